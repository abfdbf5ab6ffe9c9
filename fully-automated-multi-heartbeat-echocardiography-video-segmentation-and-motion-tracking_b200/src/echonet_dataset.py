"""Host helpers with the names and semantics of the two functions of the reference's ``src/echonet_dataset.py`` that the
inference path uses (``zeroone_normalizer`` :38-50, ``EDESpairs`` :159-172).  Written against their behaviour (pinned by
``tests/golden/host_helpers.npz``, produced by the reference functions), not transcribed; the dataset classes need the
EchoNet-Dynamic data and the ``echonet`` package and are out of scope.  On the product path the normalisation runs on the
device (``clasfv_ingest_u8``); this host form is what the oracle and the golden tests compare it with."""
import numpy as np


def zeroone_normalizer(image_data):
    """Scale every one of the three channels of a (3, ...) float array to [0, 1], IN PLACE, and return it: subtract the
    channel minimum, then divide by the channel maximum *of the shifted data* (so a constant channel divides by zero, as
    in the reference)."""
    flat = image_data.reshape(3, -1)                 # a view: the caller's array is modified, like the reference's
    flat -= flat.min(axis=1, keepdims=True)
    flat /= flat.max(axis=1, keepdims=True)
    return flat.reshape(image_data.shape)


def EDESpairs(diastole, systole):
    """[(ED frame, ES frame)]: every systolic frame is matched with the last diastolic frame before it; a diastole keeps only
    the first systole matched to it, and systoles with no earlier diastole are dropped.  Pairs come out in frame order."""
    ed_frames = np.sort(np.asarray(diastole))
    pairs, used = [], None
    for es in np.sort(np.asarray(systole)):
        n_before = int(np.searchsorted(ed_frames, es, side="left"))     # diastoles strictly before this systole
        if n_before == 0:
            continue
        ed = ed_frames[n_before - 1]
        if used is None or ed != used:
            pairs.append((ed, es))
            used = ed
    return pairs
