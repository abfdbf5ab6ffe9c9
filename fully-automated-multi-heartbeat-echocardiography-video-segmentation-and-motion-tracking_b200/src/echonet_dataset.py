"""The two helpers of the reference's ``src/echonet_dataset.py`` that the inference path uses
(``zeroone_normalizer`` :38-50 and ``EDESpairs`` :159-172); the dataset classes need the EchoNet-Dynamic
data and the ``echonet`` package and are out of scope."""
import numpy as np


def zeroone_normalizer(image_data):
    """Per-channel ``x -= min; x /= max`` over a (3, ...) array, in place (reference semantics: the
    divisor is the max *after* the shift)."""
    norm_data = image_data
    data_shape = norm_data.shape
    norm_data = norm_data.reshape(3, -1)
    norm_data -= np.min(norm_data, axis=1).reshape(3, 1)
    norm_data /= np.max(norm_data, axis=1).reshape(3, 1)
    return norm_data.reshape(data_shape)


def EDESpairs(diastole, systole):
    """Pair every systolic frame with the closest preceding diastolic frame, one pair per diastole."""
    diastole = np.sort(np.array(diastole))
    systole = np.sort(np.array(systole))
    clips = []
    inds = np.searchsorted(diastole, systole, side='left')
    for i, sf in enumerate(systole):
        if inds[i] == 0:
            continue
        best_df = diastole[inds[i] - 1]
        if len(clips) == 0 or best_df != clips[-1][0]:
            clips.append((best_df, sf))
    return clips
