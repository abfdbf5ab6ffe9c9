"""Oracle: restatement of the reference's ejection-fraction post-processing.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows, line by line, ``compute_ef_using_putative_clips`` (reference src/fuse_utils.py:105-147), ``EDESpairs``
(src/echonet_dataset.py:159-172) and ``get2dPucks`` (src/utils/echo_utils.py:259-334, the non-visualising part).
Pinning: compared live against the unmodified reference functions in the build container
(tests/test_oracle_vs_reference.py::test_ef_matches_reference_functions) and against golden vectors produced by them
(tests/golden/ef.npz, oracle/make_golden.py) - with ONE substitution on both sides: ``skimage.segmentation.
find_boundaries(mode="thick")`` (skimage is not installed here and cannot be) is replaced by ``find_boundaries_thick``
below, a restatement of skimage's definition (grey dilation != grey erosion with the 1-connected footprint, borders
ignored).  That one function is therefore **parity unpinned**; everything around it is pinned.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi
from scipy.signal import find_peaks


def find_boundaries_thick(label_img):
    """skimage.segmentation.find_boundaries(label_img, connectivity=1, mode="thick", background=0):
    ``dilation(label_img, footprint) != erosion(label_img, footprint)``, footprint = generate_binary_structure(ndim, 1);
    skimage's morphology pads with the identity of the operation, i.e. pixels outside the image are ignored (= "nearest")."""
    img = np.asarray(label_img).astype(np.uint8)
    fp = ndi.generate_binary_structure(img.ndim, 1)
    return ndi.grey_dilation(img, footprint=fp, mode="nearest") != ndi.grey_erosion(img, footprint=fp, mode="nearest")


def edes_pairs(diastole, systole):
    """src/echonet_dataset.py:159-172."""
    diastole = np.sort(np.array(diastole))
    systole = np.sort(np.array(systole))
    clips = []
    inds = np.searchsorted(diastole, systole, side="left")
    for i, sf in enumerate(systole):
        if inds[i] == 0:                      # :165 no prior diastolic frame
            continue
        best_df = diastole[inds[i] - 1]       # :167
        if len(clips) == 0 or best_df != clips[-1][0]:
            clips.append((best_df, sf))
    return clips


def get_2d_pucks(abin, apix, npucks=10):
    """src/utils/echo_utils.py:259-334: (length of the long axis, npucks disk radii)."""
    if not np.any(abin):                                         # :266
        return 1.0, np.zeros((npucks,))
    x, y = np.where(abin > 0)                                    # :269
    X = np.stack([x, y])
    X = np.multiply(X, np.array(apix)[:, None])                  # :274
    try:
        val, vec = np.linalg.eig(np.cov(X, rowvar=True))         # :276
    except Exception:
        return 0.0, np.zeros((npucks,))
    eigorder = np.argsort(val)[-1::-1]                           # :281
    vec = vec[:, eigorder]
    if vec[0, 0] < 0:                                            # :288-291
        vec[:, 0] = -1.0 * vec[:, 0]
    if vec[1, 1] < 0:
        vec[:, 1] = -1.0 * vec[:, 1]
    mu = np.expand_dims(np.mean(X, axis=1), axis=1)              # :294
    B = find_boundaries_thick(abin)                              # :301 (substituted, see module docstring)
    Xb = np.stack(np.where(B))
    Xb = np.multiply(Xb, np.array(apix)[:, None])
    proj = np.dot((Xb - mu).T, vec)                              # :304
    L_min, L_max = np.min(proj, axis=0), np.max(proj, axis=0)    # :309
    L = L_max - L_min
    part = np.linspace(L_min[0], L_max[0], npucks + 1)           # :313
    R = []
    for i in range(len(part) - 1):
        which = np.logical_and(proj[:, 0] >= part[i], proj[:, 0] < part[i + 1])
        with np.errstate(all="ignore"):
            R.append(np.median(np.abs(proj[:, 1][which])) if which.any() else np.nan)   # :327-330 (median of nothing is nan)
    return L[0], np.array(R)


def compute_ef_using_putative_clips(fused_segmentations, test_pat_index="", return_edes=False):
    """src/fuse_utils.py:105-147 (the frame size is taken from the input; the reference hard-codes 112 at :124)."""
    size = np.sum(fused_segmentations, axis=(1, 2)).ravel()      # :106
    _05cut, _85cut, _95cut = np.percentile(size, [5, 85, 95])
    trim_range = _95cut - _05cut                                 # :109-111
    systole = find_peaks(-size, distance=20, prominence=(0.50 * trim_range))[0]
    diastole = find_peaks(size, distance=20, prominence=(0.50 * trim_range))[0]
    diastole = [x for x in diastole if size[x] >= _85cut]        # :116
    if np.mean(size[:3]) >= _85cut:                              # :118
        diastole = [0] + diastole
    diastole = np.array(diastole)
    clip_pairs = edes_pairs(diastole, systole)                   # :122
    frames = fused_segmentations.reshape((-1,) + tuple(fused_segmentations.shape[-2:]))
    predicted_efs = []
    for ed, es in clip_pairs:
        length_ed, radius_ed = get_2d_pucks((frames[ed] == 1).astype("int"), (1.0, 1.0))     # :132
        length_es, radius_es = get_2d_pucks((frames[es] == 1).astype("int"), (1.0, 1.0))
        edv = np.sum(((np.pi * radius_ed * radius_ed) * length_ed / len(radius_ed)))        # :135
        esv = np.sum(((np.pi * radius_es * radius_es) * length_es / len(radius_es)))
        ef_predicted = (edv - esv) / edv * 100
        if ef_predicted < 0:                                     # :140
            continue
        predicted_efs.append(ef_predicted)
    if return_edes:
        return predicted_efs, clip_pairs
    return predicted_efs


def beating_masks(num_frames=150, size=112, period=47.0, seed=0):
    """Synthetic multi-heartbeat LV masks for the EF fixtures: a tilted ellipse whose axes breathe with `period` frames."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float64)
    cy, cx, tilt = size * 0.52, size * 0.47, 0.35
    masks = np.zeros((num_frames, size, size), dtype=np.int64)
    for t in range(num_frames):
        phase = 0.5 * (1 + np.cos(2 * np.pi * t / period))       # 1 at end-diastole
        a = size * (0.20 + 0.10 * phase) + rng.normal(0, 0.15)
        b = size * (0.11 + 0.06 * phase) + rng.normal(0, 0.15)
        u = (yy - cy) * np.cos(tilt) + (xx - cx) * np.sin(tilt)
        v = -(yy - cy) * np.sin(tilt) + (xx - cx) * np.cos(tilt)
        masks[t] = ((u / a) ** 2 + (v / b) ** 2 <= 1.0)
    return masks
