"""Oracle: restatement of the reference's full-video fusion pipeline and warp primitive.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows (paths relative to the reference checkout):

* ``src/fuse_utils.py:16-33``   divide_to_consecutive_clips          (C1)
* ``src/fuse_utils.py:36-102``  segment_a_video_with_fusion          (F1)
* ``src/transform_utils.py:14-34`` generate_2dmotion_field           (W1)
* ``src/clasfv_losses.py:86-87,112-113`` the grid_sample call-site kwargs
  (``align_corners=False, mode="bilinear", padding_mode="border"``)
* ``src/clasfv_losses.py:84-94,110-120`` flow direction conventions
* ``src/echonet_dataset.py:38-50`` zeroone_normalizer                (P1)
* ``src/clasfv_losses.py:60-68`` categorical_dice

Deliberate differences from the literal files (none changes a value):
H/W are taken from the input instead of the hard-coded 112 (``fuse_utils.py:22,27,55,68,75``),
ragged per-shift clip lists stay Python lists (``np.array(ragged)`` at ``:50`` raises on
NumPy >= 1.24), ``.cuda()`` is dropped from ``transform_utils.py:19-20``, the model runs under
``no_grad``.  The label voter (``LabelFusion.wrapper.fuse_images``, ``fuse_utils.py:95``; package
not vendored, not installed, version unpinned) is restated as plain majority voting with
ties going to class 0: PARITY UNPINNED for its SIMPLE / STAPLE methods.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

CLIP = 32


# --------------------------------------------------------------------------- P1
def zeroone_normalizer(image_data):
    """src/echonet_dataset.py:38-50 - per-channel (x - min) / max(x - min), in place."""
    shape = image_data.shape
    flat = image_data.reshape(3, -1)
    flat -= np.min(flat, axis=1).reshape(3, 1)
    flat /= np.max(flat, axis=1).reshape(3, 1)
    return flat.reshape(shape)


# --------------------------------------------------------------------------- C1
def temporal_resample(x, out_len):
    """Linear temporal resample with align_corners=False of a (C, L, H, W) array.

    The reference does this with 5-D ``F.interpolate(mode="trilinear",
    align_corners=False)`` whose spatial output size equals the input size
    (fuse_utils.py:21-23 and :74-76); with equal sizes the spatial part of the
    trilinear kernel is the identity, so only the time axis is interpolated.
    """
    t = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).unsqueeze(0)
    t = F.interpolate(t, size=(int(out_len), t.shape[3], t.shape[4]), mode="trilinear", align_corners=False)
    return t.squeeze(0).numpy()


def temporal_resample_formula(x, out_len):
    """Same as :func:`temporal_resample`, written out (what the CUDA kernel computes)."""
    x = np.asarray(x, dtype=np.float32)
    L = x.shape[1]
    scale = np.float32(L) / np.float32(out_len)
    d = np.arange(out_len, dtype=np.float32)
    src = np.maximum(scale * (d + np.float32(0.5)) - np.float32(0.5), np.float32(0)).astype(np.float32)
    i0 = np.minimum(np.floor(src).astype(np.int64), L - 1)
    i1 = np.minimum(i0 + 1, L - 1)
    lam1 = (src - i0.astype(np.float32)).astype(np.float32)
    lam0 = (np.float32(1) - lam1).astype(np.float32)
    return (lam0[None, :, None, None] * x[:, i0] + lam1[None, :, None, None] * x[:, i1]).astype(np.float32)


def num_consecutive_clips(length, clip_length=CLIP):
    """``int(np.round(L / clip_length))`` - NumPy rounds half to even (fuse_utils.py:21,29)."""
    return int(np.round(length / clip_length))


def divide_to_consecutive_clips(video, clip_length=CLIP, interpolate_last=False):
    """src/fuse_utils.py:16-33.  (3, L, H, W) -> (n, 3, clip_length, H, W) float64."""
    source_video = video.copy()
    length = video.shape[1]
    n = num_consecutive_clips(length, clip_length)
    if length % clip_length != 0 and interpolate_last:
        if n == 0:
            raise RuntimeError("Input and output sizes should be greater than 0")  # F.interpolate size 0
        source_video = temporal_resample(source_video, n * clip_length)
    clips = []
    for start in range(0, clip_length * n, clip_length):
        one_clip = source_video[:, start:start + clip_length]
        if one_clip.shape[1] != clip_length:   # np.concatenate shape mismatch in the reference (:32)
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
        clips.append(one_clip)
    if not clips:
        return np.empty((0, 3, clip_length) + video.shape[2:], dtype=np.float64)
    return np.stack(clips).astype(np.float64)


# --------------------------------------------------------------------------- voter
def majority_vote(masks):
    """Per-pixel majority over binary masks; ties -> class 0.  Stand-in for
    ``fuse_images(images, "simple", class_list=[0,1])`` (fuse_utils.py:95) - unpinned."""
    stack = np.stack([np.asarray(m).astype(np.int64) for m in masks])
    ones = stack.sum(0)
    return (2 * ones > stack.shape[0]).astype(np.uint8)


# --------------------------------------------------------------------------- F1
def plan_shifts(num_frames, step=1, num_clips=10):
    """fuse_utils.py:38-45 - the list of shift distances actually used."""
    if num_frames < CLIP + num_clips * step:
        num_clips = (num_frames - CLIP) // step
    if num_clips < 0:
        print("Video is too short")
        num_clips = 1
    return list(range(0, num_clips * step, step))


def shifted_softmax(video, model, interpolate_last=True, step=1, num_clips=10):
    """fuse_utils.py:45-68 - per shift, the concatenated per-frame softmax (2, 32 n_k, H, W)."""
    out = []
    for shift in plan_shifts(video.shape[1], step, num_clips):
        clips = divide_to_consecutive_clips(video[:, shift:], interpolate_last=interpolate_last)
        probs = []
        for i in range(clips.shape[0]):
            with torch.no_grad():
                seg, _motion = model(torch.Tensor(np.expand_dims(clips[i], 0)))   # :59
                probs.append(F.softmax(seg, 1).cpu().numpy().astype(np.float64))  # :60-61
        p = np.concatenate(probs)                                # (n, 2, 32, H, W)
        p = p.transpose([1, 0, 2, 3, 4]).reshape(2, -1, p.shape[3], p.shape[4])  # :66-68
        out.append(p)
    return out


def segment_a_video_with_fusion(video, model, interpolate_last=True, step=1, num_clips=10,
                                fuse_method="simple", class_list=(0, 1), voter=majority_vote):
    """src/fuse_utils.py:36-102.  (3, T, H, W) float in [0,1] -> (T, H, W) int64 {0,1}."""
    all_seg = shifted_softmax(video, model, interpolate_last, step, num_clips)
    masks = []
    for i, seg in enumerate(all_seg):                              # :70-80
        length = video[:, i * step:].shape[1]
        if interpolate_last and length % CLIP != 0:
            seg = temporal_resample(seg, length)
        masks.append(np.argmax(seg, 0))
    fused = [masks[0][0]]                                          # :82
    for i in range(1, video.shape[1]):                             # :84-98
        if step - 1 < i:
            voters = []
            for index in range(min(i, len(masks))):
                if i - index * step < 0:
                    break
                voters.append(masks[index][i - index * step].astype("uint8"))
            fused.append(voters[0] if len(voters) <= 1 else voter(voters).astype("uint8"))
    return np.array(fused).astype(np.int64)


# --------------------------------------------------------------------------- W1
def generate_2dmotion_field(x, offset):
    """src/transform_utils.py:14-34 without the hard-coded ``.cuda()``.

    x: (N, C, H, W); offset: (N, 2, H, W), channel 0 added to the column (x) grid and
    channel 1 to the row (y) grid.  Returns the (N, H, W, 2) grid_sample grid.
    """
    h, w = int(x.shape[2]), int(x.shape[3])
    grid_w, grid_h = torch.meshgrid([torch.linspace(-1, 1, h), torch.linspace(-1, 1, w)], indexing="ij")
    grid_w = grid_w.to(offset.device).float()   # row coordinate (despite the name)
    grid_h = grid_h.to(offset.device).float()   # column coordinate
    offset_h, offset_w = torch.split(offset, 1, 1)
    offset_w = offset_w.contiguous().view(-1, h, w)
    offset_h = offset_h.contiguous().view(-1, h, w)
    return torch.stack((grid_h + offset_h, grid_w + offset_w), 3)


def warp(src, flow):
    """W1 + its call site (clasfv_losses.py:86-87): out = bilinear src sampled at grid+flow."""
    return F.grid_sample(src, generate_2dmotion_field(src, flow), align_corners=False,
                         mode="bilinear", padding_mode="border")


# --------------------------------------------------------------------------- F2
def apply_sequence_deformation(flow_source_image, motion_output, start_index, end_index, grid_mode="bilinear", forward=True):
    """Reference ``src/visualization_utils.py:106-128`` restated on the pinned primitives (``generate_2dmotion_field`` +
    ``F.grid_sample(align_corners=False, mode=grid_mode, padding_mode='border')``): chained warps of one frame's image /
    label along the forward (channels 0:2) or backward (2:4) motion of frames range(start_index, end_index, +-1)."""
    step = 1 if forward else -1
    for frame_index in range(start_index, end_index, step):
        field = motion_output[:, :2, frame_index] if forward else motion_output[:, 2:, frame_index]
        grid = generate_2dmotion_field(flow_source_image, field)
        new_image = F.grid_sample(flow_source_image, grid, align_corners=False, mode=grid_mode, padding_mode="border")
        flow_source_image = new_image
    return new_image


def warp_fuse(prob, motion, clip_starts, num_frames, edge_hops=False, accumulate=torch.float64):
    """North-star warp-and-fuse operator (not in the reference; composition of S1 + W1).

    prob   (n, 2, 32, H, W) per-frame class probabilities of clip c (softmax of the seg logits), or (n, 1, 32, H, W) the
           LV probability alone: only the LV plane (the last) is used
    motion (n, 4, 32, H, W) tanh motion, channels [fwd x, fwd y, bwd x, bwd y]
    clip c covers global frames clip_starts[c] + t, t in [0, 32).

    Each clip frame t casts up to three votes, all *gathers* in the destination frame:
      direct   onto g = s_c + t      : prob[c, :, t]
      forward  onto g + 1            : warp(prob[c, :, t], motion[c, 0:2, t])   (t -> t+1, clasfv_losses.py:84-94)
      backward onto g - 1            : warp(prob[c, :, t], motion[c, 2:4, t])   (t -> t-1, clasfv_losses.py:110-120)
    With ``edge_hops=False`` (default) a hop must land inside its own clip (forward from
    t <= 30, backward from t >= 1): these are exactly the flows the reference's losses
    supervise (clasfv_losses.py:38-40 uses fwd[t], bwd[t+1] for t in [0, 31)).  With
    ``edge_hops=True`` the two unsupervised edge flows vote as well (SURVEY.md 8a row F2
    wording); votes that land outside [0, num_frames) are dropped either way.

    The operator fuses the LV probability ("each clip's per-frame LV softmax is warped", north-star): acc[:, 1] is the sum
    of the LV votes and acc[:, 0] = cnt - acc[:, 1] the background sum - what summing the warped background plane gives as
    well, because the two class probabilities and the four bilinear weights each sum to one.  fused = LV where
    acc[:, 1] > acc[:, 0], i.e. where the mean LV probability exceeds 1/2; ties -> background.

    Returns (acc (T, 2, H, W), cnt (T,), mask (T, H, W) uint8).
    """
    n, _, clip, h, w = prob.shape
    lv = torch.zeros(num_frames, h, w, dtype=accumulate)
    cnt = torch.zeros(num_frames, dtype=torch.int64)
    for c in range(n):
        s = int(clip_starts[c])
        for t in range(clip):
            g = s + t
            p = prob[c, -1:, t].unsqueeze(0).float()
            if 0 <= g < num_frames:
                lv[g] += p[0, 0].to(accumulate)
                cnt[g] += 1
            if (edge_hops or t + 1 < clip) and 0 <= g + 1 < num_frames:
                lv[g + 1] += warp(p, motion[c, 0:2, t].unsqueeze(0).float())[0, 0].to(accumulate)
                cnt[g + 1] += 1
            if (edge_hops or t >= 1) and 0 <= g - 1 < num_frames:
                lv[g - 1] += warp(p, motion[c, 2:4, t].unsqueeze(0).float())[0, 0].to(accumulate)
                cnt[g - 1] += 1
    acc = torch.stack([cnt.view(-1, 1, 1).to(accumulate) - lv, lv], 1)
    mask = (acc[:, 1] > acc[:, 0]).to(torch.uint8)
    return acc, cnt, mask


# --------------------------------------------------------------------------- metrics
def categorical_dice(prediction, truth, k, epsilon=1e-5):
    """src/clasfv_losses.py:60-68."""
    a = (prediction == k)
    b = (truth == k)
    return 2 * np.sum(a * b) / (np.sum(a) + np.sum(b) + epsilon)
