"""Drop-in for the reference's ``src/fuse_utils.py`` on libclasfv_b200.

Same three entry points and argument meanings:

* ``divide_to_consecutive_clips(video, clip_length=32, interpolate_last=False)``            (reference :16-33)
* ``segment_a_video_with_fusion(video, model, interpolate_last=True, step=1, num_clips=10,
                               fuse_method="simple", class_list=[0, 1])``                    (reference :36-102)
* ``compute_ef_using_putative_clips(fused_segmentations, test_pat_index, return_edes=False)``  (reference :105-147)

What changed underneath: the video is uploaded once, every shifted pass is resampled and cut into
clips on the device, all clips of all passes go through the network in batches (softmax fused into
the head kernel), and resample-back + argmax + per-frame voting are one kernel; only the final
(T,H,W) mask comes back.  H and W are taken from the input (the reference hard-codes 112).

``fuse_method``:
  "simple" / "itkvoting" / "majority"  reference-exact shifted-pass fusion with per-pixel majority voting
                (LabelFusion's SIMPLE is unpinned - package not vendored; for binary labels it reduces to
                 majority voting up to its iterative re-weighting; ties go to background)
  "warp"        the north-star warp-and-fuse operator: every ``step``-strided 32-frame window of the video,
                soft votes, each frame also voting on its neighbours along the predicted forward / backward
                motion fields (specified by oracle/fuse_ref.py:warp_fuse); ``num_clips`` is ignored
"""
from __future__ import annotations

import numpy as np
import torch

from .. import engine as _engine
from .._lib import OUT_LVPROB, OUT_PROB, ClasfvError
from .echonet_dataset import EDESpairs

CLIP = 32
MAJORITY_METHODS = ("simple", "itkvoting", "majority")

_default_engines = {}


def _default_engine():
    if not torch.cuda.is_available():
        raise ClasfvError("clasfv_b200.fuse_utils needs a CUDA device (sm_100a); there is no CPU path")
    idx = torch.cuda.current_device()
    if idx not in _default_engines:
        _default_engines[idx] = _engine.Engine(torch.device("cuda", idx))
    return _default_engines[idx]


def _unwrap(model):
    from .model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
    net = model.module if isinstance(model, torch.nn.DataParallel) else model
    if not isinstance(net, R2plus1D_18_MotionNet):
        raise TypeError("segment_a_video_with_fusion needs a clasfv_b200 R2plus1D_18_MotionNet "
                        "(optionally wrapped in nn.DataParallel); got " + type(net).__name__)
    if net.training:
        raise ClasfvError("call model.eval() first (inference only)")
    return net


def _num_clips(length, clip_length=CLIP):
    return int(np.round(length / clip_length))      # half to even, as the reference (:21, :29)


def _shift_plan_entry(start, length, clip_length, interpolate_last):
    """(start, effective length, nclips) of one shifted pass, or raises what the reference raises."""
    n = _num_clips(length, clip_length)
    if length % clip_length != 0 and not interpolate_last:
        if n * clip_length > length:
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly "
                             "(last clip is short; the reference fails in np.concatenate, fuse_utils.py:32)")
        return (start, n * clip_length, n)          # truncated, no resample
    if length % clip_length != 0 and n == 0:
        raise RuntimeError("Input and output sizes should be greater than 0 (video shorter than half a clip)")
    return (start, length, n)


def _to_device_video(video, device):
    if isinstance(video, torch.Tensor):
        v = video
    else:
        v = torch.from_numpy(np.ascontiguousarray(video))
    if v.dim() != 4 or v.shape[0] != 3:
        raise ClasfvError(f"expected a video of shape (3,T,H,W), got {tuple(v.shape)}")
    return v.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def _mask_to_host_int64(mask):
    """(T,H,W) uint8 CUDA mask -> int64 numpy array (what the reference returns, fuse_utils.py:100-102).  The widening
    runs on the device and the copy lands in pinned memory from torch's caching host allocator: a uint8 pageable copy
    followed by a host-side ``astype`` costs ~2 ms per 200-frame video, more than the whole fusion kernel."""
    wide = mask.to(torch.int64)
    host = torch.empty(wide.shape, dtype=torch.int64, pin_memory=True)
    host.copy_(wide, non_blocking=True)
    torch.cuda.current_stream(mask.device).synchronize()
    return host.numpy()


def divide_to_consecutive_clips(video, clip_length=32, interpolate_last=False):
    """(3, L, H, W) array -> (n, 3, clip_length, H, W) float64 array, n = round(L / clip_length)."""
    eng = _default_engine()
    length = video.shape[1]
    entry = _shift_plan_entry(0, length, clip_length, interpolate_last)
    if entry[2] == 0:
        return np.empty((0, 3, clip_length) + tuple(video.shape[2:]), dtype=np.float64)
    v = _to_device_video(video, eng.device)
    clips = eng.build_shift_clips(v, [entry], clip_length)
    return clips.cpu().numpy().astype(np.float64)


def plan_shifts(num_frames, step=1, num_clips=10):
    """The shift distances the reference uses (:38-45), including its clamping and its message."""
    if num_frames < CLIP + num_clips * step:
        num_clips = (num_frames - CLIP) // step
    if num_clips < 0:
        print("Video is too short")
        num_clips = 1
    return list(range(0, num_clips * step, step))


def segment_a_video_with_fusion(video, model, interpolate_last=True, step=1, num_clips=10,
                                fuse_method="simple", class_list=[0, 1], batch_clips=192, edge_hops=False,
                                return_details=False):
    net = _unwrap(model)
    eng = net.engine()
    if list(class_list) != [0, 1]:
        raise ClasfvError("class_list must be [0, 1] (background, LV)")
    method = fuse_method.lower()
    v = _to_device_video(video, eng.device)
    num_frames, h, w = int(v.shape[1]), int(v.shape[2]), int(v.shape[3])
    out_dtype = _engine.storage_dtype(eng.precision)

    if method == "warp":
        if num_frames < CLIP:
            raise ClasfvError("warp fusion needs at least one full 32-frame clip")
        starts = list(range(0, num_frames - CLIP + 1, step))
        if starts[-1] != num_frames - CLIP:
            starts.append(num_frames - CLIP)        # the tail is always covered
        n = len(starts)
        prob = torch.empty((n, 1, CLIP, h, w), dtype=out_dtype, device=v.device)      # the LV probability: all that fusion reads
        mot = torch.empty((n, 4, CLIP, h, w), dtype=out_dtype, device=v.device)
        # one call per run of equally spaced windows: the library batches internally (batch_clips) and shares the stem
        # and layer1 between the overlapping windows (dense-video schedule, csrc/api.cu)
        eng.forward_windows(v, prob, mot, OUT_LVPROB, starts, CLIP, batch_clips)
        res = eng.warp_fuse(prob, mot, starts, num_frames, edge_hops=edge_hops)
        fused = _mask_to_host_int64(res["mask"])
        if return_details:
            return fused, {"area": res["area"].cpu().numpy(), "cnt": res["cnt"].cpu().numpy(), "acc": res["acc"],
                           "clips": n, "prob": prob, "motion": mot, "starts": starts}
        return fused

    if method == "staple":
        raise NotImplementedError("STAPLE label fusion (LabelFusion/SimpleITK) is not part of this path")
    if method not in MAJORITY_METHODS:
        raise ValueError(f"unknown fuse_method {fuse_method!r}")

    shifts = plan_shifts(num_frames, step, num_clips)
    if not shifts:
        raise IndexError("list index out of range")     # reference: all_interpolated_segmentations[0] (:82)
    plan = [_shift_plan_entry(s, num_frames - s, CLIP, interpolate_last) for s in shifts]
    if plan[0][2] == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    for k, (_s, eff_len, _n) in enumerate(plan):        # frames a shift must provide (reference :87-91)
        last_needed = (num_frames - 1) - k * step
        if k < min(num_frames - 1, len(plan)) or k == 0:
            if last_needed >= eff_len and last_needed >= 0:
                raise IndexError(f"index {last_needed} is out of bounds for axis 0 with size {eff_len}")
    clips = eng.build_shift_clips(v, plan, CLIP)
    total = clips.shape[0]
    prob = torch.empty((total, 2, CLIP, h, w), dtype=out_dtype, device=v.device)
    mot = torch.empty((min(batch_clips, total), 4, CLIP, h, w), dtype=out_dtype, device=v.device)
    eng.set_option("sub_batch", batch_clips)
    for b0 in range(0, total, batch_clips):
        b1 = min(total, b0 + batch_clips)
        eng.forward_into(clips[b0:b1], prob[b0:b1], mot[:b1 - b0], OUT_PROB)
    mask, area = eng.fuse_shift_votes(prob, plan, num_frames, step)
    fused = _mask_to_host_int64(mask)
    keep = [0] + [i for i in range(1, num_frames) if step - 1 < i]   # reference skips frames 1..step-1 (:85)
    if len(keep) != num_frames:
        fused = fused[keep]
    if return_details:
        return fused, {"area": area.cpu().numpy()[keep], "clips": total, "plan": plan, "prob": prob}
    return fused


class _VideoPipeline:
    """State of segment_videos_with_fusion: two pinned staging slots and two device video slots, one copy stream per
    direction; prob / motion planes of the video in flight are reused when consecutive videos have the same shape."""

    def __init__(self, eng):
        self.eng = eng
        # one stream per direction: on a single in-order copy stream the upload of video i+1 would queue behind the mask
        # download of video i, which waits for video i's compute - and the GPU would idle for an upload per video
        self.copy_stream = torch.cuda.Stream(device=eng.device)          # host -> device
        self.back_stream = torch.cuda.Stream(device=eng.device)          # device -> host
        self.stage = [None, None]
        self.dev_video = [None, None]
        self.uploaded = [None, None]         # event: the slot's last host->device copy (its pinned buffer is reusable after it)
        self.consumed = [None, None]         # event: the compute that read the slot's device video has finished
        self.planes = None

    def upload(self, slot, video):
        """host (3,T,H,W) float array -> device, on the copy stream; returns (device tensor, event)."""
        src = video if isinstance(video, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(video))
        if src.dim() != 4 or src.shape[0] != 3:
            raise ClasfvError(f"expected a video of shape (3,T,H,W), got {tuple(src.shape)}")
        n = src.numel()
        if self.stage[slot] is None or self.stage[slot].numel() < n:
            self.stage[slot] = torch.empty(n, dtype=torch.float32, pin_memory=True)
        if self.dev_video[slot] is None or self.dev_video[slot].numel() < n:
            self.dev_video[slot] = torch.empty(n, dtype=torch.float32, device=self.eng.device)
        host = self.stage[slot][:n].view(src.shape)
        dev = self.dev_video[slot][:n].view(src.shape)
        if self.uploaded[slot] is not None:
            self.uploaded[slot].synchronize()    # the pinned buffer is still the source of an earlier copy until then
        if not src.is_cuda:
            host.copy_(src)                      # float64 / uint8 inputs are converted here
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[slot] is not None:
                self.copy_stream.wait_event(self.consumed[slot])   # the video that lived in this slot is no longer being read
            dev.copy_(src if src.is_cuda else host, non_blocking=True)
            ev = torch.cuda.Event(); ev.record()
        self.uploaded[slot] = ev
        return dev, ev, slot

    def plane_buffers(self, n, h, w, dtype):
        """(n,1,32,h,w) LV-probability and (n,4,32,h,w) motion planes as views of two flat buffers that only ever grow:
        videos of different lengths reuse them instead of sending a new size through the allocator per video."""
        need = n * CLIP * h * w
        if self.planes is None or self.planes[0] != dtype or self.planes[1].numel() < need:
            self.planes = None               # release before allocating the larger buffers
            dev = self.eng.device
            self.planes = (dtype, torch.empty(need, dtype=dtype, device=dev), torch.empty(4 * need, dtype=dtype, device=dev))
        return self.planes[1][:need].view(n, 1, CLIP, h, w), self.planes[2][:4 * need].view(n, 4, CLIP, h, w)


def segment_videos_with_fusion(videos, model, step=1, batch_clips=192, edge_hops=False, return_details=False):
    """Warp-and-fuse segmentation of MANY videos (BASELINE config 4; the many-video counterpart of
    ``segment_a_video_with_fusion(video, model, fuse_method="warp")``, reference src/fuse_utils.py:36-102): a generator that
    takes an iterable of host (3,T,H,W) float arrays and yields one (T,H,W) int64 mask per video, in order, identical to
    the single-video call.  Host and device work overlap across videos: while video i is segmented, video i+1 is staged
    through pinned memory and uploaded on a copy stream, and the mask of video i-1 travels back; every video still pays its
    own host->device and device->host copy.  With ``return_details`` yields (mask, {"area": per-frame LV pixel count})."""
    net = _unwrap(model)
    eng = net.engine()
    out_dtype = _engine.storage_dtype(eng.precision)
    pipe = _VideoPipeline(eng)
    main = torch.cuda.current_stream(eng.device)
    it = iter(videos)

    def start(slot, video):
        return pipe.upload(slot, video)

    def compute(dev_video, ready, slot):
        main.wait_event(ready)
        num_frames, h, w = int(dev_video.shape[1]), int(dev_video.shape[2]), int(dev_video.shape[3])
        if num_frames < CLIP:
            raise ClasfvError("warp fusion needs at least one full 32-frame clip")
        starts = list(range(0, num_frames - CLIP + 1, step))
        if starts[-1] != num_frames - CLIP:
            starts.append(num_frames - CLIP)
        prob, mot = pipe.plane_buffers(len(starts), h, w, out_dtype)
        eng.forward_windows(dev_video, prob, mot, OUT_LVPROB, starts, CLIP, batch_clips)
        res = eng.warp_fuse(prob, mot, starts, num_frames, edge_hops=edge_hops)
        wide = res["mask"].to(torch.int64)
        done = torch.cuda.Event(); done.record(main)
        pipe.consumed[slot] = done
        host = torch.empty(wide.shape, dtype=torch.int64, pin_memory=True)
        area = torch.empty(res["area"].shape, dtype=torch.int32, pin_memory=True) if return_details else None
        with torch.cuda.stream(pipe.back_stream):
            pipe.back_stream.wait_event(done)
            host.copy_(wide, non_blocking=True)
            if area is not None:
                area.copy_(res["area"], non_blocking=True)
            wide.record_stream(pipe.back_stream)
            res["area"].record_stream(pipe.back_stream)
            back = torch.cuda.Event(); back.record()
        return host, area, back

    def finish(pending):
        host, area, back = pending
        back.synchronize()
        return (host.numpy(), {"area": area.numpy()}) if return_details else host.numpy()

    try:
        nxt = start(0, next(it))
    except StopIteration:
        return
    slot, pending = 0, None
    while nxt is not None:
        cur = nxt
        out = compute(*cur)                         # enqueue video i on the main stream
        try:
            nxt = start(1 - slot, next(it))         # stage + upload video i+1 while the GPU works on video i
        except StopIteration:
            nxt = None
        if pending is not None:
            yield finish(pending)                   # mask of video i-1 (its copy overlapped video i's compute)
        pending, slot = out, 1 - slot
    yield finish(pending)


# ------------------------------------------------------------------------------------------- EF (host)
# Ejection fraction from the fused masks (reference src/fuse_utils.py:105-147 with get2dPucks, src/utils/echo_utils.py:259-334):
# the LV area trace picks end-diastolic / end-systolic frames, each LV mask becomes a stack of disks along its long axis
# (Simpson, single plane), EF = (EDV - ESV) / EDV.  Same results as the reference functions (oracle/ef_ref.py restates them
# line by line; tests compare the two on the same masks), organised around the device-side by-product of fusion: both
# fusion kernels already return the per-frame LV area, so the (T,H,W) mask is only touched for the few ED / ES frames.
def _find_boundaries_thick(binary):
    """Pixels where the 4-neighbourhood (plus the pixel) is not constant: skimage.segmentation.find_boundaries(mode="thick")
    for a 2-D label image; neighbours outside the image do not count."""
    from scipy import ndimage as ndi
    img = np.asarray(binary).astype(np.uint8)
    cross = ndi.generate_binary_structure(2, 1)
    return ndi.maximum_filter(img, footprint=cross, mode="nearest") != ndi.minimum_filter(img, footprint=cross, mode="nearest")


def _long_axis_frame(points):
    """Principal axes (columns, long axis first) of 2 x N pixel coordinates, oriented so that the long axis points to +row and
    the short axis to +column - the reference's sign convention."""
    weights, axes = np.linalg.eig(np.cov(points, rowvar=True))
    axes = axes[:, np.argsort(weights)[::-1]]
    for k in (0, 1):
        if axes[k, k] < 0:
            axes[:, k] = -axes[:, k]
    return axes


def get2dPucks(abin, apix, npucks=10):
    """(long-axis length, ``npucks`` disk radii) of a binary LV mask with pixel spacing ``apix`` (reference
    src/utils/echo_utils.py:259-334): boundary pixels are projected on the mask's principal axes, the long-axis extent of the
    boundary is cut into ``npucks`` equal slabs and a slab's radius is the median distance of its boundary pixels from the
    long axis (nan for a slab without boundary pixels).  An empty mask gives (1.0, zeros)."""
    filled = np.asarray(abin) > 0
    if not filled.any():
        return 1.0, np.zeros((npucks,))
    spacing = np.asarray(apix, dtype=np.float64).reshape(2, 1)
    inside = np.vstack(np.nonzero(filled)) * spacing
    try:
        axes = _long_axis_frame(inside)
    except Exception:
        return 0.0, np.zeros((npucks,))
    centre = inside.mean(axis=1, keepdims=True)
    edge = np.vstack(np.nonzero(_find_boundaries_thick(filled))) * spacing
    along, across = np.dot((edge - centre).T, axes).T
    lo, hi = along.min(), along.max()
    cuts = np.linspace(lo, hi, npucks + 1)
    radii = np.full((npucks,), np.nan)
    for k in range(npucks):
        slab = (along >= cuts[k]) & (along < cuts[k + 1])
        if slab.any():
            radii[k] = np.median(np.abs(across[slab]))
    return hi - lo, radii


def _disk_volume(mask):
    length, radii = get2dPucks((mask == 1).astype("int"), (1.0, 1.0))
    return np.sum(np.pi * radii * radii * length / len(radii))


def find_ed_es_frames(area):
    """[(ED frame, ES frame)] from the per-frame LV area: peaks / troughs at least 20 frames apart whose prominence is half
    the 5-95 percentile range; diastoles must reach the 85th percentile; frame 0 counts as a diastole when the video starts
    near one (reference src/fuse_utils.py:106-122)."""
    from scipy.signal import find_peaks
    area = np.asarray(area).ravel()
    p05, p85, p95 = np.percentile(area, [5, 85, 95])
    prominence = 0.5 * (p95 - p05)
    troughs = find_peaks(-area, distance=20, prominence=prominence)[0]
    peaks = [f for f in find_peaks(area, distance=20, prominence=prominence)[0] if area[f] >= p85]
    if np.mean(area[:3]) >= p85:
        peaks = [0] + peaks
    return EDESpairs(np.array(peaks), troughs)


def compute_ef_using_putative_clips(fused_segmentations, test_pat_index, return_edes=False, area=None):
    """Ejection fraction (%) at every identified heartbeat of a fused (T,H,W) mask video.  ``area``: the per-frame LV pixel
    count when the caller already has it (``return_details`` of segment_a_video_with_fusion: a by-product of the fusion
    kernel); otherwise the masks are summed here as the reference does."""
    frames = fused_segmentations.reshape((-1,) + tuple(fused_segmentations.shape[-2:]))
    if area is None:
        area = np.sum(fused_segmentations, axis=(1, 2)).ravel()
    clip_pairs = find_ed_es_frames(area)
    predicted_efs = []
    for ed, es in clip_pairs:
        edv, esv = _disk_volume(frames[ed]), _disk_volume(frames[es])
        ef = (edv - esv) / edv * 100
        if ef < 0:
            print("Negative EF at patient: " + str(test_pat_index))
            continue
        predicted_efs.append(ef)
    if return_edes:
        return predicted_efs, clip_pairs
    return predicted_efs
