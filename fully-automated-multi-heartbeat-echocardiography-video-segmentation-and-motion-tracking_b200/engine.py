"""Host-side wrapper of one ``clasfv_handle`` (packed network + workspace on one device).

This is plumbing shared by the reference-shaped drop-ins in ``src/``: it moves state_dict tensors
across the C ABI, and exposes forward / fusion calls on torch CUDA tensors (torch supplies device
memory and streams only).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import BF16, F16, F32, OUT_LOGITS, OUT_LVPROB, OUT_PROB, ClasfvError, check


def precision_code(precision):
    if precision in (F32, "fp32", "f32", "float32", torch.float32):
        return F32
    if precision in (BF16, "bf16", "bfloat16", torch.bfloat16):
        return BF16
    if precision in (F16, "fp16", "f16", "float16", "half", torch.float16):
        return F16
    raise ClasfvError(f"unknown precision {precision!r}: use 'fp32', 'bf16' or 'fp16'")


def storage_dtype(precision):
    """torch element type of the prob / motion planes the fused pipeline keeps resident in a given precision mode."""
    return {F32: torch.float32, BF16: torch.bfloat16, F16: torch.float16}[precision_code(precision)]


class Engine:
    """One packed CLAS-FV network on one CUDA device."""

    MAX_RUN = 1024        # windows per clasfv_forward call of forward_windows
    WORKSPACE_BUDGET = 64e9   # bytes of activation workspace forward_windows lets an internal batch take

    def __init__(self, device):
        self.lib = _lib.lib()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ClasfvError(f"clasfv_b200 runs on CUDA devices only (sm_100a); got {dev}. There is no CPU path.")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        h = C.c_void_p()
        check(self.lib.clasfv_create(self.device.index, C.byref(h)), "clasfv_create")
        self._h = h
        self.precision = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.clasfv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict, precision="fp32"):
        for key, t in state_dict.items():
            if key.startswith("module."):
                key = key[7:]
            if not torch.is_floating_point(t):
                continue                      # num_batches_tracked
            a = np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())
            shape = _lib.i64_array(a.shape) if a.ndim else None
            check(self.lib.clasfv_set_tensor(self._h, key.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim),
                  f"clasfv_set_tensor({key})")
        self.finalize(precision)

    def finalize(self, precision):
        code = precision_code(precision)
        check(self.lib.clasfv_finalize(self._h, code), "clasfv_finalize")
        self.precision = code

    # ------------------------------------------------------------------ network
    def _clip_geometry(self, x, clip_starts, clip_len):
        """Validation shared by forward() and forward_into(): device, element type, contiguity, shape, and that every clip
        window lies inside the video.  Returns (n, t, h, w, clip offsets or None, channel stride)."""
        _lib.require_cuda(x, "x")
        if x.dtype != torch.float32:
            raise ClasfvError("network input must be float32")
        if clip_starts is None:
            if x.dim() != 5 or x.shape[1] != 3:
                raise ClasfvError(f"expected input of shape (N,3,T,H,W), got {tuple(x.shape)}")
            n, _, t, h, w = x.shape
            return n, t, h, w, None, 0
        if x.dim() != 4 or x.shape[0] != 3:
            raise ClasfvError(f"expected a video of shape (3,T,H,W), got {tuple(x.shape)}")
        _, tv, h, w = x.shape
        t = int(clip_len)
        n = len(clip_starts)
        if n and (min(clip_starts) < 0 or max(clip_starts) + t > tv):
            raise ClasfvError("clip window outside the video")
        return n, t, h, w, _lib.i64_array([int(s) * h * w for s in clip_starts]), tv * h * w

    def forward(self, x, out_kind=OUT_LOGITS, out_dtype=torch.float32, clip_starts=None, clip_len=None):
        """x: (N,3,T,H,W) fp32 CUDA tensor, or with ``clip_starts`` a resident video (3,Tv,H,W) whose
        windows [s, s+clip_len) are the clips (no copy).  Returns (seg, motion)."""
        n, t, h, w, offs, ch_stride = self._clip_geometry(x, clip_starts, clip_len)
        seg = torch.empty((n, 1 if out_kind == OUT_LVPROB else 2, t, h, w), dtype=out_dtype, device=x.device)
        mot = torch.empty((n, 4, t, h, w), dtype=out_dtype, device=x.device)
        if n == 0:
            return seg, mot
        check(self.lib.clasfv_forward(self._h, x.data_ptr(), offs, ch_stride, n, t, h, w, out_kind,
                                      _lib.torch_dtype_code(out_dtype), seg.data_ptr(), mot.data_ptr(),
                                      _lib.current_stream_ptr(x.device)), "clasfv_forward")
        return seg, mot

    def forward_into(self, x, seg, mot, out_kind, clip_starts=None, clip_len=None):
        """As :meth:`forward`, writing into caller-provided slices of larger (N,2,T,H,W)/(N,4,T,H,W) buffers."""
        n, t, h, w, offs, ch_stride = self._clip_geometry(x, clip_starts, clip_len)
        _lib.require_cuda(seg, "seg"); _lib.require_cuda(mot, "mot")
        planes = 1 if out_kind == OUT_LVPROB else 2
        if tuple(seg.shape) != (n, planes, t, h, w) or tuple(mot.shape) != (n, 4, t, h, w) or seg.dtype != mot.dtype or seg.device != x.device:
            raise ClasfvError(f"forward_into: outputs must be ({n},{planes},{t},{h},{w}) and ({n},4,{t},{h},{w}) tensors of one type on "
                              f"the input's device (got {tuple(seg.shape)} {seg.dtype}, {tuple(mot.shape)} {mot.dtype})")
        if n == 0:
            return
        check(self.lib.clasfv_forward(self._h, x.data_ptr(), offs, ch_stride, n, t, h, w, out_kind,
                                      _lib.torch_dtype_code(seg.dtype), seg.data_ptr(), mot.data_ptr(),
                                      _lib.current_stream_ptr(x.device)), "clasfv_forward")

    def forward_windows(self, video, seg, mot, out_kind, clip_starts, clip_len, batch_clips=64):
        """All windows [s, s+clip_len) of a resident video (3,Tv,H,W) in as few calls as possible: every maximal run
        of equally spaced starts is ONE clasfv_forward call, so the library can share the stem and layer1 between
        the overlapping windows (dense-video schedule) and batches internally (``batch_clips`` per batch)."""
        # the library's workspace grows with the internal batch: about 130 MB per 32 x 112 x 112 clip in a 16-bit mode (twice
        # that in fp32), proportional to the clip's voxels - keep it under WORKSPACE_BUDGET whatever the frame size
        per_clip = 130e6 * (clip_len / 32.0) * (video.shape[2] * video.shape[3]) / (112.0 * 112.0) * (2 if self.precision == F32 else 1)
        batch_clips = max(1, min(int(batch_clips), int(self.WORKSPACE_BUDGET // per_clip)))
        self.set_option("sub_batch", batch_clips)
        starts = [int(s) for s in clip_starts]
        i, n = 0, len(starts)
        while i < n:
            j = i + 1
            if j < n:
                d = starts[j] - starts[i]
                # a run is cut at MAX_RUN windows: the library keeps video-level maps for the frames of a run (about
                # 5.6 MB per frame at 112 x 112, 22 MB at 224 x 224), so the cap bounds that part of the workspace
                while j + 1 < n and starts[j + 1] - starts[j] == d and j + 1 - i < self.MAX_RUN:
                    j += 1
                j += 1
            self.forward_into(video, seg[i:j], mot[i:j], out_kind, clip_starts=starts[i:j], clip_len=clip_len)
            i = j

    def ingest_u8(self, frames, height, width, bgr=False):
        """frames (T,H0,W0,3) uint8 (host or CUDA) -> (3,T,height,width) fp32 CUDA video: the reference's pre-resize
        (trilinear, align_corners=True) + zero-one normalisation (motion_segment.py:96-106) on the device."""
        f = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames))
        if f.dtype != torch.uint8 or f.dim() != 4 or f.shape[3] != 3:
            raise ClasfvError(f"ingest_u8: expected uint8 frames of shape (T,H,W,3), got {f.dtype} {tuple(f.shape)}")
        if not f.is_cuda:
            f = f.pin_memory().to(self.device, non_blocking=True)
        f = f.contiguous()
        t, h0, w0, _ = f.shape
        out = torch.empty((3, t, int(height), int(width)), dtype=torch.float32, device=f.device)
        check(self.lib.clasfv_ingest_u8(self._h, f.data_ptr(), t, h0, w0, 1 if bgr else 0, out.data_ptr(), int(height), int(width),
                                        _lib.current_stream_ptr(f.device)), "clasfv_ingest_u8")
        return out

    def set_option(self, name, value):
        """``sub_batch`` (clips per internal batch of a forward call, default 16), ``dense_video`` (0/1: share
        the stem and layer1 between overlapping windows of one resident video, default on) or ``umma_pair`` (0/1: the
        convolutions of layers 2-4 on CTA pairs, default on; the outputs are the same bits either way)."""
        check(self.lib.clasfv_set_option(self._h, name.encode(), int(value)), f"clasfv_set_option({name})")

    def profile_begin(self):
        check(self.lib.clasfv_profile_begin(self._h), "clasfv_profile_begin")

    def profile_end(self):
        """-> (dict stage -> summed ms, number of forward calls)"""
        ms = (C.c_float * 4)()
        calls = C.c_int(0)
        check(self.lib.clasfv_profile_end(self._h, ms, C.byref(calls)), "clasfv_profile_end")
        gf = (C.c_double * 4)()
        check(self.lib.clasfv_profile_gflop(self._h, gf), "clasfv_profile_gflop")
        self.last_profile_gflop = {"stem": gf[0], "trunk": gf[1], "lateral": gf[2], "head": gf[3]}
        return {"stem": ms[0], "trunk": ms[1], "lateral": ms[2], "head": ms[3]}, calls.value

    def launch_count(self):
        """Kernels launched through this handle so far."""
        return int(self.lib.clasfv_launch_count(self._h))

    def workspace_bytes(self):
        return int(self.lib.clasfv_workspace_bytes(self._h))

    # ------------------------------------------------------------------ fusion
    def warp_fuse(self, prob, motion, clip_starts, num_frames, edge_hops=False, acc=None, accumulate=False,
                  want_mask=True, want_area=True, cnt=None):
        """F2 (oracle/fuse_ref.py:warp_fuse). prob (n,2,L,H,W) class probabilities or (n,1,L,H,W) the LV probability alone.
        Returns dict(acc, cnt, mask, area)."""
        _lib.require_cuda(prob, "prob"); _lib.require_cuda(motion, "motion")
        n, c, clip_len, h, w = prob.shape
        if c not in (1, 2) or tuple(motion.shape) != (n, 4, clip_len, h, w) or motion.dtype != prob.dtype:
            raise ClasfvError("warp_fuse: prob must be (n,2,L,H,W) or (n,1,L,H,W) and motion (n,4,L,H,W) of the same dtype")
        if len(clip_starts) != n:
            raise ClasfvError("warp_fuse: one start per clip")
        dev = prob.device
        if acc is None:
            acc = torch.empty((num_frames, 2, h, w), dtype=torch.float32, device=dev)
            accumulate = False
        if cnt is None:
            cnt = torch.zeros((num_frames,), dtype=torch.int32, device=dev)
        mask = torch.empty((num_frames, h, w), dtype=torch.uint8, device=dev) if want_mask else None
        area = torch.empty((num_frames,), dtype=torch.int32, device=dev) if want_area else None
        check(self.lib.clasfv_warp_fuse(self._h, prob.data_ptr(), c, motion.data_ptr(), _lib.torch_dtype_code(prob.dtype),
                                        _lib.i32_array(clip_starts), n, clip_len, num_frames, h, w,
                                        1 if edge_hops else 0, 1 if accumulate else 0, acc.data_ptr(), cnt.data_ptr(),
                                        mask.data_ptr() if mask is not None else None,
                                        area.data_ptr() if area is not None else None,
                                        _lib.current_stream_ptr(dev)), "clasfv_warp_fuse")
        return {"acc": acc, "cnt": cnt, "mask": mask, "area": area}

    def build_shift_clips(self, video, plan, clip_len=32):
        """video (3,T,H,W) fp32 CUDA; plan = list of (start, length, nclips). Returns (total,3,clip_len,H,W)."""
        _lib.require_cuda(video, "video")
        _, t, h, w = video.shape
        starts = [p[0] for p in plan]; lens = [p[1] for p in plan]; ncl = [p[2] for p in plan]
        bases = list(np.cumsum([0] + ncl[:-1]))
        total = int(sum(ncl))
        clips = torch.empty((total, 3, clip_len, h, w), dtype=torch.float32, device=video.device)
        check(self.lib.clasfv_build_shift_clips(self._h, video.data_ptr(), t, h, w, clip_len, len(plan),
                                                _lib.i32_array(starts), _lib.i32_array(lens), _lib.i32_array(ncl),
                                                _lib.i32_array(bases), clips.data_ptr(),
                                                _lib.current_stream_ptr(video.device)), "clasfv_build_shift_clips")
        return clips

    def fuse_shift_votes(self, prob, plan, num_frames, step=1, want_area=True):
        """prob (total,2,L,H,W); plan as in build_shift_clips. Returns (mask (T,H,W) uint8, area (T,) int32)."""
        _lib.require_cuda(prob, "prob")
        total, _, clip_len, h, w = prob.shape
        lens = [p[1] for p in plan]; ncl = [p[2] for p in plan]
        bases = list(np.cumsum([0] + ncl[:-1]))
        mask = torch.empty((num_frames, h, w), dtype=torch.uint8, device=prob.device)
        area = torch.empty((num_frames,), dtype=torch.int32, device=prob.device) if want_area else None
        check(self.lib.clasfv_fuse_shift_votes(self._h, prob.data_ptr(), _lib.torch_dtype_code(prob.dtype), num_frames, h, w,
                                               clip_len, step, len(plan), _lib.i32_array(lens), _lib.i32_array(ncl),
                                               _lib.i32_array(bases), mask.data_ptr(),
                                               area.data_ptr() if area is not None else None,
                                               _lib.current_stream_ptr(prob.device)), "clasfv_fuse_shift_votes")
        return mask, area

    def decoder_head(self, g, out_kind=OUT_LOGITS, out_dtype=torch.float32):
        """The fused tensor-core head alone (test surface): g = four fp16 CUDA maps (N,T,H/2^(l+1),W/2^(l+1),64)."""
        for x in g:
            _lib.require_cuda(x, "g")
            if x.dtype != torch.float16:
                raise ClasfvError("decoder_head: lateral maps must be float16")
        n, t, h2, w2, _ = g[0].shape
        h, w = 2 * h2, 2 * w2
        seg = torch.empty((n, 2, t, h, w), dtype=out_dtype, device=g[0].device)
        mot = torch.empty((n, 4, t, h, w), dtype=out_dtype, device=g[0].device)
        check(self.lib.clasfv_decoder_head(self._h, g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(), g[3].data_ptr(), n, t, h, w,
                                           out_kind, _lib.torch_dtype_code(out_dtype), seg.data_ptr(), mot.data_ptr(),
                                           _lib.current_stream_ptr(g[0].device)), "clasfv_decoder_head")
        return seg, mot

    # ------------------------------------------------------------------ single layer (tests)
    def conv3d(self, x, weight, scale=None, shift=None, stride=(1, 1, 1), padding=(0, 0, 0), residual=None, relu=False,
               engine="umma", out_f32=False):
        """x channels-last (N,T,H,W,Cin) CUDA (fp32|bf16); weight (Cout,Cin,kt,kh,kw) fp32 host tensor."""
        _lib.require_cuda(x, "x")
        n, t, h, w, cin = x.shape
        cout, _, kt, kh, kw = weight.shape
        to = (t + 2 * padding[0] - kt) // stride[0] + 1
        ho = (h + 2 * padding[1] - kh) // stride[1] + 1
        wo = (w + 2 * padding[2] - kw) // stride[2] + 1
        out_dtype = torch.float32 if (out_f32 or x.dtype == torch.float32) else x.dtype
        out = torch.empty((n, to, ho, wo, cout), dtype=out_dtype, device=x.device)
        wa = np.ascontiguousarray(weight.detach().cpu().float().numpy())
        sa = np.ascontiguousarray(scale.detach().cpu().float().numpy()) if scale is not None else None
        ba = np.ascontiguousarray(shift.detach().cpu().float().numpy()) if shift is not None else None
        check(self.lib.clasfv_conv3d(self._h, x.data_ptr(), _lib.torch_dtype_code(x.dtype), n, t, h, w, cin,
                                     wa.ctypes.data_as(C.c_void_p), sa.ctypes.data_as(C.c_void_p) if sa is not None else None,
                                     ba.ctypes.data_as(C.c_void_p) if ba is not None else None, cout, kt, kh, kw,
                                     stride[0], stride[1], stride[2], padding[0], padding[1], padding[2],
                                     residual.data_ptr() if residual is not None else None, 1 if relu else 0,
                                     1 if engine == "umma" else 0, 1 if out_f32 else 0, out.data_ptr(),
                                     _lib.current_stream_ptr(x.device)), "clasfv_conv3d")
        return out


def warp(src, flow, mode="bilinear"):
    """W1 + its grid_sample call site on fp32 CUDA tensors: src (N,C,H,W), flow (N,2,H,W); mode "bilinear" | "nearest"."""
    _lib.require_cuda(src, "src"); _lib.require_cuda(flow, "flow")
    if src.dtype != torch.float32 or flow.dtype != torch.float32:
        raise ClasfvError("warp: float32 tensors only")
    n, c, h, w = src.shape
    if tuple(flow.shape) != (n, 2, h, w):
        raise ClasfvError("warp: flow must be (N,2,H,W)")
    if mode not in ("bilinear", "nearest"):
        raise ClasfvError(f"warp: mode must be 'bilinear' or 'nearest', got {mode!r}")
    out = torch.empty_like(src)
    check(_lib.lib().clasfv_warp_mode(src.data_ptr(), flow.data_ptr(), out.data_ptr(), n, c, h, w, 1 if mode == "nearest" else 0,
                                      _lib.current_stream_ptr(src.device)), "clasfv_warp_mode")
    return out


def motion_field(offset, height, width):
    _lib.require_cuda(offset, "offset")
    n = offset.shape[0]
    grid = torch.empty((n, height, width, 2), dtype=torch.float32, device=offset.device)
    check(_lib.lib().clasfv_motion_field(offset.data_ptr(), grid.data_ptr(), n, height, width,
                                         _lib.current_stream_ptr(offset.device)), "clasfv_motion_field")
    return grid


def finalize_mask(acc, want_area=True):
    """acc (T,2,H,W) fp32 CUDA -> (mask (T,H,W) uint8, area (T,) int32): argmax over the class sums, ties -> 0."""
    _lib.require_cuda(acc, "acc")
    t, _, h, w = acc.shape
    mask = torch.empty((t, h, w), dtype=torch.uint8, device=acc.device)
    area = torch.empty((t,), dtype=torch.int32, device=acc.device) if want_area else None
    check(_lib.lib().clasfv_finalize_mask(acc.data_ptr(), t, h, w, mask.data_ptr(), area.data_ptr() if area is not None else None,
                                          _lib.current_stream_ptr(acc.device)), "clasfv_finalize_mask")
    return mask, area


def temporal_resample(x, out_len):
    """(C,L,H,W) fp32 CUDA -> (C,out_len,H,W), linear, align_corners=False."""
    _lib.require_cuda(x, "x")
    c, l, h, w = x.shape
    out = torch.empty((c, out_len, h, w), dtype=torch.float32, device=x.device)
    check(_lib.lib().clasfv_temporal_resample(x.data_ptr(), out.data_ptr(), c, l, out_len, h * w,
                                              _lib.current_stream_ptr(x.device)), "clasfv_temporal_resample")
    return out
