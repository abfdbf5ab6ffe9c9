"""ctypes binding of the C-ABI library ``csrc/libclasfv_b200.so`` (``include/clasfv_b200.h``).

The library is built in-tree by ``csrc/build.sh`` (``__graft_entry__.build()``).  There is no
fallback of any kind: if the shared object is missing, cannot be loaded, or no sm_100 device is
present, the operators raise ``ClasfvError``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

F32, BF16, F16 = 0, 1, 2
OUT_LOGITS, OUT_PROB, OUT_LVPROB = 0, 1, 2

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(CSRC, "libclasfv_b200.so")

# every symbol include/clasfv_b200.h declares
EXPORTS = (
    "clasfv_abi_version", "clasfv_last_error", "clasfv_create", "clasfv_destroy", "clasfv_set_tensor",
    "clasfv_finalize", "clasfv_forward", "clasfv_workspace_bytes", "clasfv_warp", "clasfv_motion_field",
    "clasfv_warp_fuse", "clasfv_build_shift_clips", "clasfv_fuse_shift_votes", "clasfv_temporal_resample",
    "clasfv_conv3d", "clasfv_profile_begin", "clasfv_profile_end", "clasfv_finalize_mask",
    "clasfv_set_option", "clasfv_profile_gflop", "clasfv_warp_mode", "clasfv_ingest_u8", "clasfv_decoder_head", "clasfv_launch_count",
)


class ClasfvError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a (nvcc cross-compiles; no GPU needed)."""
    res = subprocess.run(["sh", os.path.join(CSRC, "build.sh")], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout, res.stderr)
    if res.returncode != 0:
        raise ClasfvError("building libclasfv_b200.so failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def lib():
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise ClasfvError(f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(csrc/build.sh). There is no CPU or PyTorch fallback.")
        try:
            l = C.CDLL(LIB_PATH)
        except OSError as e:
            raise ClasfvError(f"cannot load {LIB_PATH}: {e}") from e
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        l.clasfv_abi_version.restype = i32
        l.clasfv_last_error.restype = C.c_char_p
        l.clasfv_create.argtypes = [i32, C.POINTER(vp)]
        l.clasfv_destroy.argtypes = [vp]
        l.clasfv_destroy.restype = None
        l.clasfv_set_tensor.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32]
        l.clasfv_finalize.argtypes = [vp, i32]
        l.clasfv_forward.argtypes = [vp, vp, C.POINTER(i64), i64, i32, i32, i32, i32, i32, i32, vp, vp, vp]
        l.clasfv_workspace_bytes.argtypes = [vp]
        l.clasfv_workspace_bytes.restype = i64
        l.clasfv_warp.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
        l.clasfv_ingest_u8.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, i32, vp]
        l.clasfv_warp_mode.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp]
        l.clasfv_motion_field.argtypes = [vp, vp, i32, i32, i32, vp]
        l.clasfv_warp_fuse.argtypes = [vp, vp, i32, vp, i32, C.POINTER(C.c_int32), i32, i32, i32, i32, i32, i32, i32,
                                       vp, vp, vp, vp, vp]
        i32p = C.POINTER(C.c_int32)
        l.clasfv_build_shift_clips.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32p, i32p, i32p, i32p, vp, vp]
        l.clasfv_fuse_shift_votes.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32p, i32p, i32p, vp, vp, vp]
        l.clasfv_finalize_mask.argtypes = [vp, i32, i32, i32, vp, vp, vp]
        l.clasfv_profile_begin.argtypes = [vp]
        l.clasfv_profile_end.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
        l.clasfv_profile_gflop.argtypes = [vp, C.POINTER(C.c_double)]
        l.clasfv_set_option.argtypes = [vp, C.c_char_p, i32]
        l.clasfv_temporal_resample.argtypes = [vp, vp, i32, i32, i32, i64, vp]
        l.clasfv_conv3d.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32,
                                    i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp]
        l.clasfv_launch_count.argtypes = [vp]
        l.clasfv_launch_count.restype = i64
        l.clasfv_decoder_head.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]
        for name in EXPORTS:
            getattr(l, name)                  # every declared symbol must resolve
        _lib = l
        return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().clasfv_last_error().decode("utf-8", "replace")
        raise ClasfvError(f"{what or 'libclasfv_b200'} failed (code {rc}): {msg}")


def i32_array(values):
    arr = (C.c_int32 * len(values))(*[int(v) for v in values])
    return arr


def i64_array(values):
    arr = (C.c_int64 * len(values))(*[int(v) for v in values])
    return arr


def torch_dtype_code(dtype):
    import torch
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    if dtype == torch.float16:
        return F16
    raise ClasfvError(f"unsupported element type {dtype}; use torch.float32, torch.bfloat16 or torch.float16")


def current_stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    if not t.is_cuda:
        raise ClasfvError(f"{name} must be a CUDA tensor: clasfv_b200 has no CPU path (got device {t.device})")
    if not t.is_contiguous():
        raise ClasfvError(f"{name} must be contiguous")
    return t
