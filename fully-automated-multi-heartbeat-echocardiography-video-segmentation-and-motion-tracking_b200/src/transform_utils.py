"""Drop-in for ``generate_2dmotion_field`` of the reference's ``src/transform_utils.py:14-34`` (the rest of
that file is CAMUS training-time augmentation, out of scope) plus the warp it feeds."""
from __future__ import annotations

from .. import engine as _engine


def generate_2dmotion_field(x, offset):
    """Sampling grid (N,H,W,2) for ``F.grid_sample(x, grid, align_corners=False, padding_mode='border')``:
    base ``linspace(-1, 1, S)`` mesh + the normalised displacement (channel 0 -> x, channel 1 -> y)."""
    return _engine.motion_field(offset.contiguous().float(), int(x.shape[2]), int(x.shape[3]))


def warp(x, offset, mode="bilinear"):
    """``F.grid_sample(x, generate_2dmotion_field(x, offset), align_corners=False, mode=mode,
    padding_mode='border')`` (clasfv_losses.py:86-87, visualization_utils.py:123-126) in one kernel."""
    return _engine.warp(x.contiguous().float(), offset.contiguous().float(), mode)
