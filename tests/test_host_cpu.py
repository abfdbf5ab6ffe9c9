"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the host logic of
the drop-ins (planning, error behaviour) works without a GPU, and the product never falls back."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import clasfv_b200
from clasfv_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "clasfv_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(clasfv_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTS)
    lib = _lib.lib()
    assert os.path.dirname(_lib.LIB_PATH).startswith(clasfv_b200.PACKAGE_DIR)      # built in-tree
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.clasfv_abi_version() == 2
    assert lib.clasfv_last_error() is not None


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_cpu_fallback_anywhere():
    from clasfv_b200.src import fuse_utils
    from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
    lib = _lib.lib()
    h = C.c_void_p()
    assert lib.clasfv_create(0, C.byref(h)) == 5                                   # CLASFV_EUNSUPPORTED
    assert b"no CPU path" in lib.clasfv_last_error()
    net = R2plus1D_18_MotionNet(pretrained=False).eval()
    with pytest.raises(_lib.ClasfvError):
        net(torch.zeros(1, 3, 8, 16, 16))
    with pytest.raises(_lib.ClasfvError):
        fuse_utils.divide_to_consecutive_clips(np.zeros((3, 64, 16, 16), np.float32))
    with pytest.raises(_lib.ClasfvError):
        fuse_utils.segment_a_video_with_fusion(np.zeros((3, 64, 16, 16), np.float32), net)


def test_state_dict_is_the_reference_state_dict():
    from clasfv_b200 import synthetic
    from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
    net = R2plus1D_18_MotionNet(pretrained=False)
    sd = net.state_dict()
    spec = synthetic.state_dict_spec()
    assert [k for k, _s, _k in spec] == list(sd.keys()) and len(sd) == 242
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == 31575731
    net.load_state_dict(synthetic.random_state_dict(1))                            # same keys, same shapes
    wrapped = torch.nn.DataParallel(net)
    assert all(k.startswith("module.") for k in wrapped.state_dict())
    assert float(net.motion_head.weight.std()) > 0
    with pytest.raises(_lib.ClasfvError):                                          # training mode is not implemented
        net.train()(torch.zeros(1, 3, 8, 16, 16))


def test_shift_planning_matches_oracle(capsys):
    from clasfv_b200.src import fuse_utils
    from oracle import fuse_ref
    for t, step, f in ((200, 1, 32), (128, 1, 1), (40, 1, 10), (20, 1, 10), (175, 2, 6), (64, 3, 4)):
        assert fuse_utils.plan_shifts(t, step, f) == fuse_ref.plan_shifts(t, step, f)
    fuse_utils.plan_shifts(20, 1, 10)
    assert "Video is too short" in capsys.readouterr().out
    for length in (48, 80, 112, 175, 176, 200):
        assert fuse_utils._num_clips(length) == fuse_ref.num_consecutive_clips(length)
    assert fuse_utils._shift_plan_entry(3, 75, 32, True) == (3, 75, 2)
    assert fuse_utils._shift_plan_entry(0, 70, 32, False) == (0, 64, 2)           # truncation without resample
    with pytest.raises(ValueError):
        fuse_utils._shift_plan_entry(0, 60, 32, False)                            # rounds up: the reference fails in concatenate


def test_host_helpers_match_golden(golden_dir):
    from clasfv_b200.src.echonet_dataset import EDESpairs, zeroone_normalizer
    g = np.load(os.path.join(golden_dir, "host_helpers.npz"))
    np.testing.assert_array_equal(zeroone_normalizer(g["v"].copy()), g["norm"])
    pairs = EDESpairs([0, 31, 62, 95], [14, 47, 49, 80, 120])
    np.testing.assert_array_equal(np.array(pairs), g["pairs"])


def test_ef_from_synthetic_beating_masks():
    from clasfv_b200.src.fuse_utils import compute_ef_using_putative_clips, get2dPucks
    t, h, w = 120, 112, 112
    yy, xx = np.mgrid[0:h, 0:w]
    masks = np.zeros((t, h, w), np.int64)
    for i in range(t):
        s = 1.0 - 0.25 * (0.5 - 0.5 * np.cos(2 * np.pi * i / 40.0))               # 3 beats
        masks[i] = (((yy - 60) / (30 * s)) ** 2 + ((xx - 56) / (18 * s)) ** 2) <= 1
    efs, pairs = compute_ef_using_putative_clips(masks, "synthetic", return_edes=True)
    assert len(pairs) >= 2 and all(es > ed for ed, es in pairs)
    # volume ~ s^3: EF = 1 - 0.75^3 = 57.8 %
    assert all(45 < ef < 70 for ef in efs), efs
    length, radii = get2dPucks(masks[0].astype(int), (1.0, 1.0))
    assert 55 < length < 65 and len(radii) == 10


def test_forward_windows_groups_equally_spaced_runs_into_single_calls():
    """Engine.forward_windows hands every maximal run of equally spaced windows to ONE clasfv_forward call (that is
    what lets the library share the stem and layer1 between overlapping windows); host logic only, no GPU."""
    import torch
    from clasfv_b200.engine import Engine

    class Recorder(Engine):
        def __init__(self):
            self.calls, self.options, self.precision = [], {}, 1

        def set_option(self, name, value):
            self.options[name] = value

        def forward_into(self, x, seg, mot, out_kind, clip_starts=None, clip_len=None):
            assert seg.shape[0] == len(clip_starts) == mot.shape[0]
            self.calls.append(list(clip_starts))

        def __del__(self):
            pass

    video = torch.zeros(3, 64, 4, 4)
    for starts, expect in [
        (list(range(0, 33)), [list(range(0, 33))]),                               # stride 1: one call
        ([0, 2, 4, 6, 7], [[0, 2, 4, 6], [7]]),                                   # the appended tail window is off the grid
        ([5], [[5]]),
        ([0, 3, 6, 10, 14, 18], [[0, 3, 6], [10, 14, 18]]),
    ]:
        eng = Recorder()
        seg = torch.zeros(len(starts), 2, 32, 4, 4); mot = torch.zeros(len(starts), 4, 32, 4, 4)
        eng.forward_windows(video, seg, mot, 1, starts, 32, batch_clips=7)
        assert eng.calls == expect and eng.options == {"sub_batch": 7}


def test_config4_lengths_and_video_assignment():
    """BASELINE config 4: 1 277 lengths ~N(175,55) in [64,400]; every video lands on exactly one rank and
    the longest-first assignment keeps the per-rank clip counts within 0.1 % of each other (round-robin: several %)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "config4_many_videos", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "config4_many_videos.py"))
    c4 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(c4)
    lengths = c4.config4_lengths()
    assert len(lengths) == 1277 and lengths.min() >= 64 and lengths.max() <= 400
    assert abs(float(lengths.mean()) - 175.0) < 5.0
    assert (lengths == c4.config4_lengths()).all()                      # seeded
    for world in (1, 2, 4, 8):
        for how in ("lpt", "round_robin"):
            seen = sorted(i for r in range(world) for i in c4.assign(lengths, r, world, how))
            assert seen == list(range(1277))
        loads = c4.rank_loads(lengths, world, "lpt")
        assert max(loads) / (sum(loads) / world) < 1.001


def test_clamped_corner_taps_equal_grid_sample_conditional_taps_bit_for_bit():
    """The staged warp-fuse kernel clamps the north-west bilinear corner to (H-2, W-2) and reads all four taps
    unconditionally (csrc/fusion.cu taps_setup / tap_sum); grid_sample (and warp_fuse_kernel) floor the clamped coordinate and
    skip out-of-range corners.  Emulated here in numpy with the kernels' operation order (fp32, one rounding per fma): the
    two must give the same BITS for interior points, exact integers and every clamped border."""
    import numpy as np
    f32 = np.float32
    rng = np.random.default_rng(0)
    h, w, n = 48, 64, 300_000
    ix = rng.uniform(-3, w + 2, n).astype(f32)
    iy = rng.uniform(-3, h + 2, n).astype(f32)
    ix[:2000] = rng.integers(0, w, 2000).astype(f32); iy[:2000] = rng.integers(0, h, 2000).astype(f32)
    ix[2000:3000] = f32(w - 1); iy[3000:4000] = f32(h - 1)
    ix[4000:4500] = np.nextafter(f32(w - 1), f32(0)); iy[4500:5000] = np.nextafter(f32(h - 1), f32(0))
    ix = np.minimum(f32(w - 1), np.maximum(ix, f32(0))); iy = np.minimum(f32(h - 1), np.maximum(iy, f32(0)))
    plane = rng.uniform(0, 1, (h, w)).astype(f32)

    def fma(a, b, c):          # fp32 fma: the product of two fp32 numbers is exact in fp64
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)

    # grid_sample / warp_fuse_kernel: conditional taps (a skipped tap == the clamped neighbour with weight 0)
    fx0, fy0 = np.floor(ix), np.floor(iy)
    x0, y0 = fx0.astype(int), fy0.astype(int)
    x1, y1 = fx0 + f32(1), fy0 + f32(1)
    nw, ne, sw, se = (x1 - ix) * (y1 - iy), (ix - fx0) * (y1 - iy), (x1 - ix) * (iy - fy0), (ix - fx0) * (iy - fy0)
    x1ok, y1ok = x0 + 1 < w, y0 + 1 < h
    xs, ys = np.minimum(x0 + 1, w - 1), np.minimum(y0 + 1, h - 1)
    v = plane[y0, x0] * nw
    v = np.where(x1ok, fma(plane[y0, xs], ne, v), v)
    v = np.where(y1ok, fma(plane[ys, x0], sw, v), v)
    ref = np.where(x1ok & y1ok, fma(plane[ys, xs], se, v), v)

    # staged kernel: clamped corner, 1 - (ix - x0) weights, unconditional taps
    gx0, gy0 = np.minimum(np.floor(ix), f32(w - 2)), np.minimum(np.floor(iy), f32(h - 2))
    wx1, wy1 = ix - gx0, iy - gy0
    wx0, wy0 = f32(1) - wx1, f32(1) - wy1
    X, Y = gx0.astype(int), gy0.astype(int)
    v = plane[Y, X] * (wx0 * wy0)
    v = fma(plane[Y, X + 1], wx1 * wy0, v)
    v = fma(plane[Y + 1, X], wx0 * wy1, v)
    new = fma(plane[Y + 1, X + 1], wx1 * wy1, v)
    assert np.array_equal(ref.view(np.uint32), new.view(np.uint32))


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside the GPU arm): stdout must be the one JSON line of the
    contract even when libraries print to file descriptor 1, with the keys the driver reads."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[1]")
