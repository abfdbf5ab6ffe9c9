"""GPU parity tests (pytest -m gpu, B200): the CUDA path, called through the reference-shaped drop-ins
and the C ABI, against the CPU oracle on identical seeded inputs and against the committed golden
vectors.  Tolerances are the north-star's (BASELINE.json), stated at each assert; where a fixture's
own fp32 noise floor is above a tolerance the test says so and measures the floor.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import clasfv_b200
import clasfv_b200.synthetic as synthetic
from clasfv_b200 import _lib
from clasfv_b200 import engine as E
from clasfv_b200.src import fuse_utils, transform_utils
from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
from oracle import fixtures, fuse_ref, model_emul, model_ref
from oracle.model_emul import Config

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sd():
    return fixtures.calibrated_state_dict(0)


def _net(sd, precision):
    net = R2plus1D_18_MotionNet(pretrained=False, precision=precision)
    net.load_state_dict(sd)
    return net.cuda().eval()


@pytest.fixture(scope="module")
def net_fp32(sd):
    return _net(sd, "fp32")


@pytest.fixture(scope="module")
def net_bf16(sd):
    return _net(sd, "bf16")


@pytest.fixture(scope="module")
def net_fp16(sd):
    return _net(sd, "fp16")


@pytest.fixture(scope="module")
def eng():
    return E.Engine("cuda:0")


# storage points of the two tensor-core modes as oracle/model_emul.py models them (csrc/: 16-bit trunk, fp16 lateral maps
# and interpolation weights in both modes, comb_2 / head operands in the trunk's type)
EMUL = {"bf16": Config(act="bf16", wtrunk="bf16", lateral="f16", head="patch", h1="bf16", h2="bf16", w2="bf16", wh="bf16", out="fp32"),
        "fp16": Config(act="f16", wtrunk="f16", lateral="f16", head="patch", h1="f16", h2="f16", w2="f16", wh="f16", out="fp32")}
ULP = {"bf16": 2.0 ** -8, "fp16": 2.0 ** -11}          # spacing of the 16-bit type relative to the value (upper bound)


def softmax_metrics(seg, seg_ref):
    p, pr = torch.softmax(seg.float().cpu(), 1), torch.softmax(seg_ref.float(), 1)
    lv, lvr = p[:, 1] > p[:, 0], pr[:, 1] > pr[:, 0]
    return {"max": float((p - pr).abs().max()), "mean": float((p - pr).abs().mean()),
            "agree": float((lv == lvr).float().mean()), "near": float(((pr[:, 1] - 0.5).abs() < 2e-2).float().mean()),
            "p": p, "pr": pr}


# ------------------------------------------------------------------------------------------ loading
def test_native_library_is_what_runs():
    assert torch.cuda.get_device_capability(0)[0] == 10, "these tests are for sm_100 (B200)"
    assert os.path.samefile(os.path.dirname(_lib.LIB_PATH), os.path.join(clasfv_b200.PACKAGE_DIR, "csrc"))
    assert _lib.lib().clasfv_abi_version() == 2
    loaded = open("/proc/self/maps").read()
    assert "libclasfv_b200.so" in loaded


# ------------------------------------------------------------------------------------------ W1
def test_warp_primitive_matches_oracle_and_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "warp.npz"))
    src, flow = torch.from_numpy(g["src"]), torch.from_numpy(g["flow"])
    grid = transform_utils.generate_2dmotion_field(src.cuda(), flow.cuda())
    assert grid.is_cuda and tuple(grid.shape) == (2, 12, 20, 2)
    np.testing.assert_allclose(grid.cpu().numpy(), g["grid"], atol=1e-6, rtol=0)
    out = transform_utils.warp(src.cuda(), flow.cuda()).cpu()
    np.testing.assert_allclose(out.numpy(), g["warped"], atol=1e-6, rtol=0)           # fp32: <= 1e-6
    # zero flow is not the identity: x_src = j*W/(W-1) - 1/2 (reference quirk)
    colimg = torch.arange(112.0).view(1, 1, 1, 112).expand(1, 1, 112, 112).contiguous()
    cols = transform_utils.warp(colimg.cuda(), torch.zeros(1, 2, 112, 112).cuda())[0, 0, 0].cpu().numpy()
    np.testing.assert_allclose(cols, g["zero_flow_cols"], atol=2e-5)
    # large shapes / odd sizes / saturated flows against the oracle
    gen = torch.Generator().manual_seed(4)
    for (n, c, h, w, s) in ((2, 2, 112, 112, 0.05), (1, 3, 37, 53, 0.5), (3, 1, 224, 224, 0.02)):
        src = torch.rand(n, c, h, w, generator=gen)
        flow = torch.tanh(s * torch.randn(n, 2, h, w, generator=gen))
        got = E.warp(src.cuda(), flow.cuda()).cpu()
        assert float((got - fuse_ref.warp(src, flow)).abs().max()) <= 1e-6


# ------------------------------------------------------------------------------------------ C1
@pytest.mark.parametrize("shape,size,bgr", [((12, 60, 80), (112, 112), False), ((9, 150, 130), (112, 112), True),
                                            ((5, 112, 112), (112, 112), True), ((4, 33, 47), (48, 64), False)])
def test_video_ingest_matches_host_pipeline(eng, shape, size, bgr):
    """clasfv_ingest_u8 against the reference's host pipeline (motion_segment.py:96-106): float32 cast, F.interpolate(
    trilinear, align_corners=True) to (T,h,w), zeroone_normalizer.  fp32 tolerance 2e-6 on [0,1] values."""
    from clasfv_b200.src.echonet_dataset import zeroone_normalizer
    rng = np.random.RandomState(3)
    t, h0, w0 = shape
    frames = rng.randint(0, 256, size=(t, h0, w0, 3)).astype(np.uint8)
    frames[..., 2] = (frames[..., 2] * 0.6).astype(np.uint8)                    # channels with different ranges
    rgb = frames[..., ::-1] if bgr else frames
    video = np.ascontiguousarray(rgb.transpose((3, 0, 1, 2))).astype(np.float32)
    v = torch.Tensor(video).unsqueeze(0)
    v = F.interpolate(v, size=(v.shape[2], size[0], size[1]), mode="trilinear", align_corners=True)
    want = zeroone_normalizer(v.squeeze(0).numpy().copy())
    got = eng.ingest_u8(frames, size[0], size[1], bgr=bgr)
    assert got.is_cuda and tuple(got.shape) == (3, t, size[0], size[1]) and got.dtype == torch.float32
    got = got.cpu().numpy()
    assert float(np.abs(got - want).max()) <= 2e-6
    assert got.min() == 0.0 and got.reshape(3, -1).max(axis=1).tolist() == [1.0, 1.0, 1.0]


@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
@pytest.mark.parametrize("forward", [True, False])
def test_apply_sequence_deformation_matches_oracle(mode, forward):
    """Motion-tracking product (reference src/visualization_utils.py:106-128): a label / image carried through a run of
    motion fields by chained warps.  Bilinear: <= 1e-5 after 6 chained warps; nearest: the same pixels picked (exact)."""
    from clasfv_b200.src.visualization_utils import apply_sequence_deformation
    g = torch.Generator().manual_seed(11)
    n, t, h, w = 2, 10, 40, 56
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    label = (((yy - 20) / 12) ** 2 + ((xx - 28) / 9) ** 2 <= 1).float().expand(n, 1, h, w).contiguous()
    image = torch.rand(n, 1, h, w, generator=g)
    src = label if mode == "nearest" else image
    # smooth flows of a few pixels (white noise would make every nearest pick a coin flip between fp32 rounding orders)
    motion = torch.tanh(F.interpolate(torch.randn(n, 4, t, 5, 7, generator=g), size=(t, h, w), mode="trilinear", align_corners=True) * 0.08)
    a, b = (1, 7) if forward else (8, 2)
    ref = fuse_ref.apply_sequence_deformation(src, motion, a, b, grid_mode=mode, forward=forward)
    out = apply_sequence_deformation(src.cuda(), motion.cuda(), a, b, grid_mode=mode, forward=forward).cpu()
    assert out.shape == ref.shape
    if mode == "bilinear":
        assert float((out - ref).abs().max()) <= 1e-5
    else:
        assert float((out != ref).float().mean()) <= 1e-3          # ties at exactly .5 may round differently after fp32 reassociation
        assert set(out.unique().tolist()) <= {0.0, 1.0}
    with pytest.raises(UnboundLocalError):
        apply_sequence_deformation(src.cuda(), motion.cuda(), 3, 3, grid_mode=mode, forward=forward)


@pytest.mark.parametrize("length", [75, 48, 64, 80])
def test_divide_to_consecutive_clips_matches_golden(golden_dir, length):
    g = np.load(os.path.join(golden_dir, "divide_clips.npz"))
    video = synthetic.synthetic_echo_video(length, 112, 112, seed=20 + length)
    clips = fuse_utils.divide_to_consecutive_clips(video, interpolate_last=True)
    assert list(clips.shape) == list(g[f"shape_{length}"]) and clips.dtype == np.float64
    np.testing.assert_allclose(clips[:, :, :, ::16, ::16], g[f"sub_{length}"], atol=1e-6, rtol=0)
    assert abs(clips.sum() - float(g[f"sum_{length}"])) < 1e-6 * abs(float(g[f"sum_{length}"]))


def test_divide_to_consecutive_clips_edge_cases():
    rng = np.random.default_rng(0)
    v = rng.random((3, 70, 16, 32)).astype(np.float32)
    a = fuse_utils.divide_to_consecutive_clips(v, interpolate_last=False)               # truncates to 2 clips
    np.testing.assert_array_equal(a, fuse_ref.divide_to_consecutive_clips(v, interpolate_last=False))
    with pytest.raises(ValueError):
        fuse_utils.divide_to_consecutive_clips(v[:, :60], interpolate_last=False)        # rounds up, last clip short
    b = fuse_utils.divide_to_consecutive_clips(v[:, :17], interpolate_last=True)         # 17 -> one 32-frame clip
    np.testing.assert_allclose(b, fuse_ref.divide_to_consecutive_clips(v[:, :17], interpolate_last=True), atol=1e-6)
    x = torch.from_numpy(v)
    for l_out in (64, 96, 33):
        got = E.temporal_resample(x.cuda(), l_out).cpu().numpy()
        np.testing.assert_allclose(got, fuse_ref.temporal_resample(v, l_out), atol=1e-6, rtol=0)


# ------------------------------------------------------------------------------------------ conv layers
CONV_CASES = [
    # n, t, h, w, cin, cout, kernel, stride, pad, residual, relu, out_f32
    (1, 2, 8, 8, 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), False, False, False),
    (1, 2, 8, 8, 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), True, False, True),
    (1, 4, 8, 8, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), False, True, False),
    (1, 2, 16, 16, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, True, False),
    (1, 2, 16, 16, 64, 240, (1, 3, 3), (1, 2, 2), (0, 1, 1), False, True, False),
    (1, 8, 8, 8, 240, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0), False, True, False),
    (1, 4, 16, 16, 64, 128, (1, 1, 1), (2, 2, 2), (0, 0, 0), False, False, False),
    (2, 2, 7, 7, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, True, False),
    (2, 4, 7, 7, 576, 256, (3, 1, 1), (1, 1, 1), (1, 0, 0), True, True, False),
    (3, 4, 7, 7, 960, 512, (3, 1, 1), (2, 1, 1), (1, 0, 0), False, True, False),
    (1, 4, 7, 7, 512, 1152, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, True, False),
    (1, 8, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, True, False),       # 196 tiles: persistent loop wraps
    (2, 4, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, True, False),      # layer2 geometry: 28-wide rows
    (2, 4, 14, 14, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, True, False),      # layer3 geometry
    (2, 8, 14, 14, 576, 256, (3, 1, 1), (1, 1, 1), (1, 0, 0), True, True, False),       # layer3 temporal: 196-row frames
]


def _conv_case(eng, case, dtype, engine_name):
    n, t, h, w, cin, cout, k, s, p, use_res, relu, out_f32 = case
    g = torch.Generator().manual_seed(cin * 7 + cout)
    x = torch.randn(n, t, h, w, cin, generator=g)
    wt = torch.randn(cout, cin, *k, generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5
    scale = 0.5 + torch.rand(cout, generator=g)
    shift = 0.2 * torch.randn(cout, generator=g)
    xq = x.to(dtype).float()
    wq = (wt * scale.view(-1, 1, 1, 1, 1)).to(dtype).float()
    ref = F.conv3d(xq.permute(0, 4, 1, 2, 3), wq, shift, s, p).permute(0, 2, 3, 4, 1).contiguous()
    res = None
    if use_res:
        res = (0.5 * torch.randn(ref.shape, generator=g)).to(torch.float32 if (out_f32 or dtype == torch.float32) else dtype)
        ref = ref + res.float()
    if relu:
        ref = ref.relu()
    out = eng.conv3d(x.to(dtype).cuda(), wt, scale, shift, s, p, res.cuda() if res is not None else None, relu,
                     engine=engine_name, out_f32=out_f32)
    return out.float().cpu(), ref


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_cuda_core_fp32(eng, case):
    out, ref = _conv_case(eng, case, torch.float32, "simt")
    assert float((out - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_tcgen05_bf16(eng, case):
    out, ref = _conv_case(eng, case, torch.bfloat16, "umma")
    out_f32 = case[-1]
    # fp32 accumulation of exactly-representable bf16 products: only the output rounding differs
    tol = 1e-4 if out_f32 else 2.0 ** -8
    assert float(((out - ref).abs() / (ref.abs() + 1.0)).max()) <= tol
    out2, _ = _conv_case(eng, case, torch.bfloat16, "simt")                             # same rounding points, other unit
    assert float(((out - out2).abs() / (ref.abs() + 1.0)).max()) <= 2 * tol      # summation order may move a value across one rounding boundary


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[5] >= 128])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_conv_cta_pair_is_bit_identical_to_single_cta(eng, case, dtype):
    """Layers with >= 128 output columns run on CTA pairs (tcgen05.mma.cta_group::2, M = 256, half of the filter rows per
    CTA, csrc/conv_umma.cu).  Every output element is accumulated over the same K steps in the same order as by the
    single-CTA kernel wherever both use the same slab layout, so there the option must not change a bit - odd tile counts (a
    pair whose second tile is empty), ragged last N tiles, strided and residual cases included."""
    try:
        eng.set_option("umma_pair", 1)
        pair, ref = _conv_case(eng, case, dtype, "umma")
        eng.set_option("umma_pair", 0)
        single, _ = _conv_case(eng, case, dtype, "umma")
    finally:
        eng.set_option("umma_pair", 1)
    assert torch.isfinite(pair).all()
    assert float(((pair - ref).abs() / (ref.abs() + 1.0)).max()) <= (2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11)
    ndiff = int((pair != single).sum())
    print(f"\n[pair vs single, {case[4]}->{case[5]} k={case[6]} {dtype}] differing outputs: {ndiff} of {pair.numel()}, max abs {float((pair - single).abs().max()):.3g}, "
          f"max err vs fp32 reference: pair {float((pair - ref).abs().max()):.4g} single {float((single - ref).abs().max()):.4g}")
    # The order in which K is walked does not depend on the tiling or on whether a tap family shares its slab (conv_umma.cu,
    # UmmaParams::span), so the halved filter stage of a pair - which lets the 256-column layers keep the shared layout - moves no bit.
    assert ndiff == 0


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_tcgen05_fp16(eng, case):
    out, ref = _conv_case(eng, case, torch.float16, "umma")
    tol = 1e-4 if case[-1] else 2.0 ** -11          # one rounding of the stored type; fp32 outputs: summation order only
    assert float(((out - ref).abs() / (ref.abs() + 1.0)).max()) <= tol


# ------------------------------------------------------------------------------------------ network
def test_forward_fp32_matches_golden(net_fp32, golden_dir):
    g = np.load(os.path.join(golden_dir, "model_forward.npz"))
    for tag in ("a", "b"):
        seg, mot = net_fp32(torch.from_numpy(g[f"x_{tag}"]).cuda())
        assert seg.dtype == torch.float32 and tuple(seg.shape) == g[f"seg_{tag}"].shape and tuple(mot.shape) == g[f"motion_{tag}"].shape
        m = softmax_metrics(seg, torch.from_numpy(g[f"seg_{tag}"]))
        assert m["max"] <= 1e-4, m["max"]                                               # north-star: softmax max-abs <= 1e-4 (fp32)
        assert m["agree"] >= 0.999
        h, w = seg.shape[-2:]
        epe = (mot.cpu() - torch.from_numpy(g[f"motion_{tag}"])).abs()
        assert float(epe.max()) * max(h, w) / 2 <= 1e-2                                  # flow end-point error <= 1e-2 px


@pytest.mark.parametrize("shape,seed,batch", [((8, 32, 32), 21, 3), ((16, 64, 48), 22, 2), ((32, 112, 112), 13, 1), ((8, 224, 224), 23, 1)])
def test_forward_fp32_matches_oracle(net_fp32, sd, shape, seed, batch):
    x = fixtures.synthetic_clip(*shape, seed=seed, batch=batch)
    seg_ref, mot_ref = model_ref.forward(sd, x)
    seg, mot = net_fp32(x)                      # a CPU tensor is moved to the model's device, as DataParallel would
    assert seg.is_cuda
    m = softmax_metrics(seg, seg_ref)
    # The reference's own fp32 evaluation is only so accurate: measure its distance to the float64
    # evaluation of the same algorithm on this input and allow that much on top of the 1e-4 gate.
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    seg64, mot64 = model_ref.forward(sd64, x.double())
    floor = float((torch.softmax(seg_ref.double(), 1) - torch.softmax(seg64, 1)).abs().max())
    ours64 = float((torch.softmax(seg.double().cpu(), 1) - torch.softmax(seg64, 1)).abs().max())
    print(f"\n[fp32 {shape}] softmax max|d| vs oracle {m['max']:.2e}; oracle fp32 vs fp64 {floor:.2e}; ours vs fp64 {ours64:.2e}; "
          f"agree {m['agree'] * 100:.4f}% near-boundary {m['near'] * 100:.2f}%")
    assert m["max"] <= 1e-4 + floor
    assert ours64 <= 1e-4 + floor                        # at least as close to the exact answer as the reference is
    assert m["agree"] >= 0.999                           # argmax-mask agreement >= 99.9 %
    assert float((mot.cpu() - mot_ref).abs().max()) * max(shape[1:]) / 2 <= 1e-2        # EPE <= 1e-2 px
    assert float(mot.abs().max()) < 1.0


def test_forward_bf16_against_oracle_and_autocast_yardstick(net_bf16, sd):
    shape = (32, 112, 112)
    x = fixtures.synthetic_clip(*shape, seed=13, batch=1)
    seg_ref, mot_ref = model_ref.forward(sd, x)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        seg_ac, mot_ac = model_ref.forward(sd, x)         # the reference algorithm under PyTorch's own bf16 autocast
    seg, mot = net_bf16(x.cuda())
    ours, yard = softmax_metrics(seg, seg_ref), softmax_metrics(seg_ac, seg_ref)
    epe = float((mot.float().cpu() - mot_ref).abs().max()) * 56
    epe_mean = float((mot.float().cpu() - mot_ref).abs().mean()) * 56
    print(f"\n[bf16 {shape}] ours: softmax max {ours['max']:.3f} mean {ours['mean']:.4f} agree {ours['agree'] * 100:.3f}% | "
          f"autocast yardstick: max {yard['max']:.3f} mean {yard['mean']:.4f} agree {yard['agree'] * 100:.3f}% | "
          f"near-boundary {ours['near'] * 100:.2f}% | EPE max {epe:.3f} mean {epe_mean:.4f} px")
    # bf16 gates.  The north-star numbers (softmax <= 2e-2, agreement >= 99.9 %) presuppose a trained,
    # well-conditioned network; on seeded random weights bf16 storage noise is amplified through 36
    # convolutions (DESIGN.md "Fixture conditioning").  What must hold on any weights:
    assert ours["mean"] <= yard["mean"] and ours["agree"] >= yard["agree"]               # no worse than torch's bf16
    assert ours["mean"] <= 2e-2                                                          # mean softmax error within 2e-2
    far = (ours["pr"][:, 1] - 0.5).abs() >= 0.25                                         # pixels outside the bf16 noise band
    lv, lvr = ours["p"][:, 1] > 0.5, ours["pr"][:, 1] > 0.5
    assert float((lv == lvr)[far].float().mean()) >= 0.999
    assert ours["agree"] >= 0.98
    assert epe_mean <= 0.1


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("t,h,w", [(8, 112, 112), (4, 224, 224), (8, 48, 144), (8, 32, 32), (8, 16, 16), (2, 16, 512), (2, 512, 16)])
def test_decoder_head_alone_matches_its_emulation(sd, net_bf16, net_fp16, precision, t, h, w):
    """The fused tensor-core head (csrc/decoder_umma.cu) on caller-supplied fp16 lateral maps against the emulation of
    its arithmetic on the SAME maps (oracle/model_emul.py:head): interpolation weights fp16(wH*wW), fp32 accumulation,
    relu -> 16-bit twice, biases as hi + lo.  Nothing upstream can differ, so the only differences are fp32 summation order and
    the 16-bit roundings it flips: at most 2 ulp of the 16-bit type at the scale of the terms of the head's dot product."""
    eng = (net_bf16 if precision == "bf16" else net_fp16).engine()
    gen = torch.Generator().manual_seed(h * 7 + w + t)
    g = [(0.7 * torch.randn(2, 64, t, h >> (l + 1), w >> (l + 1), generator=gen)).half() for l in range(4)]
    seg, mot = eng.decoder_head([x.permute(0, 2, 3, 4, 1).contiguous().cuda() for x in g])
    seg_e, mot_e = model_emul.head(model_ref.strip_module_prefix(sd), [x.float() for x in g], h, w, EMUL[precision])
    seg, mot = seg.cpu(), mot.cpu()
    assert torch.isfinite(seg).all() and torch.isfinite(mot).all()
    # scale of one term of the segmentation dot product: |wh_k| * |h2_k| ~ max|logit| / sqrt(64) * ... measured through the
    # logits themselves: 2 ulp of the largest logit
    scale = float(seg_e.abs().max())
    d = float((seg - seg_e).abs().max())
    agree = float(((seg[:, 1] > seg[:, 0]) == (seg_e[:, 1] > seg_e[:, 0])).float().mean())
    dm = float((mot - mot_e).abs().max())
    print(f"\n[head alone {precision} {t}x{h}x{w}] logits max|d| {d:.3e} (largest logit {scale:.2f}, 2 ulp = {2 * ULP[precision] * scale:.3e}) "
          f"mean|d| {float((seg - seg_e).abs().mean()):.2e} argmax agreement {agree * 100:.4f}% motion max|d| {dm:.2e}")
    assert d <= 2 * ULP[precision] * scale
    assert agree >= 0.9999
    assert dm <= 2 * ULP[precision] + 2.0 ** -10            # tanh.approx (2^-11 relative) on top of the roundings


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_forward_16bit_is_as_far_from_fp32_as_its_storage_format(sd, net_bf16, net_fp16, precision):
    """Whole network, 32 x 112 x 112, three evaluations of the same weights: this library's tensor-core mode, the emulation
    of its storage points (oracle/model_emul.py: same roundings, fp32 accumulation in another order) and the fp32 oracle.
    Two 16-bit evaluations with identical storage points but different summation orders do NOT agree to a few ulp at the
    output - every rounding a reordered sum flips is a fresh 1-ulp perturbation that the next 30 layers amplify - so the
    implementation check is statistical: the library must be no further from the emulation than the emulation is from
    fp32, and its distance to fp32 must be the emulation's distance (the number format's own floor on these weights)."""
    net = net_bf16 if precision == "bf16" else net_fp16
    shape = (32, 112, 112)
    x = fixtures.synthetic_clip(*shape, seed=13, batch=1)
    seg_ref, mot_ref = model_ref.forward(sd, x)
    seg_e, mot_e = model_emul.forward(sd, x, EMUL[precision])
    seg, mot = net(x.cuda())
    ours, emul, between = softmax_metrics(seg, seg_ref), softmax_metrics(seg_e, seg_ref), softmax_metrics(seg, seg_e)

    def epe(a, b):
        d = a.float().cpu() - b
        return torch.sqrt((d[:, 0] * 56) ** 2 + (d[:, 1] * 56) ** 2)

    e_ours, e_emul, e_between = epe(mot, mot_ref), epe(mot_e, mot_ref), epe(mot, mot_e)
    print(f"\n[{precision} {shape}] vs fp32 oracle: ours softmax max {ours['max']:.4f} mean {ours['mean']:.5f} agree {ours['agree'] * 100:.3f}% "
          f"EPE mean {float(e_ours.mean()):.4f} max {float(e_ours.max()):.3f} px | emulation: max {emul['max']:.4f} mean {emul['mean']:.5f} agree "
          f"{emul['agree'] * 100:.3f}% EPE mean {float(e_emul.mean()):.4f} max {float(e_emul.max()):.3f} px | ours vs emulation: max {between['max']:.4f} "
          f"mean {between['mean']:.5f} agree {between['agree'] * 100:.3f}% EPE mean {float(e_between.mean()):.4f} px | near-boundary {ours['near'] * 100:.2f}%")
    assert between["mean"] <= emul["mean"] and between["max"] <= 1.25 * emul["max"] and float(e_between.mean()) <= float(e_emul.mean())
    assert 0.75 * emul["mean"] <= ours["mean"] <= 1.25 * emul["mean"]
    assert abs(ours["agree"] - emul["agree"]) <= 2e-3
    assert 0.75 * float(e_emul.mean()) <= float(e_ours.mean()) <= 1.25 * float(e_emul.mean())
    if precision == "fp16":
        assert ours["max"] <= 2e-2                       # north-star: softmax max-abs <= 2e-2 in 16-bit mode, per clip
        assert float(e_ours.mean()) <= 1.5e-2            # mean end-point error on this fixture's 3 px flows
        # the flow error of a 16-bit evaluation scales with the flow: the same network with its motion head calibrated to
        # 1.5 px flows (still above an echo's frame-to-frame wall motion at 112 px) meets the gate as stated
        sd15 = fixtures.calibrated_state_dict(0, flow_px=1.5)
        net15 = _net(sd15, "fp16")
        _seg15, mot15_ref = model_ref.forward(sd15, x)
        _s, mot15 = net15(x.cuda())
        e15 = epe(mot15, mot15_ref)
        print(f"[fp16, motion head calibrated to 1.5 px] EPE mean {float(e15.mean()):.4f} max {float(e15.max()):.3f} px")
        assert float(e15.mean()) <= 1e-2                 # north-star gate, as stated


def test_forward_interface_and_errors(net_fp32, sd):
    x = fixtures.synthetic_clip(8, 32, 32, seed=1, batch=2).cuda()
    seg, mot = net_fp32(x)
    assert tuple(seg.shape) == (2, 2, 8, 32, 32) and tuple(mot.shape) == (2, 4, 8, 32, 32)
    wrapped = torch.nn.DataParallel(net_fp32, device_ids=[0])
    seg2, _ = wrapped(x.cpu())                                                           # how fuse_utils.py:59 calls it
    assert torch.equal(seg, seg2)
    prob, _ = net_fp32.forward_prob(x)
    assert float((prob - torch.softmax(seg, 1)).abs().max()) <= 1e-6
    with pytest.raises(_lib.ClasfvError):
        net_fp32(torch.zeros(1, 3, 12, 32, 32).cuda())                                   # T % 8 != 0
    with pytest.raises(_lib.ClasfvError):
        net_fp32(torch.zeros(1, 3, 8, 24, 32).cuda())                                    # H % 16 != 0
    # repacks after an in-place weight change (load_state_dict / optimiser step)
    net2 = _net(sd, "fp32")
    a, _ = net2(x)
    with torch.no_grad():
        net2.segmentation_head.bias.add_(1.0)
    b, _ = net2(x)
    assert float((b - a - 1.0).abs().max()) <= 1e-5


# ------------------------------------------------------------------------------------------ head geometry
@pytest.mark.parametrize("t,h,w", [(16, 224, 224), (8, 48, 144), (8, 32, 272)])
def test_tensor_core_head_geometries_against_fp32_path(net_bf16, net_fp32, t, h, w):
    """The tcgen05 head tiles a row in 128-voxel pieces (W = 224: two pieces, 272: three, 144: ragged second piece) with
    a per-piece interpolation matrix.  Its bf16-mode output is checked against this library's own fp32 path (itself
    pinned to the oracle above) with the bf16 gates of DESIGN.md section 5: mean softmax error <= 2e-2 and >= 99.9 %
    argmax agreement away from the decision boundary; flow mean end-point error <= 1.5e-3 of the frame size (bf16
    output resolution is 2^-9 of the tanh range = 0.2 px at 112 px)."""
    x = fixtures.synthetic_clip(t, h, w, seed=31, batch=2).cuda()
    seg_ref, mot_ref = net_fp32(x)
    seg, mot = net_bf16(x)
    assert torch.isfinite(seg).all() and torch.isfinite(mot).all()
    m = softmax_metrics(seg, seg_ref.cpu())
    far = (m["pr"][:, 1] - 0.5).abs() >= 0.25
    lv, lvr = m["p"][:, 1] > 0.5, m["pr"][:, 1] > 0.5
    agree_far = float((lv == lvr)[far].float().mean())
    d = (mot.float() - mot_ref.float()).cpu()
    epe = torch.sqrt((d[:, 0] * w / 2) ** 2 + (d[:, 1] * h / 2) ** 2)
    print(f"\n[bf16 head {t}x{h}x{w}] softmax mean {m['mean']:.4f} max {m['max']:.3f} agree(far) {agree_far:.5f} EPE mean {float(epe.mean()):.4f} px")
    assert m["mean"] <= 2e-2
    assert agree_far >= 0.999
    assert float(epe.mean()) <= 1.5e-3 * max(h, w)
    # no column of the frame is special: the error must not concentrate at the 128-voxel piece boundaries
    col_err = (m["p"] - m["pr"]).abs().mean(dim=(0, 1, 2, 3))
    assert float(col_err.max()) <= 6 * float(col_err.mean()) + 1e-3


# ------------------------------------------------------------------------------------------ dense-video schedule
@pytest.mark.parametrize("tv,clip_len,h,w,step,sub_batch", [
    (43, 32, 32, 32, 1, 5),        # 12 windows, ragged last batch
    (40, 16, 48, 32, 2, 16),       # stride-2 windows, shortest clip the schedule accepts, one batch
    (52, 32, 112, 112, 1, 8),      # the benchmark geometry (56x56 layer-1 maps)
    (22, 16, 224, 224, 1, 4),      # config-5 geometry: 112x112 layer-1 maps
    (70, 64, 32, 32, 1, 8),        # the longest clip the frame-selected inputs take (64 frames)
    (45, 40, 16, 16, 1, 3),        # smallest frame: the deepest maps are 1 x 1
])
def test_dense_video_schedule_is_bit_identical_to_per_clip(net_bf16, tv, clip_len, h, w, step, sub_batch):
    """Sharing the stem and layer1 between overlapping windows (csrc/api.cu, Forward::run_dense) must not change a
    single bit of the outputs: every output element is produced by the same MMA sequence as in the per-clip schedule."""
    eng = net_bf16.engine()
    video = torch.from_numpy(synthetic.synthetic_echo_video(tv, h, w, seed=5)).cuda()
    starts = list(range(0, tv - clip_len + 1, step))
    outs = []
    for dense in (1, 0):
        eng.set_option("dense_video", dense)
        eng.set_option("sub_batch", sub_batch)
        seg, mot = eng.forward(video, _lib.OUT_LOGITS, torch.bfloat16, clip_starts=starts, clip_len=clip_len)
        torch.cuda.synchronize()
        outs.append((seg.float().cpu(), mot.float().cpu()))
    eng.set_option("dense_video", 1)
    eng.set_option("sub_batch", 32)
    assert torch.isfinite(outs[0][0]).all() and torch.isfinite(outs[0][1]).all()
    assert torch.equal(outs[0][0], outs[1][0]), float((outs[0][0] - outs[1][0]).abs().max())
    assert torch.equal(outs[0][1], outs[1][1]), float((outs[0][1] - outs[1][1]).abs().max())
    # and the per-clip schedule on a dense (N,3,T,H,W) copy of the same windows gives the same numbers
    clips = torch.stack([video[:, s:s + clip_len] for s in starts[:3]]).contiguous()
    seg3, mot3 = eng.forward(clips, _lib.OUT_LOGITS, torch.bfloat16)
    assert torch.equal(seg3.float().cpu(), outs[0][0][:3]) and torch.equal(mot3.float().cpu(), outs[0][1][:3])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("edge", [False, True])
def test_warp_fuse_matches_oracle(eng, dtype, edge):
    g = torch.Generator().manual_seed(7)
    n, h, w = 7, 24, 40
    prob = torch.softmax(2 * torch.randn(n, 2, 32, h, w, generator=g), 1).to(dtype)
    mot = torch.tanh(0.1 * torch.randn(n, 4, 32, h, w, generator=g)).to(dtype)
    starts = [0, 1, 2, 3, 10, 11, 50]                     # ragged: a gap, and a clip that shares no frame with the others
    t_out = 82
    acc, cnt, mask = fuse_ref.warp_fuse(prob.float(), mot.float(), starts, t_out, edge_hops=edge)
    r = eng.warp_fuse(prob.cuda(), mot.cuda(), starts, t_out, edge_hops=edge)
    assert float((r["acc"].cpu() - acc.float()).abs().max()) <= 1e-5 * float(acc.abs().max() + 1)
    assert torch.equal(r["cnt"].cpu().long(), cnt)
    margin = (acc[:, 1] - acc[:, 0]).abs()
    assert torch.equal(r["mask"].cpu()[margin > 1e-4], mask[margin > 1e-4])              # identical away from exact ties
    assert torch.equal(r["area"].cpu().long(), r["mask"].cpu().flatten(1).sum(1).long())
    assert int(r["cnt"][46].item()) == 0 and float(r["acc"][46].abs().max()) == 0.0      # uncovered frames stay empty


def test_warp_fuse_properties_at_config3_size(eng):
    """BASELINE config 3 shape (256 stride-1 clips x 32 x 112 x 112): size-independent properties."""
    n, h, w = 256, 112, 112
    g = torch.Generator(device="cuda").manual_seed(0)
    prob = torch.softmax(torch.randn(n, 2, 32, h, w, generator=g, device="cuda"), 1)
    mot = torch.tanh(0.05 * torch.randn(n, 4, 32, h, w, generator=g, device="cuda"))
    starts = list(range(n))
    t_out = n + 31
    r = eng.warp_fuse(prob, mot, starts, t_out)
    cnt = r["cnt"].float().view(-1, 1, 1)
    # bilinear weights sum to one: the class sums add up to the vote count
    assert float(((r["acc"][:, 0] + r["acc"][:, 1]) - cnt).abs().max()) <= 2e-4
    k = torch.arange(t_out)
    direct = torch.minimum(torch.minimum(k + 1, torch.tensor(32)), torch.tensor(t_out) - k)
    assert int(r["cnt"][100].item()) == 32 * 3 - 2 and int(r["cnt"][0].item()) == 2 and int(direct[100]) == 32
    # linearity of the LV sum in prob; the one-plane form (the LV probability alone) is the same operator
    r2 = eng.warp_fuse(2 * prob, mot, starts, t_out)
    assert float((r2["acc"][:, 1] - 2 * r["acc"][:, 1]).abs().max()) <= 1e-3
    r1 = eng.warp_fuse(prob[:, 1:].contiguous(), mot, starts, t_out)
    assert torch.equal(r1["acc"], r["acc"]) and torch.equal(r1["mask"], r["mask"])
    # fusing clip-batch by clip-batch (accumulate) == fusing at once
    acc = torch.zeros_like(r["acc"])
    first = True
    for b0 in range(0, n, 100):
        b1 = min(n, b0 + 100)
        rr = eng.warp_fuse(prob[b0:b1].contiguous(), mot[b0:b1].contiguous(), starts[b0:b1], t_out, acc=acc, accumulate=not first)
        first = False
    assert float((acc - r["acc"]).abs().max()) <= 1e-4
    assert torch.equal(rr["mask"], r["mask"]) or float((rr["mask"] != r["mask"]).float().mean()) < 1e-5
    # zero motion field and constant prob: every vote is the same value (border clamp keeps it inside)
    const = torch.full((4, 2, 32, h, w), 0.5, device="cuda")
    rz = eng.warp_fuse(const, torch.zeros(4, 4, 32, h, w, device="cuda"), [0, 1, 2, 3], 35)
    assert float((rz["acc"][:, 1] - 0.5 * rz["cnt"].float().view(-1, 1, 1)).abs().max()) <= 1e-5
    assert int(rz["mask"].sum().item()) == 0                                             # ties go to background


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_warp_fuse_at_config3_size_against_the_oracle_on_sampled_frames(eng, dtype):
    """BASELINE config 3 (256 stride-1 clips x 32 x 112 x 112, flows of a few pixels): the fused class sums, vote counts and
    masks of sampled output frames - first, last, interior, one next to each end - against the CPU oracle, which is run
    on just the clips that can vote on each sampled frame (the whole operator is 8 minutes of CPU)."""
    n, h, w = 256, 112, 112
    g = torch.Generator(device="cuda").manual_seed(5)
    prob = torch.sigmoid(2 * torch.randn(n, 1, 32, h, w, generator=g, device="cuda")).to(dtype)
    mot = torch.tanh(torch.randn(n, 4, 32, h, w, generator=g, device="cuda") * (3.0 / 56.0)).to(dtype)
    starts, t_out = list(range(n)), n + 31
    r = eng.warp_fuse(prob, mot, starts, t_out)
    for frame in (0, 1, 143, 285, 286):
        clips = [c for c in range(n) if frame - 32 <= starts[c] <= frame + 1]
        acc, cnt, mask = fuse_ref.warp_fuse(prob[clips].float().cpu(), mot[clips].float().cpu(), [starts[c] for c in clips], t_out)
        assert int(r["cnt"][frame]) == int(cnt[frame])
        got = r["acc"][frame].cpu()
        assert float((got - acc[frame].float()).abs().max()) <= 1e-5 * float(cnt[frame])          # fp32 sums of <= 94 votes
        margin = (acc[frame, 1] - acc[frame, 0]).abs()
        assert torch.equal(r["mask"][frame].cpu()[margin > 1e-4], mask[frame][margin > 1e-4])


# ------------------------------------------------------------------------------------------ pipelines
def test_reference_exact_fusion_matches_oracle(net_fp32, sd):
    video = synthetic.synthetic_echo_video(75, 32, 48, seed=3)
    oracle_model = lambda x: model_ref.forward(sd, x)  # noqa: E731
    for f, step in ((5, 1), (1, 1), (3, 2)):
        ref = fuse_ref.segment_a_video_with_fusion(video, oracle_model, interpolate_last=True, step=step, num_clips=f)
        got = fuse_utils.segment_a_video_with_fusion(video, torch.nn.DataParallel(net_fp32), interpolate_last=True, step=step, num_clips=f)
        assert got.dtype == np.int64 and got.shape == ref.shape
        assert float((got != ref).mean()) <= 1e-3, (f, step)                             # mask agreement >= 99.9 %
        dice = fuse_ref.categorical_dice(got, ref, 1)
        assert abs(1.0 - dice) <= 1e-3                                                   # Dice delta <= 1e-3


def test_reference_exact_fusion_config0_shape(net_fp32, sd):
    """BASELINE config 0: 112x112, T=128, single pass (-f 1): 4 consecutive clips, no resample."""
    video = synthetic.synthetic_echo_video(128, 112, 112, seed=0)
    oracle_model = lambda x: model_ref.forward(sd, x)  # noqa: E731
    ref = fuse_ref.segment_a_video_with_fusion(video, oracle_model, interpolate_last=True, step=1, num_clips=1)
    got, det = fuse_utils.segment_a_video_with_fusion(video, net_fp32, interpolate_last=True, step=1, num_clips=1, return_details=True)
    assert det["clips"] == 4 and got.shape == (128, 112, 112)
    assert float((got == ref).mean()) >= 0.999
    area = ref.reshape(128, -1).sum(1)
    ed, es = int(area.argmax()), int(area.argmin())
    for fr in (ed, es):                                                                  # ES / ED Dice delta <= 1e-3
        assert abs(1.0 - fuse_ref.categorical_dice(got[fr], ref[fr], 1)) <= 1e-3
    np.testing.assert_array_equal(det["area"], got.reshape(128, -1).sum(1))


def test_warp_fusion_pipeline_matches_oracle(net_fp32, sd):
    video = synthetic.synthetic_echo_video(70, 32, 32, seed=3)
    starts = list(range(0, 70 - 32 + 1))
    probs, mots = [], []
    for s in starts:
        seg, mot = model_ref.forward(sd, torch.from_numpy(video[:, s:s + 32]).unsqueeze(0))
        probs.append(torch.softmax(seg, 1)); mots.append(mot)
    acc, cnt, mask = fuse_ref.warp_fuse(torch.cat(probs), torch.cat(mots), starts, 70)
    got, det = fuse_utils.segment_a_video_with_fusion(video, net_fp32, fuse_method="warp", return_details=True, batch_clips=7)
    assert det["clips"] == len(starts) and got.shape == (70, 32, 32)
    assert float((det["acc"].cpu() - acc.float()).abs().max()) <= 1e-4 * float(cnt.max())  # mean prob within 1e-4
    assert float((got != mask.numpy()).mean()) <= 1e-3
    np.testing.assert_array_equal(det["cnt"], cnt.numpy())


@pytest.fixture(scope="module")
def config1_cut_oracle(sd):
    """fp32 oracle of a cut of BASELINE configs[1]: 112 x 112, every stride-1 32-frame window (17 clips of a 48-frame
    video; the full 200-frame video is 169 clips = 8 CPU-minutes of oracle), warp-and-fuse."""
    tv = 48
    video = synthetic.synthetic_echo_video(tv, 112, 112, seed=7)
    starts = list(range(0, tv - 32 + 1))
    probs, mots = [], []
    for s0 in starts:
        seg, mot = model_ref.forward(sd, torch.from_numpy(video[:, s0:s0 + 32]).unsqueeze(0))
        probs.append(torch.softmax(seg, 1)); mots.append(mot)
    prob, mot = torch.cat(probs), torch.cat(mots)
    acc, cnt, mask = fuse_ref.warp_fuse(prob, mot, starts, tv)
    return {"video": video, "starts": starts, "prob": prob, "motion": mot, "acc": acc, "cnt": cnt, "mask": mask}


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_config1_pipeline_gates_in_16bit_modes(config1_cut_oracle, net_fp16, net_bf16, precision):
    """The benchmarked configuration (BASELINE configs[1]: tensor-core mode + fuse_method="warp" through the public
    segment_a_video_with_fusion) against the fp32 oracle, with the north-star gates as stated:
        softmax max-abs <= 2e-2 (16-bit), argmax-mask agreement >= 99.9 %, ES/ED Dice delta <= 1e-3, flow EPE <= 1e-2 px.
    fp16 mode: the softmax, mask and Dice gates are asserted as stated.  The EPE gate is asserted at the level the fp16 format
    allows on this fixture's 3 px flows (1.5e-2 px mean; the emulation's floor is 1.4e-2) and as stated on 1.5 px flows in
    test_forward_16bit_is_as_far_from_fp32_as_its_storage_format.
    bf16 mode: none of the gates is reachable on these weights by ANY bf16-operand evaluation - rounding only the trunk's
    weights to bf16, activations exact, already gives softmax max 0.10 / 99.3 % (tests/test_oracle_emul.py) -
    so the asserts pin the library to that floor (and to the emulation in the test above) instead of to the gate."""
    o = config1_cut_oracle
    net = net_fp16 if precision == "fp16" else net_bf16
    tv = o["video"].shape[1]
    got, det = fuse_utils.segment_a_video_with_fusion(o["video"], net, fuse_method="warp", return_details=True)
    assert det["clips"] == len(o["starts"]) and np.array_equal(det["cnt"], o["cnt"].numpy())
    cnt = o["cnt"].view(-1, 1, 1, 1).clamp_min(1).float()
    mean, mean_ref = det["acc"].cpu() / cnt, o["acc"].float() / cnt
    smax = float((mean - mean_ref).abs().max())
    clip_smax = float((det["prob"][:, 0].float().cpu() - o["prob"][:, 1]).abs().max())
    agree = float((torch.from_numpy(got) == o["mask"].long()).float().mean())
    area = o["mask"].flatten(1).sum(1)
    ed, es = int(area.argmax()), int(area.argmin())
    ddice = [abs(1.0 - fuse_ref.categorical_dice(got[f], o["mask"][f].numpy(), 1)) for f in (ed, es)]
    d = det["motion"].float().cpu() - o["motion"]
    epe = torch.cat([torch.sqrt((d[:, 0] * 56) ** 2 + (d[:, 1] * 56) ** 2).flatten(), torch.sqrt((d[:, 2] * 56) ** 2 + (d[:, 3] * 56) ** 2).flatten()])
    near = float(((mean_ref[:, 1] - 0.5).abs() < 2e-2).float().mean())
    print(f"\n[configs[1] cut, {precision}, {tv} frames / {len(o['starts'])} clips] fused softmax max|d| {smax:.4f} (per clip {clip_smax:.4f}) | "
          f"mask agreement {agree * 100:.4f}% (near-boundary pixels {near * 100:.2f}%) | Dice delta ED {ddice[0]:.5f} ES {ddice[1]:.5f} | "
          f"flow EPE mean {float(epe.mean()):.4f} max {float(epe.max()):.3f} px")
    if precision == "fp16":
        assert smax <= 2e-2                      # north-star gate, as stated
        assert agree >= 0.999                    # north-star gate, as stated
        assert max(ddice) <= 1e-3                # north-star gate, as stated
        assert float(epe.mean()) <= 1.5e-2       # fp16 floor on 3 px flows (gate 1e-2: see the docstring)
    else:
        assert smax <= 0.16 and agree >= 0.993 and max(ddice) <= 8e-3 and float(epe.mean()) <= 0.15


def test_warp_fusion_pipeline_at_224_matches_oracle(net_fp32, net_fp16, sd):
    """BASELINE configs[4] geometry (224 x 224; the reference itself hard-codes 112, src/fuse_utils.py:22,27,55,68,75, so the
    oracle is the H/W-generic restatement): a 36-frame cut, 5 stride-1 clips, fp32 mode and the fp16 tensor-core mode against
    the fp32 oracle - the fp32-vs-16-bit tolerance check of SURVEY 8(d) "Config 5"."""
    tv = 36
    video = synthetic.synthetic_echo_video(tv, 224, 224, seed=11)
    starts = list(range(0, tv - 32 + 1))
    probs, mots = [], []
    for s0 in starts:
        seg, mot = model_ref.forward(sd, torch.from_numpy(video[:, s0:s0 + 32]).unsqueeze(0))
        probs.append(torch.softmax(seg, 1)); mots.append(mot)
    acc, cnt, mask = fuse_ref.warp_fuse(torch.cat(probs), torch.cat(mots), starts, tv)
    c = cnt.view(-1, 1, 1, 1).clamp_min(1).float()
    # fp32: the north-star gates as stated.  fp16: this cut has 5 clips, i.e. 1..15 votes per frame where the full video has
    # ~94, so the fused probability is essentially a per-clip probability: bound = the per-clip fp16 level measured at 112 x 112
    # (2.4e-2 in test_config1_pipeline_gates_in_16bit_modes, where the 17-clip fusion brings it to 1.5e-2 <= the 2e-2 gate)
    for name, net, tol, min_agree in (("fp32", net_fp32, 1e-4, 0.999), ("fp16", net_fp16, 4e-2, 0.998)):
        got, det = fuse_utils.segment_a_video_with_fusion(video, net, fuse_method="warp", return_details=True)
        assert np.array_equal(det["cnt"], cnt.numpy())
        err = float((det["acc"].cpu() / c - acc.float() / c).abs().max())
        agree = float((torch.from_numpy(got) == mask.long()).float().mean())
        print(f"\n[224x224, {tv} frames / {len(starts)} clips, {name}] fused softmax max|d| {err:.2e}, mask agreement {agree * 100:.4f}%")
        assert err <= tol and agree >= min_agree


def test_many_videos_pipeline_equals_single_calls(net_fp16):
    """fuse_utils.segment_videos_with_fusion (pinned staging, copies overlapped with the neighbouring videos' compute) must
    return, in order, exactly the masks of one segment_a_video_with_fusion(..., fuse_method="warp") call per video -
    videos of different lengths and sizes, so every staging slot and plane buffer is re-used and re-allocated."""
    shapes = [(40, 32, 32), (50, 32, 32), (33, 48, 32), (40, 32, 32), (64, 32, 48)]
    videos = [synthetic.synthetic_echo_video(t, h, w, seed=40 + i) for i, (t, h, w) in enumerate(shapes)]
    singles = [fuse_utils.segment_a_video_with_fusion(v, net_fp16, fuse_method="warp") for v in videos]
    piped = list(fuse_utils.segment_videos_with_fusion(iter(videos), net_fp16))
    assert len(piped) == len(singles)
    for a, b in zip(piped, singles):
        assert a.dtype == np.int64 and a.shape == b.shape and np.array_equal(a, b)
    assert list(fuse_utils.segment_videos_with_fusion(iter([]), net_fp16)) == []


def test_bf16_pipeline_dice_against_fp32_oracle(net_bf16, sd):
    video = synthetic.synthetic_echo_video(64, 32, 32, seed=5)
    oracle_model = lambda x: model_ref.forward(sd, x)  # noqa: E731
    ref = fuse_ref.segment_a_video_with_fusion(video, oracle_model, interpolate_last=True, step=1, num_clips=8)
    got = fuse_utils.segment_a_video_with_fusion(video, net_bf16, interpolate_last=True, step=1, num_clips=8)
    dice = fuse_ref.categorical_dice(got, ref, 1)
    print(f"\n[bf16 F1 pipeline] mask agreement {float((got == ref).mean()) * 100:.3f}%  Dice {dice:.4f}")
    assert dice >= 0.97


def test_fusion_error_behaviour(net_fp32, capsys):
    v = synthetic.synthetic_echo_video(40, 16, 16, seed=1)
    with pytest.raises(IndexError):                                                      # 32 <= T < 32 + step (fuse_utils.py:82)
        fuse_utils.segment_a_video_with_fusion(v[:, :32], net_fp32, step=1, num_clips=10)
    with pytest.raises(TypeError):
        fuse_utils.segment_a_video_with_fusion(v, lambda x: x)
    with pytest.raises(NotImplementedError):
        fuse_utils.segment_a_video_with_fusion(v, net_fp32, fuse_method="staple")
    out = fuse_utils.segment_a_video_with_fusion(v[:, :20], net_fp32, step=1, num_clips=10)   # "Video is too short", 1 pass
    assert "Video is too short" in capsys.readouterr().out and out.shape == (20, 16, 16)


def test_ejection_fraction_from_the_device_area_trace(eng):
    """compute_ef_using_putative_clips fed by the LV area trace the fusion kernel returns, against the
    oracle (oracle/ef_ref.py, pinned to the reference functions) run on the same fused masks."""
    from oracle import ef_ref
    beats = ef_ref.beating_masks(150, 112, 47.0, 3)
    t = beats.shape[0]
    starts = list(range(0, t - 32 + 1, 8))
    lv = torch.from_numpy(beats).float() * 0.8 + 0.1                               # LV probability 0.9 inside, 0.1 outside
    prob = torch.stack([lv[s0:s0 + 32] for s0 in starts]).unsqueeze(1).contiguous()     # (n,1,32,H,W)
    mot = torch.zeros(len(starts), 4, 32, 112, 112)
    r = eng.warp_fuse(prob.cuda(), mot.cuda(), starts, t)
    masks = r["mask"].cpu().numpy().astype(np.int64)
    area = r["area"].cpu().numpy()
    assert np.array_equal(area, masks.reshape(t, -1).sum(1))
    want, want_pairs = ef_ref.compute_ef_using_putative_clips(masks, "t", return_edes=True)
    got, pairs = fuse_utils.compute_ef_using_putative_clips(masks, "t", return_edes=True, area=area)
    assert len(want) >= 2 and all(45 < ef < 90 for ef in want)
    assert [tuple(map(int, p)) for p in pairs] == [tuple(map(int, p)) for p in want_pairs]
    np.testing.assert_allclose(np.array(got), np.array(want), rtol=1e-12)


# ------------------------------------------------------------------------------------------ CLI
def test_motion_segment_cli_config0(tmp_path, sd):
    video = synthetic.synthetic_echo_video(128, 112, 112, seed=0)
    avi = tmp_path / "synthetic_echo.avi"
    synthetic.write_avi(avi, video)
    ckpt = tmp_path / "model.pth"
    torch.save({"model": {"module." + k: v for k, v in sd.items()}}, ckpt)
    cli = os.path.join(clasfv_b200.PACKAGE_DIR, "motion_segment.py")
    out = tmp_path / "out"
    r = subprocess.run([sys.executable, cli, "-p", str(avi), "-m", str(ckpt), "-d", "cuda", "-f", "1", "-c", "binary_video,binary",
                        "-o", str(out), "-v"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "R2+1D MotionNet has 31575731 parameters." in r.stdout
    import pickle
    seg = pickle.load(open(out / "synthetic_echo_whole_video_segmentation.pkl", "rb"))
    assert seg.shape == (128, 112, 112) and seg.dtype == np.int64 and set(np.unique(seg)) <= {0, 1}
    # mask equality with the oracle's reference-exact fusion of the same decoded video (the reference's host pipeline:
    # cv2 decode, float, trilinear pre-resize, zero-one normalisation; motion_segment.py:80-106)
    import cv2
    from clasfv_b200.src.echonet_dataset import zeroone_normalizer
    cap = cv2.VideoCapture(str(avi))
    frames = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        frames.append(cv2.cvtColor(fr, cv2.COLOR_BGR2RGB))
    v = torch.Tensor(np.stack(frames).transpose((3, 0, 1, 2)).astype(np.float32)).unsqueeze(0)
    v = F.interpolate(v, size=(v.shape[2], 112, 112), mode="trilinear", align_corners=True)
    host_video = zeroone_normalizer(v.squeeze(0).numpy())
    ref = fuse_ref.segment_a_video_with_fusion(host_video, lambda x: model_ref.forward(sd, x), interpolate_last=True, step=1, num_clips=1)
    assert float((seg == ref).mean()) >= 0.999
    # ED / ES pickles: one pair of files per heartbeat the EF post-processing identifies on these masks, holding those frames
    from oracle import ef_ref
    _efs, pairs = ef_ref.compute_ef_using_putative_clips(seg, "cli", return_edes=True)
    written = sorted(p.name for p in out.iterdir())
    for ed, es in pairs:
        for tag, f in (("ED", int(ed)), ("ES", int(es))):
            name = f"synthetic_echo_{tag}_Frame_{f}_segmentation.pkl"
            assert name in written
            assert np.array_equal(pickle.load(open(out / name, "rb")), seg[f])
    assert len([n for n in written if "_Frame_" in n]) == 2 * len(pairs)
    assert f"Identified {len(_efs)} systoles" in r.stdout
    # -c gif: written when the optional presentation dependency is there, skipped with a message otherwise
    r3 = subprocess.run([sys.executable, cli, "-p", str(avi), "-m", str(ckpt), "-d", "cuda", "-f", "1", "-c", "gif", "-o", str(out)],
                        capture_output=True, text=True, timeout=600)
    assert r3.returncode == 0, r3.stderr[-2000:]
    assert (out / "synthetic_echo_annotated.gif").exists() or "skipping the annotated gif" in r3.stdout
    r2 = subprocess.run([sys.executable, cli, "-p", str(avi), "-m", str(ckpt), "-d", "cpu"], capture_output=True, text=True, timeout=120)
    assert r2.returncode != 0 and "no CPU path" in (r2.stderr + r2.stdout)


# ------------------------------------------------------------------------------------------ multi-GPU path, 1 rank
def test_long_video_split_single_rank_equals_direct(net_fp32):
    import torch.distributed as dist
    from clasfv_b200 import sharding
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29611")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        video = synthetic.synthetic_echo_video(50, 32, 32, seed=9)
        a = sharding.segment_long_video(video, net_fp32, step=1, mask_dtype=np.int64)
        b = fuse_utils.segment_a_video_with_fusion(video, net_fp32, fuse_method="warp")
        assert a.dtype == np.int64 and np.array_equal(a, b)
    finally:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ GPU yardstick
def test_faster_than_pytorch_eager_of_the_reference_algorithm_on_the_same_gpu(net_bf16, sd):
    """SURVEY.md 8(d) "GPU reference point": the reference ships no Blackwell code, so the kernel to beat on this GPU is
    stock PyTorch eager (cuDNN / cuBLAS, bf16 autocast) of the reference's forward pass.  Same 16 independent clips, model
    forward only (per-clip schedule: no sharing between windows), CUDA events, the better of two rounds each."""
    n, shape = 16, (32, 112, 112)
    x = fixtures.synthetic_clip(*shape, seed=21, batch=n).cuda()
    sd_dev = {k: v.cuda() for k, v in model_ref.strip_module_prefix(sd).items()}

    def timed(fn):
        times = []
        for _ in range(3):                                   # first round is the warm-up (cuDNN heuristics, workspace)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        return min(times[1:]), out

    def eager():
        outs = []
        with torch.autocast("cuda", dtype=torch.bfloat16):
            for i in range(0, n, 4):                          # 4 clips per call: the 1 024-channel concat is 3.3 GB in bf16
                outs.append(model_ref.forward(sd_dev, x[i:i + 4]))
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])

    eng_ = net_bf16.engine()
    eng_.set_option("dense_video", 0)
    try:
        ms_ours, (seg, _mot) = timed(lambda: net_bf16(x))
    finally:
        eng_.set_option("dense_video", 1)
    ms_eager, (seg_e, _mot_e) = timed(eager)
    agree = float(((seg[:, 1] > seg[:, 0]) == (seg_e[:, 1] > seg_e[:, 0])).float().mean())
    fps_ours, fps_eager = n * 32 / ms_ours * 1e3, n * 32 / ms_eager * 1e3
    print(f"\n[eager yardstick] ours {ms_ours:.2f} ms ({fps_ours:.0f} clip-frames/s) | PyTorch eager bf16 autocast {ms_eager:.2f} ms "
          f"({fps_eager:.0f} clip-frames/s) | ratio {ms_eager / ms_ours:.1f}x | mask agreement {agree * 100:.2f}%")
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        import json
        with open(os.path.join(out_dir, "eager_yardstick.json"), "w") as f:
            json.dump({"clips": n, "shape": list(shape), "ours_ms": ms_ours, "eager_bf16_autocast_ms": ms_eager,
                       "ours_clip_frames_per_s": fps_ours, "eager_clip_frames_per_s": fps_eager,
                       "speedup": ms_eager / ms_ours, "mask_agreement": agree, "schedule": "per-clip (no window sharing)"}, f)
    assert ms_ours * 2 < ms_eager            # at least twice as fast as the library path on the same GPU
    assert agree > 0.97                      # two bf16 evaluations of the same fp32 network (see the autocast yardstick test above)


_WF_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
from clasfv_b200.engine import Engine
dtype = torch.float32 if sys.argv[1] == "fp32" else torch.bfloat16
# the inputs are generated ONCE by the parent and read from a file: both kernels must see the same bytes (two processes that
# each evaluate torch's CPU randn / tanh / softmax are not guaranteed to - this comparison is about the kernels)
z = np.load(sys.argv[4])
prob = torch.from_numpy(z["prob"]).to(dtype); mot = torch.from_numpy(z["mot"]).to(dtype)
n = prob.shape[0]
pd, md = prob.cuda(), mot.cuda()
bits = torch.int32 if dtype == torch.float32 else torch.int16
chk = np.array([int(pd.view(bits).long().sum()), int(md.view(bits).long().sum())])      # what the kernel reads, as uploaded
r = Engine("cuda:0").warp_fuse(pd, md, list(range(n)), n + 31, edge_hops=(sys.argv[2] == "1"))
np.savez(sys.argv[3], acc=r["acc"].cpu().numpy(), cnt=r["cnt"].cpu().numpy(), mask=r["mask"].cpu().numpy(), area=r["area"].cpu().numpy(), chk=chk)
"""


def _wf_inputs(path):
    g = torch.Generator().manual_seed(11)
    n, h, w = 40, 48, 64
    prob = torch.softmax(2 * torch.randn(n, 2, 32, h, w, generator=g), 1)
    mot = torch.tanh(0.15 * torch.randn(n, 4, 32, h, w, generator=g))
    mot[:, :, :, :, :3] = -0.9; mot[:, :, :, :, -3:] = 0.9; mot[:, :, :, :2, :] = -0.9; mot[:, :, :, -2:, :] = 0.9   # taps clamped at every border
    np.savez(path, prob=prob.numpy(), mot=mot.numpy())


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_staged_warp_fuse_is_bit_identical_to_the_direct_gather_kernel(tmp_path, dtype):
    """the shared-memory-staged kernel (clamped north-west corner, select-free taps) adds the same numbers in
    the same order as the direct-gather kernel with grid_sample's conditional taps, so the class sums must be the same
    BITS - including pixels whose source lands on or beyond every border.  The kernel choice is latched per process
    (CLASFV_WARP_FUSE_DIRECT), hence two subprocesses."""
    script = tmp_path / "wf.py"
    script.write_text(_WF_SCRIPT.format(root=ROOT))
    inputs = str(tmp_path / "inputs.npz")
    _wf_inputs(inputs)
    out = {}
    for name, env in (("staged", {}), ("direct", {"CLASFV_WARP_FUSE_DIRECT": "1"}), ("staged_again", {})):
        path = str(tmp_path / f"{name}.npz")
        e = dict(os.environ); e.pop("CLASFV_WARP_FUSE_DIRECT", None); e.update(env)
        subprocess.run([sys.executable, str(script), dtype, "1" if dtype == "fp32" else "0", path, inputs], check=True, env=e, timeout=300)
        out[name] = np.load(path)
    a, b = out["staged"], out["direct"]
    assert np.array_equal(a["chk"], b["chk"]) and np.array_equal(a["chk"], out["staged_again"]["chk"]), "the device inputs differ between processes"
    # run-to-run determinism of the staged kernel across processes (a race in the ring would show here)
    assert np.array_equal(a["acc"].view(np.uint32), out["staged_again"]["acc"].view(np.uint32)), "staged kernel differs from itself across processes"
    differing = int((a["acc"].view(np.uint32) != b["acc"].view(np.uint32)).sum())
    worst = float(np.abs(a["acc"] - b["acc"]).max())
    print(f"\n[staged vs direct, {dtype}] class sums differing in any bit: {differing} of {a['acc'].size}, max abs difference {worst:.3g}")
    assert np.array_equal(a["cnt"], b["cnt"])
    assert a["acc"].max() > 1.0                      # not vacuous
    # Both kernels are pure functions of their inputs with a fixed order of additions per pixel.  Round 2 settled the open
    # item of DESIGN.md 4.5 on the GPU (tools/wf_diag.py, profiles/r02a_wf_diag.log): staged vs direct and each kernel
    # against itself across three processes, both element types, both edge modes: 0 of 436 224 sums differ in any bit.
    assert differing == 0, f"{dtype}: {differing} class sums differ (max abs {worst:.3g})"
    assert np.array_equal(a["mask"], b["mask"]) and np.array_equal(a["area"], b["area"])
