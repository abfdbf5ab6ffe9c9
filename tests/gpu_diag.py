"""GPU bring-up diagnostics: prints numeric error of every kernel family against the oracle.

Not collected by pytest (no test_ prefix).  Usage on the B200 box:

    python tests/gpu_diag.py            # all sections, each in its own subprocess with a timeout
    python tests/gpu_diag.py conv_umma  # one section in-process

Sections run in separate processes because a device-side trap poisons the CUDA context.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SECTIONS = ["fusion_small", "conv_simt", "conv_umma", "forward_fp32", "forward_bf16_simt", "forward_bf16", "pipeline"]


def _stats(name, got, ref, extra=""):
    import torch
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    d = (got - ref).abs()
    print(f"  {name:44s} max|d|={d.max().item():.3e} mean|d|={d.mean().item():.3e} max|ref|={ref.abs().max().item():.3e} {extra}",
          flush=True)
    return d.max().item()


def conv_cases():
    # name, N,T,H,W,Cin,Cout,k,stride,pad, residual, relu, out_f32
    return [
        ("gemm 1x1x1 64->64 (1 tile)", 1, 2, 8, 8, 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), False, False, False),
        ("gemm 1x1x1 128->64 (2 slabs)", 1, 2, 8, 8, 128, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), False, False, False),
        ("gemm 1x1x1 64->64 f32 out + res", 1, 2, 8, 8, 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), True, False, True),
        ("temporal 3x1x1 144->64 pad1", 1, 4, 8, 8, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), False, True, False),
        ("spatial 1x3x3 64->144 pad1", 1, 2, 16, 16, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, True, False),
        ("spatial s2 1x3x3 64->240", 1, 2, 16, 16, 64, 240, (1, 3, 3), (1, 2, 2), (0, 1, 1), False, True, False),
        ("temporal s2 3x1x1 240->128", 1, 8, 8, 8, 240, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0), False, True, False),
        ("downsample 1x1x1 s2 64->128", 1, 4, 16, 16, 64, 128, (1, 1, 1), (2, 2, 2), (0, 0, 0), False, False, False),
        ("spatial 1x3x3 256->576 7x7 n2", 2, 2, 7, 7, 256, 576, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, True, False),
        ("temporal 3x1x1 576->256 +res relu", 2, 4, 7, 7, 576, 256, (3, 1, 1), (1, 1, 1), (1, 0, 0), True, True, False),
        ("spatial 1x3x3 512->1152 4x7x7", 1, 4, 7, 7, 512, 1152, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, True, False),
        ("spatial 1x3x3 64->144 56x56 (196 tiles)", 1, 8, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, True, False),
        ("temporal 3x1x1 144->64 56x56 +res", 1, 8, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), True, True, False),
    ]


def run_conv(engine_name, dtype_name):
    import torch
    import torch.nn.functional as F
    from clasfv_b200.engine import Engine
    eng = Engine("cuda:0")
    dtype = torch.float32 if dtype_name == "fp32" else torch.bfloat16
    worst = 0.0
    for (name, n, t, h, w, cin, cout, k, s, p, use_res, relu, out_f32) in conv_cases():
        g = torch.Generator().manual_seed(hash(name) % 1000)
        x = torch.randn(n, t, h, w, cin, generator=g)
        wt = torch.randn(cout, cin, *k, generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5
        scale = 0.5 + torch.rand(cout, generator=g)
        shift = 0.2 * torch.randn(cout, generator=g)
        xq = x.to(dtype).float()
        wq = (wt * scale.view(-1, 1, 1, 1, 1)).to(dtype).float()
        ref = F.conv3d(xq.permute(0, 4, 1, 2, 3), wq, shift, s, p).permute(0, 2, 3, 4, 1).contiguous()
        res = None
        if use_res:
            res_dtype = torch.float32 if (out_f32 or dtype == torch.float32) else torch.bfloat16
            res = (0.5 * torch.randn(ref.shape, generator=g)).to(res_dtype)
            ref = ref + res.float()
        if relu:
            ref = ref.relu()
        try:
            out = eng.conv3d(x.to(dtype).cuda(), wt, scale, shift, s, p, res.cuda() if res is not None else None, relu,
                             engine=engine_name, out_f32=out_f32)
            torch.cuda.synchronize()
            worst = max(worst, _stats(f"[{engine_name}/{dtype_name}] {name}", out, ref))
        except Exception as e:  # noqa: BLE001
            print(f"  [{engine_name}/{dtype_name}] {name}: FAILED {type(e).__name__}: {e}", flush=True)
            return 1
    print(f"  worst max|d| = {worst:.3e}")
    return 0


def section_fusion_small():
    import numpy as np
    import torch
    from clasfv_b200 import engine as E
    from oracle import fuse_ref
    g = torch.Generator().manual_seed(0)
    src = torch.rand(3, 2, 24, 40, generator=g)
    flow = torch.tanh(0.1 * torch.randn(3, 2, 24, 40, generator=g))
    flow[0, :, 0, 0] = torch.tensor([-0.9, 0.9])
    _stats("warp", E.warp(src.cuda(), flow.cuda()), fuse_ref.warp(src, flow))
    _stats("motion_field", E.motion_field(flow.cuda(), 24, 40), fuse_ref.generate_2dmotion_field(src, flow))
    x = torch.rand(3, 75, 16, 16, generator=g)
    _stats("temporal_resample 75->64", E.temporal_resample(x.cuda(), 64), torch.from_numpy(fuse_ref.temporal_resample(x.numpy(), 64)))
    _stats("temporal_resample 64->75", E.temporal_resample(x[:, :64].contiguous().cuda(), 75),
           torch.from_numpy(fuse_ref.temporal_resample(x[:, :64].numpy(), 75)))
    eng = E.Engine("cuda:0")
    # F2 small
    n, h, w = 6, 16, 32
    prob = torch.softmax(2 * torch.randn(n, 2, 32, h, w, generator=g), 1)
    mot = torch.tanh(0.08 * torch.randn(n, 4, 32, h, w, generator=g))
    starts = [0, 1, 2, 5, 6, 9]
    for edge in (False, True):
        acc, cnt, mask = fuse_ref.warp_fuse(prob, mot, starts, 43, edge_hops=edge)
        r = eng.warp_fuse(prob.cuda(), mot.cuda(), starts, 43, edge_hops=edge)
        _stats(f"warp_fuse acc edge={edge}", r["acc"], acc.float())
        print("    cnt equal:", bool((r["cnt"].cpu() == cnt.int()).all()), " mask mismatches:", int((r["mask"].cpu() != mask).sum()),
              " area ok:", bool((r["area"].cpu() == mask.flatten(1).sum(1).int()).all()))
    rb = eng.warp_fuse(prob.cuda().bfloat16(), mot.cuda().bfloat16(), starts, 43)
    accb, _, _ = fuse_ref.warp_fuse(prob.bfloat16().float(), mot.bfloat16().float(), starts, 43)
    _stats("warp_fuse acc bf16 inputs", rb["acc"], accb.float())
    # F1 pieces
    video = torch.rand(3, 75, 16, 16, generator=g)
    plan = [(0, 75, 2), (1, 74, 2), (11, 64, 2)]
    clips = eng.build_shift_clips(video.cuda(), plan)
    ref = np.concatenate([fuse_ref.divide_to_consecutive_clips(video[:, s:].numpy(), interpolate_last=True) for s, _l, _n in plan])
    _stats("build_shift_clips", clips, torch.from_numpy(ref).float())
    return 0


def section_conv_simt():
    return run_conv("simt", "fp32") or run_conv("simt", "bf16")


def section_conv_umma():
    return run_conv("umma", "bf16")


def _forward(precision, shapes, force_simt=False):
    import torch
    from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
    from oracle import fixtures, model_ref
    if force_simt:
        os.environ["CLASFV_FORCE_SIMT"] = "1"
    sd = fixtures.calibrated_state_dict(0)
    net = R2plus1D_18_MotionNet(pretrained=False, precision=precision)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    for shape, seed, batch in shapes:
        x = fixtures.synthetic_clip(*shape, seed=seed, batch=batch)
        t0 = time.time()
        seg_ref, mot_ref = model_ref.forward(sd, x)
        t_ref = time.time() - t0
        seg, mot = net(x.cuda())
        torch.cuda.synchronize()
        t0 = time.time()
        seg, mot = net(x.cuda())
        torch.cuda.synchronize()
        t_gpu = time.time() - t0
        print(f" shape {shape} batch {batch}: oracle {t_ref:.2f}s  gpu {t_gpu * 1e3:.1f} ms", flush=True)
        _stats("seg logits", seg, seg_ref)
        _stats("motion (tanh)", mot, mot_ref, extra=f"EPE px max={(mot.cpu() - mot_ref).abs().max().item() * shape[2] / 2:.3e}")
        p = torch.softmax(seg.float().cpu(), 1)
        pr = torch.softmax(seg_ref, 1)
        _stats("softmax", p, pr)
        agree = ((p[:, 1] > p[:, 0]) == (pr[:, 1] > pr[:, 0])).float().mean().item()
        print(f"    argmax agreement {agree * 100:.4f}%   near-boundary(|p-.5|<2e-2) {((pr[:, 1] - 0.5).abs() < 2e-2).float().mean().item() * 100:.3f}%")
    return 0


def section_forward_fp32():
    return _forward("fp32", [((8, 32, 32), 11, 1), ((16, 48, 32), 12, 2), ((32, 112, 112), 13, 1)])


def section_forward_bf16_simt():
    return _forward("bf16", [((8, 32, 32), 11, 1), ((32, 112, 112), 13, 1)], force_simt=True)


def section_forward_bf16():
    return _forward("bf16", [((8, 32, 32), 11, 1), ((16, 48, 32), 12, 2), ((32, 112, 112), 13, 2)])


def section_pipeline():
    import numpy as np
    import torch
    import clasfv_b200.synthetic as synthetic
    from clasfv_b200.src import fuse_utils
    from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
    from oracle import fixtures, fuse_ref, model_ref
    sd = fixtures.calibrated_state_dict(0)
    net = R2plus1D_18_MotionNet(pretrained=False, precision="fp32")
    net.load_state_dict(sd)
    net = torch.nn.DataParallel(net).cuda().eval()
    video = synthetic.synthetic_echo_video(70, 32, 32, seed=3)
    oracle_model = lambda x: model_ref.forward(sd, x)  # noqa: E731
    t0 = time.time()
    ref = fuse_ref.segment_a_video_with_fusion(video, oracle_model, interpolate_last=True, step=1, num_clips=4)
    print(f"  oracle F1 {time.time() - t0:.1f}s")
    got = fuse_utils.segment_a_video_with_fusion(video, net, interpolate_last=True, step=1, num_clips=4)
    print("  F1 exact fusion: shape", got.shape, got.dtype, "mismatching pixels", int((got != ref).sum()), "of", ref.size,
          " LV frac", float(ref.mean()))
    # F2
    starts = list(range(0, 70 - 32 + 1))
    probs, mots = [], []
    for s in starts:
        seg, mot = model_ref.forward(sd, torch.from_numpy(video[:, s:s + 32]).unsqueeze(0))
        probs.append(torch.softmax(seg, 1)); mots.append(mot)
    acc, cnt, mask = fuse_ref.warp_fuse(torch.cat(probs), torch.cat(mots), starts, 70)
    got2, det = fuse_utils.segment_a_video_with_fusion(video, net, fuse_method="warp", return_details=True)
    print("  F2 warp fusion: mismatching pixels", int((got2 != mask.numpy()).sum()), "of", mask.numel(), " LV frac", float(mask.float().mean()))
    _stats("F2 acc", det["acc"], acc.float())
    return 0


def main():
    if len(sys.argv) > 1:
        import torch
        print(f"== {sys.argv[1]} on {torch.cuda.get_device_name(0)}", flush=True)
        return globals()["section_" + sys.argv[1]]()
    rc = 0
    for s in SECTIONS:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), s], timeout=420).returncode
        except subprocess.TimeoutExpired:
            r = -9
            print(f"== {s}: TIMEOUT")
        print(f"== {s}: exit {r} in {time.time() - t0:.0f}s", flush=True)
        rc = rc or r
    return rc


if __name__ == "__main__":
    sys.exit(main())
