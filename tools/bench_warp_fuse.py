#!/usr/bin/env python
"""BASELINE.json configs[2]: the warp+fusion kernel alone, HBM-bandwidth sweep.

256 stride-1 clips x 32 frames x 112x112 LV softmax + fwd/bwd flow fields, sweeping the clip count, the element type,
the frame size and the flow magnitude (the gather pattern).  Prints one JSON line per point: algorithmic bytes (the LV
plane + 4 flow planes per clip-frame read once + the fused outputs written once; `GBps_6plane` counts the background plane
the operator no longer reads, i.e. SURVEY 8(d)'s figure) / CUDA-event time.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from clasfv_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, nargs="*", default=[16, 64, 256])
    ap.add_argument("--flow-px", type=float, nargs="*", default=[0.0, 1.0, 4.0, 8.0])
    ap.add_argument("--dtypes", nargs="*", default=["fp32", "bf16", "fp16"])
    ap.add_argument("--size", type=int, default=112)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--flow-kinds", nargs="*", default=["white", "smooth"],
                    help="white: independent flow per pixel (the worst gather pattern: neighbouring pixels sample unrelated places); "
                         "smooth: a 14x14 field upsampled bilinearly, like the decoder's (neighbouring pixels sample neighbouring places)")
    ap.add_argument("--once", action="store_true", help="one point per dtype (256 clips, 4 px), 1 warm-up + 1 launch: for ncu captures")
    args = ap.parse_args()
    warm = 3
    if args.once:
        args.clips, args.flow_px, args.iters, warm, args.flow_kinds = [256], [4.0], 1, 1, ["white"]
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    eng = Engine("cuda:0")
    h = w = args.size
    g = torch.Generator(device="cuda").manual_seed(0)
    for dt in args.dtypes:
        dtype = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[dt]
        for n in args.clips:
            prob = torch.sigmoid(torch.randn(n, 1, 32, h, w, generator=g, device="cuda")).to(dtype)      # the LV probability
            for px, kind in [(px, kind) for px in args.flow_px for kind in (args.flow_kinds if px > 0 else args.flow_kinds[:1])]:
                # tanh(N(0, sigma)) with sigma chosen so that the rms displacement is `px` pixels
                if kind == "white":
                    mot = torch.tanh(torch.randn(n, 4, 32, h, w, generator=g, device="cuda") * (px / (h / 2.0))).to(dtype)
                else:
                    low = torch.randn(n * 4, 32, 14, 14, generator=g, device="cuda")
                    up = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=True).view(n, 4, 32, h, w)
                    up = up / up.std()
                    mot = torch.tanh(up * (px / (h / 2.0))).to(dtype)
                    del low, up
                starts = list(range(n))
                t_out = n + 31
                for _ in range(warm):
                    eng.warp_fuse(prob, mot, starts, t_out)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    eng.warp_fuse(prob, mot, starts, t_out)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                nbytes = n * 32 * h * w * 5 * prob.element_size() + t_out * h * w * 9 + t_out * 8
                nbytes6 = nbytes + n * 32 * h * w * prob.element_size()
                print(json.dumps({"kernel": "warp_fuse", "dtype": dt, "size": h, "clips": n, "flow_px_rms": px, "flow_kind": kind if px > 0 else "zero", "ms": round(ms, 4),
                                  "algorithmic_bytes": nbytes, "GBps": round(nbytes / ms / 1e6, 1), "frac_of_measured_hbm": round(nbytes / ms / 1e6 / peak, 4),
                                  "GBps_6plane": round(nbytes6 / ms / 1e6, 1),
                                  "note": "inputs of 16 clips (39-77 MB) fit the 126 MB L2" if n <= 16 else ""}), flush=True)
                del mot
            del prob


if __name__ == "__main__":
    main()
