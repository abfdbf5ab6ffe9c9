#!/usr/bin/env python
"""Per-convolution timing of one full-video forward (CLASFV_CONV_TRACE=1, csrc/api.cu:run_conv): aggregates the lines the
library prints into a table by layer geometry.  Development aid; every launch is timed alone with the stream drained."""
import os, sys, subprocess, re, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("CLASFV_CONV_TRACE") is None:
    env = dict(os.environ, CLASFV_CONV_TRACE="1")
    r = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, capture_output=True, text=True)
    lines = [l for l in r.stderr.splitlines() if l.startswith("conv ")]
    n_per = len(lines) // 2                       # two forwards: keep the second (warm)
    lines = lines[n_per:]
    agg = collections.OrderedDict()
    for l in lines:
        key = l.split("  ")[0]
        ms = float(re.search(r"([\d.]+) ms", l).group(1)); gf = float(re.search(r"([\d.]+) GFLOP", l).group(1))
        a = agg.setdefault(key, [0, 0.0, 0.0]); a[0] += 1; a[1] += ms; a[2] += gf
    tot = sum(a[1] for a in agg.values())
    print(f"{len(lines)} launches, {tot:.3f} ms summed")
    for k, (c, ms, gf) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{c:<3d} {gf / ms:6.0f} TFLOP/s  {k}")
    print(r.stdout[-300:], r.stderr[-300:] if not lines else "")
    sys.exit(0)
sys.path.insert(0, ROOT)
import torch
from clasfv_b200 import synthetic
from clasfv_b200._lib import OUT_LVPROB
from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
tv = int(sys.argv[1]) if len(sys.argv) > 1 else 200
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
net = R2plus1D_18_MotionNet(pretrained=False, precision=prec); net.load_state_dict(synthetic.random_state_dict(0)); net = net.cuda().eval()
eng = net.engine(); eng.set_option("sub_batch", 64)
video = torch.from_numpy(synthetic.synthetic_echo_video(tv, 112, 112, seed=0)).cuda()
n = tv - 31
dt = torch.bfloat16 if prec == "bf16" else torch.float16
prob = torch.empty((n, 1, 32, 112, 112), dtype=dt, device="cuda"); mot = torch.empty((n, 4, 32, 112, 112), dtype=dt, device="cuda")
for _ in range(2):
    eng.forward_into(video, prob, mot, OUT_LVPROB, clip_starts=list(range(n)), clip_len=32)
torch.cuda.synchronize()
print("ok")
