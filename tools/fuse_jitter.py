#!/usr/bin/env python
"""Hunts the sporadic ~20-60 ms stall seen in front of the fusion kernel: times every fusion call of many steps with
CUDA events, in three variants (fusion alone / after a forward / after a forward with pre-allocated outputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clasfv_b200 import synthetic
from clasfv_b200._lib import OUT_PROB
from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet

net = R2plus1D_18_MotionNet(pretrained=False, precision="bf16")
net.load_state_dict(synthetic.random_state_dict(0))
net = net.cuda().eval()
eng = net.engine()
T, H, W, CLIP = 200, 112, 112, 32
n = T - CLIP + 1
video = torch.from_numpy(synthetic.synthetic_echo_video(T, H, W, seed=0)).cuda()
prob = torch.empty((n, 2, CLIP, H, W), dtype=torch.bfloat16, device="cuda")
mot = torch.empty((n, 4, CLIP, H, W), dtype=torch.bfloat16, device="cuda")
starts = list(range(n))
eng.forward_windows(video, prob, mot, OUT_PROB, starts, CLIP, 64)
acc = torch.empty((T, 2, H, W), dtype=torch.float32, device="cuda")
cnt = torch.zeros((T,), dtype=torch.int32, device="cuda")

def run(label, steps, with_forward, prealloc, sync_each):
    evs = []
    for _ in range(steps):
        if with_forward:
            eng.forward_windows(video, prob, mot, OUT_PROB, starts, CLIP, 64)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if prealloc:
            eng.warp_fuse(prob, mot, starts, T, acc=acc, cnt=cnt, want_mask=False, want_area=False)
        else:
            eng.warp_fuse(prob, mot, starts, T)
        b.record()
        evs.append((a, b))
        if sync_each:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    ts = [x.elapsed_time(y) for x, y in evs]
    out = [(i, round(t, 2)) for i, t in enumerate(ts) if t > 2.0]
    print(f"{label}: steps {steps} median {sorted(ts)[len(ts)//2]:.3f} ms outliers(>2ms) {out}", flush=True)

run("fusion alone", 300, False, False, False)
run("forward+fusion", 60, True, False, False)
run("forward+fusion prealloc no mask/area", 60, True, True, False)
run("forward+fusion, sync each step", 60, True, False, True)
run("forward+fusion (again)", 60, True, False, False)
