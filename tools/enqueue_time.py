#!/usr/bin/env python
"""Host time to ENQUEUE one configs[1] video (forward of 169 windows + fusion) vs the GPU time it takes to run."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clasfv_b200 import synthetic
from clasfv_b200._lib import OUT_LVPROB
from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
net = R2plus1D_18_MotionNet(pretrained=False, precision="bf16"); net.load_state_dict(synthetic.random_state_dict(0)); net = net.cuda().eval()
eng = net.engine()
video = torch.from_numpy(synthetic.synthetic_echo_video(200, 112, 112, seed=0)).cuda()
n = 169; starts = list(range(n))
prob = torch.empty((n, 1, 32, 112, 112), dtype=torch.bfloat16, device="cuda"); mot = torch.empty((n, 4, 32, 112, 112), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    eng.forward_windows(video, prob, mot, OUT_LVPROB, starts, 32, 192); r = eng.warp_fuse(prob, mot, starts, 200)
torch.cuda.synchronize()
cpu, gpu = [], []
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.forward_windows(video, prob, mot, OUT_LVPROB, starts, 32, 192); r = eng.warp_fuse(prob, mot, starts, 200)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    cpu.append((t1 - t0) * 1e3); gpu.append((t2 - t0) * 1e3)
print("enqueue ms", sorted(cpu)[5], "total ms", sorted(gpu)[5], "launches", eng.launch_count() // 13)
