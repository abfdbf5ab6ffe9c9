#!/usr/bin/env python
"""Per-call wall times of the public e2e path (pinned host video in, int64 host mask out): looks for host-side jitter."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clasfv_b200 import synthetic
from clasfv_b200.src import fuse_utils
from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet

net = R2plus1D_18_MotionNet(pretrained=False, precision="bf16")
net.load_state_dict(synthetic.random_state_dict(0))
net = net.cuda().eval()
video_host = torch.from_numpy(synthetic.synthetic_echo_video(200, 112, 112, seed=0)).pin_memory()
ts = []
for i in range(24):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = fuse_utils.segment_a_video_with_fusion(video_host, net, fuse_method="warp", batch_clips=64)
    ts.append((time.perf_counter() - t0) * 1e3)
print("ms per call:", " ".join(f"{t:.1f}" for t in ts))
