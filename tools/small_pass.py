#!/usr/bin/env python
"""A small pass over every kernel family (compute-sanitizer is closed on this GPU pool, so this is a plain smoke run): a 40-frame 32x32 video through both 16-bit modes'
full-video pipeline (dense schedule: time-segmented and frame-selected convs, patch-GEMM head, staged fusion), the fp32 mode,
the F1 fusion and the ingest.  Prints 'sanitize pass ok'."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clasfv_b200 import synthetic
from clasfv_b200.src import fuse_utils
from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
sd = synthetic.random_state_dict(0)
video = synthetic.synthetic_echo_video(40, 32, 32, seed=2)
for prec in ("fp16", "bf16", "fp32"):
    net = R2plus1D_18_MotionNet(pretrained=False, precision=prec); net.load_state_dict(sd); net = net.cuda().eval()
    m = fuse_utils.segment_a_video_with_fusion(video, net, fuse_method="warp")
    assert m.shape == (40, 32, 32)
    if prec != "bf16":
        m2 = fuse_utils.segment_a_video_with_fusion(video, net, step=1, num_clips=3)
        assert m2.shape == (40, 32, 32)
    seg, mot = net(torch.from_numpy(video[:, :16]).unsqueeze(0).cuda())
    assert torch.isfinite(seg).all()
frames = np.random.RandomState(0).randint(0, 256, size=(6, 40, 50, 3)).astype(np.uint8)
v = net.engine().ingest_u8(frames, 32, 32, bgr=True)
torch.cuda.synchronize()
print("sanitize pass ok")
