#!/usr/bin/env python
"""Dense-video vs per-clip schedule, bf16, at the benchmark geometry, with CTA pairs on and off: where do the outputs differ?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clasfv_b200.synthetic as synthetic
from clasfv_b200 import _lib
from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet

tv, clip_len, h, w, step, sub_batch = 52, 32, 112, 112, 1, 8
net = R2plus1D_18_MotionNet(pretrained=False, precision="bf16")
net.load_state_dict(synthetic.random_state_dict(0))
net = net.to("cuda:0").eval()
eng = net.engine()
video = torch.from_numpy(synthetic.synthetic_echo_video(tv, h, w, seed=5)).cuda()
starts = list(range(0, tv - clip_len + 1, step))
res = {}
for pair in (1, 0):
    for dense in (1, 0):
        eng.set_option("umma_pair", pair); eng.set_option("dense_video", dense); eng.set_option("sub_batch", sub_batch)
        seg, mot = eng.forward(video, _lib.OUT_LOGITS, torch.bfloat16, clip_starts=starts, clip_len=clip_len)
        torch.cuda.synchronize()
        res[(pair, dense)] = (seg.float().cpu(), mot.float().cpu())
for a, b in (((1, 1), (1, 0)), ((0, 1), (0, 0)), ((1, 1), (0, 1)), ((1, 0), (0, 0))):
    d = (res[a][0] - res[b][0]).abs()
    clips = sorted(set(torch.nonzero(d.flatten(1).amax(1) > 0).flatten().tolist()))
    frames = sorted(set(torch.nonzero(d.amax((0, 1, 3, 4)) > 0).flatten().tolist()))
    print(f"(pair,dense)={a} vs {b}: seg differing {int((d > 0).sum())} of {d.numel()}, max {float(d.max()):.4g}; clips {clips[:30]} frames {frames[:40]}", flush=True)
# repeatability of the pair kernel
eng.set_option("umma_pair", 1); eng.set_option("dense_video", 1)
seg2, _ = eng.forward(video, _lib.OUT_LOGITS, torch.bfloat16, clip_starts=starts, clip_len=clip_len)
print("pair dense run-to-run equal:", torch.equal(seg2.float().cpu(), res[(1, 1)][0]))
