#!/usr/bin/env python
"""Run a few bf16 forwards of one clip batch (for ncu captures of individual kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from clasfv_b200 import synthetic  # noqa: E402
from clasfv_b200._lib import OUT_LVPROB  # noqa: E402
from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
net = R2plus1D_18_MotionNet(pretrained=False, precision="bf16")
net.load_state_dict(synthetic.random_state_dict(0))
net = net.cuda().eval()
eng = net.engine()
if len(sys.argv) > 3:
    eng.set_option("dense_video", int(sys.argv[3]))
    eng.set_option("sub_batch", n)
video = torch.from_numpy(synthetic.synthetic_echo_video(32 + n - 1, 112, 112, seed=0)).cuda()
prob = torch.empty((n, 1, 32, 112, 112), dtype=torch.bfloat16, device="cuda")
mot = torch.empty((n, 4, 32, 112, 112), dtype=torch.bfloat16, device="cuda")
for _ in range(iters):
    eng.forward_into(video, prob, mot, OUT_LVPROB, clip_starts=list(range(n)), clip_len=32)
torch.cuda.synchronize()
print("ok", float(prob.float().mean()))
