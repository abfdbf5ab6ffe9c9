#!/usr/bin/env python
"""BASELINE config 5 check: one long video split by clip range across the ranks of a torchrun job, NCCL halo
exchange, compared on rank 0 with the single-GPU result; prints a JSON line with timing (max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/long_video_check.py [T H W]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from clasfv_b200 import sharding, synthetic  # noqa: E402
from clasfv_b200.src import fuse_utils  # noqa: E402
from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet  # noqa: E402

t, h, w = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (400, 112, 112)
precision = sys.argv[4] if len(sys.argv) > 4 else "bf16"
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
net = R2plus1D_18_MotionNet(pretrained=False, precision=precision)
net.load_state_dict(synthetic.random_state_dict(0))
net = net.cuda().eval()
video = synthetic.synthetic_echo_video(t, h, w, seed=3)
sharding.segment_long_video(video, net)                      # warm-up with the real shape (workspace, staging, NCCL channels)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
full = sharding.segment_long_video(video, net, mask_dtype=np.int64)
torch.cuda.synchronize(); dist.barrier()
dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    t1 = time.perf_counter()
    single = fuse_utils.segment_a_video_with_fusion(video, net, fuse_method="warp")
    torch.cuda.synchronize()
    t_single = time.perf_counter() - t1
    mism = int((single != full).sum())
    print(json.dumps({"check": "long_video_clip_range_split", "ranks": world, "frames": t, "height": h, "width": w, "precision": precision,
                      "mismatching_pixels_vs_single_gpu": mism, "pixels": int(full.size), "seconds_split": float(dt), "seconds_single_gpu": t_single,
                      "frames_per_s_split": t / float(dt)}), flush=True)
dist.destroy_process_group()
