#!/usr/bin/env python
"""BASELINE config 4: many videos of EchoNet-like lengths, sharded by video across the ranks of a torchrun job.

    python tools/config4_many_videos.py [--videos 1277] [--assign lpt|round_robin] [--precision bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/config4_many_videos.py ...

1 277 synthetic 112x112 videos, lengths ~ N(175, 55) clipped to [64, 400], seed 0.  Every video goes through
the public call ``segment_a_video_with_fusion(video, model, fuse_method="warp")``: pinned host float video in, int64 host
mask out, dense stride-1 clips - the timed region of a video is that call, host<->device copies included.  Videos are
independent units: no collective on the data path; the ranks only all-reduce their counters for the report.  Rank 0 prints
one JSON line: frames/s = all frames of all ranks / the slowest rank's summed call time.

Frame content: every video is a window of one 464-frame synthetic echo sequence (start offset varies with the video
index), copied into a pinned staging buffer before the timed call - filling that buffer is synthetic-data generation, not
part of the path, and is reported separately as ``host_fill_s``.
"""
import argparse
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

H = W = 112
MAX_LEN, MIN_LEN = 400, 64


def config4_lengths(n_videos=1277, seed=0):
    """Video lengths of BASELINE config 4: N(175, 55) rounded, clipped to [64, 400]."""
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(rng.normal(175.0, 55.0, size=n_videos)), MIN_LEN, MAX_LEN).astype(np.int64)


def assign(lengths, rank, world, how):
    if how == "round_robin":
        return list(range(rank, len(lengths), world))
    return sharding.shard_videos([int(x) for x in lengths], rank, world)


def rank_loads(lengths, world, how):
    """Clips every rank runs under an assignment (host arithmetic; used for the imbalance figure)."""
    return [int(sum(max(1, int(lengths[i]) - 31) for i in assign(lengths, r, world, how))) for r in range(world)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=1277)
    ap.add_argument("--assign", default="lpt", choices=["lpt", "round_robin"])
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--check-every", type=int, default=0, help="compare every k-th video of rank 0 with a second run of the same call")
    ap.add_argument("--single-calls", action="store_true", help="one synchronous segment_a_video_with_fusion call per video (round-1 mode) "
                    "instead of the pipelined many-video API")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    net = R2plus1D_18_MotionNet(pretrained=False, precision=args.precision)
    net.load_state_dict(synthetic.random_state_dict(0))
    net = net.cuda().eval()

    lengths = config4_lengths(args.videos)
    mine = assign(lengths, rank, world, args.assign)
    base = torch.from_numpy(synthetic.synthetic_echo_video(MAX_LEN + 64, H, W, seed=7))          # (3, 464, H, W)
    stage = torch.empty(3 * MAX_LEN * H * W, dtype=torch.float32).pin_memory()

    def staged(i):
        n = int(lengths[i])
        v = stage[: 3 * n * H * W].view(3, n, H, W)
        o = i % 64
        v.copy_(base[:, o:o + n])
        return v

    # warm-up on the longest and the shortest video: library workspace and torch's caching allocator reach their
    # high-water marks outside the timed calls (a cudaMalloc inside a call stalls the stream for tens of ms)
    for n in (MAX_LEN, MAX_LEN, MIN_LEN, 200):
        v = stage[: 3 * n * H * W].view(3, n, H, W)
        v.copy_(base[:, :n])
        keep = fuse_utils.segment_a_video_with_fusion(v, net, fuse_method="warp")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    if not args.single_calls:
        return pipelined(args, net, lengths, mine, base, rank, world)
    import gc
    gc.collect()
    gc.disable()
    call_s, fill_s, frames, clips, lv_pixels, slowest = 0.0, 0.0, 0, 0, 0, (0.0, 0)
    prev = None
    wall0 = time.perf_counter()
    for k, i in enumerate(mine):
        t0 = time.perf_counter()
        v = staged(i)
        t1 = time.perf_counter()
        mask = fuse_utils.segment_a_video_with_fusion(v, net, fuse_method="warp")     # returns after its own stream sync
        t2 = time.perf_counter()
        fill_s += t1 - t0
        call_s += t2 - t1
        n = int(lengths[i])
        if mask.shape != (n, H, W) or mask.dtype != np.int64:
            raise SystemExit(f"video {i}: mask {mask.shape} {mask.dtype}")
        frames += n
        clips += n - 31
        lv_pixels += int(mask.sum())
        per_frame = (t2 - t1) / n
        if per_frame > slowest[0]:
            slowest = (per_frame, n)
        if args.check_every and rank == 0 and k % args.check_every == 0:
            again = fuse_utils.segment_a_video_with_fusion(v, net, fuse_method="warp")
            if not np.array_equal(again, mask):
                raise SystemExit(f"video {i}: second run of the same call differs")
        prev = mask                                                                    # previous result stays alive, as a caller's would
    wall = time.perf_counter() - wall0
    gc.enable()
    del prev, keep

    tot = torch.tensor([call_s, wall, fill_s], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([frames, clips, lv_pixels, len(mine)], dtype=torch.int64, device="cuda")
    mx, mn = tot.clone(), tot.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        loads = rank_loads(lengths, world, args.assign)
        other = "round_robin" if args.assign == "lpt" else "lpt"
        loads_other = rank_loads(lengths, world, other)
        print(json.dumps({
            "check": "config4_many_videos", "ranks": world, "videos": int(cnt[3]), "frames": int(cnt[0]), "clips": int(cnt[1]),
            "height": H, "width": W, "precision": args.precision, "assign": args.assign,
            "lengths": {"min": int(lengths.min()), "mean": float(lengths.mean()), "max": int(lengths.max())},
            "frames_per_s": int(cnt[0]) / float(mx[0]),
            "frames_per_s_wall_incl_synthetic_fill": int(cnt[0]) / float(mx[1]),
            "call_s_slowest_rank": float(mx[0]), "call_s_fastest_rank": float(mn[0]), "host_fill_s_slowest_rank": float(mx[2]),
            "clips_per_rank_max_over_mean": max(loads) / (sum(loads) / world),
            f"clips_per_rank_max_over_mean_{other}": max(loads_other) / (sum(loads_other) / world),
            "slowest_video_rank0": {"ms_per_frame": slowest[0] * 1e3, "frames": slowest[1]},
            "lv_pixel_fraction": int(cnt[2]) / (int(cnt[0]) * H * W),
            "h2d_bytes": int(cnt[0]) * 3 * H * W * 4, "d2h_bytes": int(cnt[0]) * H * W * 8,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def pipelined(args, net, lengths, mine, base, rank, world):
    """All videos of this rank through fuse_utils.segment_videos_with_fusion: host NumPy (3,T,H,W) float32 videos in (windows of
    the synthetic sequence: strided views, made contiguous by the API's own staging), host int64 masks out; the timed region is
    the whole loop, host staging included."""
    base_np = base.numpy()

    def videos(idx):
        for i in idx:
            o, n = i % 64, int(lengths[i])
            yield base_np[:, o:o + n]

    for _m in fuse_utils.segment_videos_with_fusion(videos(mine[:4]), net):       # pipeline warm-up: staging slots, copy streams
        pass
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    frames = clips = lv_pixels = 0
    t0 = time.perf_counter()
    for i, mask in zip(mine, fuse_utils.segment_videos_with_fusion(videos(mine), net)):
        n = int(lengths[i])
        if mask.shape != (n, H, W) or mask.dtype != np.int64:
            raise SystemExit(f"video {i}: mask {mask.shape} {mask.dtype}")
        frames += n; clips += n - 31; lv_pixels += int(mask[::16].sum())
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    tot = torch.tensor([wall], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([frames, clips, len(mine)], dtype=torch.int64, device="cuda")
    mx, mn = tot.clone(), tot.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        loads = rank_loads(lengths, world, args.assign)
        print(json.dumps({
            "check": "config4_many_videos", "api": "fuse_utils.segment_videos_with_fusion (host NumPy in, host int64 masks out, pipelined)",
            "ranks": world, "videos": int(cnt[2]), "frames": int(cnt[0]), "clips": int(cnt[1]), "height": H, "width": W,
            "precision": args.precision, "assign": args.assign,
            "lengths": {"min": int(lengths.min()), "mean": float(lengths.mean()), "max": int(lengths.max())},
            "frames_per_s": int(cnt[0]) / float(mx[0]), "seconds_slowest_rank": float(mx[0]), "seconds_fastest_rank": float(mn[0]),
            "clips_per_rank_max_over_mean": max(loads) / (sum(loads) / world),
            "h2d_bytes": int(cnt[0]) * 3 * H * W * 4, "d2h_bytes": int(cnt[0]) * H * W * 8}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
