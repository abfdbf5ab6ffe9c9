"""Host-side multi-GPU logic on CPU: world_size 2 and 3 over gloo.  Every rank fuses its clip range with the
oracle's warp_fuse (CPU tensors), exchanges partial sums through clasfv_b200.sharding.exchange_partials, and
the stitched result must equal fusing all clips in one process."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clasfv_b200 import sharding
from oracle import fuse_ref


def _worker(rank, world, port, n_frames, step, edge, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        h, w = 8, 12
        starts = sharding.clip_starts_for_video(n_frames, step)
        prob = torch.softmax(2 * torch.randn(len(starts), 2, 32, h, w, generator=g), 1)
        mot = torch.tanh(0.1 * torch.randn(len(starts), 4, 32, h, w, generator=g))
        ranges = sharding.partition_clips(len(starts), world)
        owners = sharding.frame_owners(starts, ranges, n_frames)
        c0, c1 = ranges[rank]
        mine = starts[c0:c1]
        lo, hi = sharding.touched_window(mine, n_frames, edge)
        if mine:
            acc, cnt, _ = fuse_ref.warp_fuse(prob[c0:c1], mot[c0:c1], [s - lo for s in mine], hi - lo, edge_hops=edge,
                                             accumulate=torch.float32)
            cnt = cnt.to(torch.int32)
        else:
            acc, cnt = torch.zeros(0, 2, h, w), torch.zeros(0, dtype=torch.int32)
        acc_own, cnt_own = sharding.exchange_partials(acc, cnt, (lo, hi), owners)
        f0, f1 = owners[rank]
        ref_acc, ref_cnt, _ = fuse_ref.warp_fuse(prob, mot, starts, n_frames, edge_hops=edge, accumulate=torch.float32)
        np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([
            float((acc_own - ref_acc[f0:f1]).abs().max()) if f1 > f0 else 0.0,
            float((cnt_own.long() - ref_cnt[f0:f1]).abs().max()) if f1 > f0 else 0.0, f0, f1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames,step,edge", [(2, 90, 1, False), (2, 70, 3, True), (3, 40, 1, True), (3, 64, 2, False)])
def test_clip_range_split_matches_single_process(tmp_path, world, n_frames, step, edge):
    port = 29500 + (os.getpid() + world * 7 + n_frames) % 2000
    mp.spawn(_worker, args=(world, port, n_frames, step, edge, str(tmp_path)), nprocs=world, join=True)
    covered = []
    for r in range(world):
        err_acc, err_cnt, f0, f1 = np.load(tmp_path / f"ok_{r}.npy")
        assert err_acc <= 1e-4 and err_cnt == 0, (r, err_acc, err_cnt)   # fp32 sums of up to 96 votes in another order
        covered.append((int(f0), int(f1)))
    assert covered[0][0] == 0 and covered[-1][1] == n_frames
    assert all(covered[i][1] == covered[i + 1][0] for i in range(world - 1))     # owned ranges tile [0, T)


def test_partition_and_video_sharding_bookkeeping():
    assert sharding.partition_clips(169, 8) == [(0, 22), (22, 43), (43, 64), (64, 85), (85, 106), (106, 127), (127, 148), (148, 169)]
    assert sharding.partition_clips(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    starts = sharding.clip_starts_for_video(2000, 1)
    assert len(starts) == 1969                                                    # BASELINE config 5
    owners = sharding.frame_owners(starts, sharding.partition_clips(len(starts), 8), 2000)
    assert owners[0][0] == 0 and owners[-1][1] == 2000 and all(a[1] == b[0] for a, b in zip(owners, owners[1:]))
    assert sharding.clip_starts_for_video(70, 3)[-1] == 38                        # the tail is always covered
    rng = np.random.default_rng(0)
    lengths = list(np.clip(rng.normal(175, 55, 1277).astype(int), 64, 400))       # BASELINE config 4
    shards = [sharding.shard_videos(lengths, r, 8) for r in range(8)]
    assert sorted(i for s in shards for i in s) == list(range(1277))
    loads = [sum(lengths[i] - 31 for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(lengths)                                # balanced to within one video
