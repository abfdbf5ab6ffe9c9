"""Multi-GPU partitioning of the full-video path.  One process per GPU (torchrun).

* Many videos: independent units - ``shard_videos`` assigns them to ranks, no collective on the data path.
* One long video: clips are independent, only fusion couples neighbours (output frame g receives votes from
  clips starting in [g-32, g+1]).  Each rank runs its contiguous range of clips, fuses them into *partial*
  class sums over the frames they touch, and ``exchange_partials`` adds every rank's overflow frames into
  the rank that owns them with point-to-point sends (NCCL over NVLink on GPUs; the same code runs on gloo
  with CPU tensors, which is how tests/ exercise it).  No all-reduce: payload is <= 33 frames per boundary.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

CLIP = 32


def shard_videos(lengths, rank, world):
    """Indices of the videos rank ``rank`` processes: longest-processing-time-first bin packing on the
    number of clips each video needs (videos differ in length; round-robin would leave ranks idle)."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    loads = [0] * world
    mine = []
    for i in order:
        r = int(np.argmin(loads))
        loads[r] += max(1, lengths[i] - CLIP + 1)
        if r == rank:
            mine.append(i)
    return sorted(mine)


def clip_starts_for_video(num_frames, step=1):
    starts = list(range(0, num_frames - CLIP + 1, step))
    if starts[-1] != num_frames - CLIP:
        starts.append(num_frames - CLIP)
    return starts


def partition_clips(n_clips, world):
    """Contiguous, balanced clip ranges [(c0, c1)] - one per rank (ranks may be empty when world > n_clips)."""
    base, extra = divmod(n_clips, world)
    out, c = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((c, c + n))
        c += n
    return out


def frame_owners(starts, ranges, num_frames):
    """Output frames [f0, f1) each rank finalises: from its first clip's first frame to the next non-empty
    rank's first frame; rank 0 starts at frame 0, the last non-empty rank ends at num_frames."""
    world = len(ranges)
    firsts = [starts[c0] if c1 > c0 else None for c0, c1 in ranges]
    owners, nxt = [None] * world, num_frames
    for r in range(world - 1, -1, -1):
        if firsts[r] is None:
            owners[r] = (nxt, nxt)
        else:
            owners[r] = (firsts[r], nxt)
            nxt = firsts[r]
    for r in range(world):                      # frames before the first clip belong to the first non-empty rank
        if firsts[r] is not None:
            owners[r] = (0, owners[r][1])
            break
        owners[r] = (0, 0)
    return owners


def touched_window(my_starts, num_frames, edge_hops=False):
    """Frames [lo, hi) the clips of one rank can vote on."""
    if not my_starts:
        return (0, 0)
    e = 1 if edge_hops else 0
    return (max(0, my_starts[0] - e), min(num_frames, my_starts[-1] + CLIP + e))


def exchange_partials(acc, cnt, window, owners, group=None, windows=None):
    """acc (hi-lo, 2, H, W), cnt (hi-lo,) = this rank's partial sums over frames window=[lo,hi).
    Returns (acc_owned, cnt_owned) over this rank's owned frames with every rank's votes added.
    ``windows`` = the [lo,hi) of every rank when the caller can derive them (the clip plan is deterministic), else they
    are all-gathered as 2 integers each; the frame data moves point-to-point between neighbours."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = window
    if windows is not None:
        wins = [(int(a), int(b)) for a, b in windows]
    else:
        wins = [None] * world
        dist.all_gather_object(wins, (int(lo), int(hi)), group=group)
    f0, f1 = owners[rank]
    h, w = acc.shape[-2:]
    acc_own = torch.zeros((f1 - f0, 2, h, w), dtype=acc.dtype, device=acc.device)
    cnt_own = torch.zeros((f1 - f0,), dtype=cnt.dtype, device=cnt.device)
    a, b = max(lo, f0), min(hi, f1)              # my own contribution
    if b > a:
        acc_own[a - f0:b - f0] += acc[a - lo:b - lo]
        cnt_own[a - f0:b - f0] += cnt[a - lo:b - lo]
    ops, recv_bufs = [], []
    for q in range(world):
        if q == rank:
            continue
        qf0, qf1 = owners[q]
        a, b = max(lo, qf0), min(hi, qf1)        # what I hold of q's frames
        if b > a:
            ops.append(dist.P2POp(dist.isend, acc[a - lo:b - lo].contiguous(), q, group))
            ops.append(dist.P2POp(dist.isend, cnt[a - lo:b - lo].contiguous(), q, group))
        qlo, qhi = wins[q]
        a, b = max(qlo, f0), min(qhi, f1)        # what q holds of my frames
        if b > a:
            ra = torch.empty((b - a, 2, h, w), dtype=acc.dtype, device=acc.device)
            rc = torch.empty((b - a,), dtype=cnt.dtype, device=cnt.device)
            ops.append(dist.P2POp(dist.irecv, ra, q, group))
            ops.append(dist.P2POp(dist.irecv, rc, q, group))
            recv_bufs.append((q, a, b, ra, rc))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for _q, a, b, ra, rc in sorted(recv_bufs, key=lambda x: x[0]):   # fixed (rank) order: deterministic sums
        acc_own[a - f0:b - f0] += ra
        cnt_own[a - f0:b - f0] += rc
    return acc_own, cnt_own


_stage = {}       # (device index, elements) -> pinned staging buffer for a rank's frame range
_copy_streams = {}


class _FrameRangeUpload:
    """video[:, fa:fb] of a host (3,T,H,W) array -> a contiguous fp32 CUDA tensor, chunk by chunk: a few host threads copy
    frame chunks into a reusable pinned buffer (a strided 600 MB copy on one thread runs at ~6 GB/s, and torchrun pins
    OMP_NUM_THREADS to 1), every chunk's transfer is enqueued on a copy stream as soon as it is staged, and
    ``ready(frame)`` returns the event after which all frames below ``frame`` are on the device - so the first clips of
    the range are segmented while the rest of the range is still in flight."""

    def __init__(self, video, fa, fb, dev):
        self.n = fb - fa
        shape = (3, self.n) + tuple(video.shape[2:])
        self.out = torch.empty(shape, dtype=torch.float32, device=dev)
        self.events = []                                         # (frames on the device so far, event)
        if isinstance(video, torch.Tensor) and video.is_cuda:
            self.out.copy_(video[:, fa:fb])
            self.pending = None
            return
        src = video if isinstance(video, torch.Tensor) else torch.from_numpy(video)
        key = (dev.index, int(np.prod(shape)))
        buf = _stage.get(key)
        if buf is None:
            _stage.clear()
            buf = _stage[key] = torch.empty(int(np.prod(shape)), dtype=torch.float32, pin_memory=True)
        self.host = buf.view(shape)
        self.stream = _copy_streams.setdefault(dev.index, torch.cuda.Stream(device=dev))
        self.stream.wait_stream(torch.cuda.current_stream(dev))      # the previous call's reads of the staging area are done
        chunk = max(8, -(-self.n // 16))
        pieces = [(a, min(self.n, a + chunk)) for a in range(0, self.n, chunk)]

        def stage(piece):
            a, b = piece
            self.host[:, a:b].copy_(src[:, fa + a:fa + b])
            return piece

        from concurrent.futures import ThreadPoolExecutor
        self.pool = ThreadPoolExecutor(max_workers=4)
        self.pending = [self.pool.submit(stage, p) for p in pieces]

    def ready(self, frame):
        """Event after which frames [0, frame) of the range are on the device (None: already there)."""
        if self.pending is None:
            return None
        while self.pending and (not self.events or self.events[-1][0] < frame):
            a, b = self.pending.pop(0).result()
            with torch.cuda.stream(self.stream):
                self.out[:, a:b].copy_(self.host[:, a:b], non_blocking=True)
                ev = torch.cuda.Event(); ev.record()
            self.events.append((b, ev))
        if not self.pending:
            self.pool.shutdown(wait=False)
        return next(ev for upto, ev in self.events if upto >= frame)


def segment_long_video(video, model, step=1, edge_hops=False, batch_clips=64, group=None, gather=True, mask_dtype=np.uint8,
                       timings=None):
    """One long video (3,T,H,W) split by clip range across the ranks of ``group`` (BASELINE config 5).
    Every rank passes the same host video.  ``gather``: True / "all" - the full (T,H,W) mask on every rank; "rank0" - on
    rank 0 only (None elsewhere); False - (owned mask, (f0, f1)).  Masks are ``mask_dtype`` (uint8 by default: a 2000-frame
    224 x 224 int64 mask is 800 MB of host traffic per rank; pass np.int64 for the reference's element type).
    ``timings``: optional dict that receives the seconds of each stage measured with CUDA events on this rank."""
    from . import engine as _engine
    from ._lib import OUT_LVPROB
    from .src.fuse_utils import _unwrap
    net = _unwrap(model)
    eng = net.engine()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if getattr(video, "ndim", None) != 4 or video.shape[0] != 3:
        raise ValueError(f"expected a video of shape (3,T,H,W), got {tuple(video.shape)}")
    num_frames, h, w = int(video.shape[1]), int(video.shape[2]), int(video.shape[3])
    starts = clip_starts_for_video(num_frames, step)
    ranges = partition_clips(len(starts), world)
    owners = frame_owners(starts, ranges, num_frames)
    c0, c1 = ranges[rank]
    mine = starts[c0:c1]
    lo, hi = touched_window(mine, num_frames, edge_hops)
    from .engine import storage_dtype
    out_dtype = storage_dtype(eng.precision)
    dev = eng.device
    cuda = dev.type == "cuda"
    marks = []

    def mark(name):
        if timings is not None and cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    if mine:
        # only the frames this rank's clips read go to its GPU (a 2000-frame 224x224 video is 1.2 GB as fp32)
        fa, fb = mine[0], mine[-1] + CLIP
        up = _FrameRangeUpload(video, fa, fb, dev)
        v = up.out
        prob = torch.empty((len(mine), 1, CLIP, h, w), dtype=out_dtype, device=dev)
        mot = torch.empty((len(mine), 4, CLIP, h, w), dtype=out_dtype, device=dev)
        # the range is segmented in a few sub-ranges of clips, each as soon as its frames have arrived
        n_sub = max(1, min(4, len(mine) // 48))
        bounds = [round(i * len(mine) / n_sub) for i in range(n_sub + 1)]
        main = torch.cuda.current_stream(dev) if cuda else None
        for a, b in zip(bounds, bounds[1:]):
            ev = up.ready(mine[b - 1] + CLIP - fa)
            if ev is not None:
                main.wait_event(ev)
            if a == 0:
                mark("upload")                   # what the first sub-range had to wait for; the rest overlaps the forward
            eng.forward_windows(v, prob[a:b], mot[a:b], OUT_LVPROB, [s - fa for s in mine[a:b]], CLIP, batch_clips)
        mark("forward")
        res = eng.warp_fuse(prob, mot, [s - lo for s in mine], hi - lo, edge_hops=edge_hops, want_mask=False, want_area=False)
        acc, cnt = res["acc"], res["cnt"]
        mark("fuse")
    else:
        acc = torch.zeros((0, 2, h, w), dtype=torch.float32, device=dev)
        cnt = torch.zeros((0,), dtype=torch.int32, device=dev)
    windows = [touched_window(starts[a:b], num_frames, edge_hops) for a, b in ranges]
    acc_own, cnt_own = exchange_partials(acc, cnt, (lo, hi), owners, group, windows=windows)
    mark("halo")
    f0, f1 = owners[rank]
    if f1 > f0:
        mask, _area = _engine.finalize_mask(acc_own.contiguous())
    else:
        mask = torch.zeros((0, h, w), dtype=torch.uint8, device=dev)

    def to_host(m):
        host = torch.empty(m.shape, dtype=torch.uint8, pin_memory=cuda)
        host.copy_(m, non_blocking=True)
        if cuda:
            torch.cuda.current_stream(dev).synchronize()
        out = host.numpy()
        return out if mask_dtype == np.uint8 else out.astype(mask_dtype)

    def finish(result):
        if timings is not None and cuda:
            mark("end")
            torch.cuda.current_stream(dev).synchronize()
            for (_n0, e0), (n1, e1) in zip(marks, marks[1:]):
                timings[n1] = timings.get(n1, 0.0) + e0.elapsed_time(e1) * 1e-3
        return result

    if not gather:
        out = to_host(mask)
        mark("d2h")
        return finish((out, (f0, f1)))
    # device-side gather of the owned uint8 masks (padded to the largest owned range), one host copy at the end
    most = max(b - a for a, b in owners)
    padded = torch.zeros((most, h, w), dtype=torch.uint8, device=dev)
    padded[:f1 - f0] = mask
    if gather == "rank0":
        parts = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
        dist.gather(padded, parts, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    else:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
    mark("gather")
    if parts is None:
        return finish(None)
    full = torch.cat([p[:b - a] for p, (a, b) in zip(parts, owners)], 0)
    assert full.shape[0] == num_frames
    out = to_host(full)
    mark("d2h")
    return finish(out)
