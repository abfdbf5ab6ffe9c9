#!/usr/bin/env python
"""bench.py - frames/sec segmented + tracked (fusion on) for the CLAS-FV full-video inference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

    python bench.py --workload long-video [--gpus N] ...      (BASELINE configs[4], see run_long_video)

One "step" = one whole pass of the hot path over one synthetic video of BASELINE.json configs[1]
(112x112, 200 frames, bf16): every stride-1 32-frame window (169 clips) through the R(2+1)D-18
encoder + decoder heads, then warp-and-fuse into one (200,112,112) mask.  With N > 1 every rank
processes its own video (videos shard across GPUs with no collective: weak scaling).

  value  fused frames/s with the video already resident in HBM (CUDA events, max over ranks)
  e2e    the same through the public API with HOST NumPy videos in and host int64 masks out, every step:
         fuse_utils.segment_videos_with_fusion (the many-video form of segment_a_video_with_fusion(video, model,
         fuse_method="warp"): pinned staging + copies overlapped with the neighbouring videos' compute); median of
         three K-step brackets.  e2e.single_call is the one-video-at-a-time synchronous call.
  fp16   the same two numbers in the fp16 tensor-core mode (same kernels; the 16-bit mode that meets the parity gates)
  roofline      trunk convolutions (tcgen05 implicit GEMM): algorithmic FLOP/s vs measured bf16 peak
  roofline_warp_fuse  the fusion kernel: algorithmic bytes/s vs measured HBM copy bandwidth
  cpu_baseline  the oracle (PyTorch CPU restatement of the reference path) on a bounded sample

--impl reference times that CPU path alone (the reference is PyTorch; /root/reference is not on the
GPU box, so the oracle's line-by-line restatement of it is what runs).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The contract is ONE JSON line on stdout.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
# version banner there when a torchrun job creates its first communicator), so fd 1 is pointed at stderr for the whole
# run and the JSON line goes to a private duplicate of the original stdout.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_VIDEO, H, W, CLIP = 200, 112, 112, 32
N_CLIPS = T_VIDEO - CLIP + 1                       # 169 stride-1 windows
# trunk MACs per 32x112x112 clip executed by the tensor-core kernel
# (encoder 81.038 GMAC minus the 1x7x7 stem conv 0.664 GMAC, which runs on CUDA cores)
TRUNK_GFLOP_PER_CLIP = 2 * (81.038 - 0.664)
ENCODER_GFLOP_PER_CLIP = 2 * 81.038
KERNELS_PER_FORWARD = 42                           # stem + 40 convolutions + head


TRAFFIC_FILE = "r02_dram_traffic_per_step.json"


def ncu_traffic():
    """DRAM bytes per bench step and kernel, from the committed ncu launch list of this workload and these kernels
    (profiles/r02_dram_traffic_per_step.json, written by tools/launch_list.sh + tools/launch_list_aggregate.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)) as f:
            return json.load(f)["per_kernel"]
    except Exception:
        return {}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "bf16_tflops_burst": float(p["bf16_tflops"]), "source": "MEASURED_PEAKS.json"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nme in enumerate(names):
                if len(r) > 2 + i and r[2 + i].lower().startswith("active"):
                    reasons.add(nme)
        smax = next((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(sample_clips, threads=None):
    """Oracle = PyTorch-CPU restatement of the reference path, on `sample_clips` stride-1 windows."""
    import torch
    from clasfv_b200 import synthetic
    from oracle import fuse_ref, model_ref
    # every host core this process may run on - torchrun exports OMP_NUM_THREADS=1, which would otherwise leave the
    # reference arm of an N>1 launch on a single thread
    torch.set_num_threads(threads or len(os.sched_getaffinity(0)))
    sd = synthetic.random_state_dict(0)
    video = synthetic.synthetic_echo_video(CLIP + sample_clips - 1, H, W, seed=0)
    starts = list(range(sample_clips))
    t0 = time.perf_counter()
    probs, mots = [], []
    for s in starts:
        seg, mot = model_ref.forward(sd, torch.from_numpy(video[:, s:s + CLIP]).unsqueeze(0))
        probs.append(torch.softmax(seg, 1)); mots.append(mot)
    t_model = time.perf_counter() - t0
    t0 = time.perf_counter()
    fuse_ref.warp_fuse(torch.cat(probs), torch.cat(mots), starts, CLIP + sample_clips - 1, accumulate=torch.float32)
    t_fuse = time.perf_counter() - t0
    # the whole workload = 169 clips: model cost scales with clips, fusion with clip-frames
    sec_per_video = (t_model + t_fuse) / sample_clips * N_CLIPS
    return {"frames_per_s": T_VIDEO / sec_per_video, "sec_per_clip_model": t_model / sample_clips, "sec_per_clip_fuse": t_fuse / sample_clips,
            "threads": torch.get_num_threads()}


def run_reference_arm(args, rank):
    if rank != 0:
        return 0
    import torch
    sample = 2
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_sample(1)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_reference_sample(sample))
    wall = time.perf_counter() - t0
    fps = sum(v["frames_per_s"] for v in vals) / len(vals)
    line = {
        "impl": "reference", "metric": "frames/sec segmented+tracked (fusion on)", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "value_is": "EXTRAPOLATED: seconds per clip of the sample x 169 clips (a whole video is 2 CPU-minutes per step)",
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": vals[0]["threads"], "kind": "port",
                         "sample": f"{sample} of {N_CLIPS} stride-1 clips per step (reference network on PyTorch CPU + oracle warp-fuse), "
                                   f"scaled to the 169-clip video; {wall:.1f}s of CPU work"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    return 0


def workload_config():
    return {"workload": "configs[1]: full-video fusion, every stride-1 32-frame clip of one 112x112 200-frame video "
                        "(169 clips -> 200 fused frames), warp-and-fuse on, bf16 tensor-core mode",
            "frames": T_VIDEO, "height": H, "width": W, "clips_per_video": N_CLIPS, "clip_len": CLIP, "fusion": "warp",
            "weights": "random-init R2plus1D_18_MotionNet (seed 0), 31,575,731 parameters",
            "l2": "no explicit flush: per-step activation traffic (~5 GB) is far larger than the 126 MB L2"}


# ------------------------------------------------------------------------------------------------ GPU arm
def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def measure_precision(args, precision, dev, rank, world, barrier, full):
    """value (video resident, CUDA events) and e2e (public API, NumPy in / NumPy out) of one precision mode.  `full`: also the
    stage profile, clocks and launch count (the primary mode)."""
    import gc
    import numpy as np
    import torch
    from clasfv_b200 import synthetic
    from clasfv_b200._lib import OUT_LVPROB
    from clasfv_b200.engine import storage_dtype
    from clasfv_b200.src import fuse_utils
    from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet

    net = R2plus1D_18_MotionNet(pretrained=False, precision=precision)
    net.load_state_dict(synthetic.random_state_dict(0))
    net = net.to(dev).eval()
    eng = net.engine()
    out_dtype = storage_dtype(precision)
    videos_np = [synthetic.synthetic_echo_video(T_VIDEO, H, W, seed=2 * rank + i) for i in range(2)]     # host NumPy fp32 (3,T,H,W)
    video = torch.from_numpy(videos_np[0]).to(dev)
    starts = list(range(N_CLIPS))
    prob = torch.empty((N_CLIPS, 1, CLIP, H, W), dtype=out_dtype, device=dev)      # LV probability: all that fusion reads
    mot = torch.empty((N_CLIPS, 4, CLIP, H, W), dtype=out_dtype, device=dev)
    bc = args.batch_clips
    eng.set_option("dense_video", 0 if args.per_clip else 1)

    # warm-up keeps the previous step's result alive while the next one is produced, exactly like the timed loop below:
    # otherwise the second set of output buffers is cudaMalloc'ed by torch's caching allocator inside the SECOND TIMED step
    res = None
    for _ in range(args.warmup):
        eng.forward_windows(video, prob, mot, OUT_LVPROB, starts, CLIP, bc)
        res = eng.warp_fuse(prob, mot, starts, T_VIDEO)
    barrier()
    gc.collect()
    gc.disable()                          # Python's cyclic garbage collector is off inside the timed regions (as timeit does)
    sampler = ClockSampler(dev.index)
    if full:
        sampler.start()
        eng.profile_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fuse_events = []
    launches0 = eng.launch_count()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        eng.forward_windows(video, prob, mot, OUT_LVPROB, starts, CLIP, bc)
        fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fa.record()
        res = eng.warp_fuse(prob, mot, starts, T_VIDEO)
        fb.record()
        fuse_events.append((fa, fb))
    ev1.record()
    barrier()
    out = {"ms_total": ev0.elapsed_time(ev1), "launches": eng.launch_count() - launches0}
    if full:
        out["clocks"] = sampler.stop()
        out["stage_ms"], _calls = eng.profile_end()
        out["stage_gflop"] = eng.last_profile_gflop
    out["fuse_ms_each"] = [a.elapsed_time(b) for a, b in fuse_events]
    out["flow_px"] = float((mot.float().abs().mean() * (W / 2)).item())      # what the gather pattern of warp_fuse depends on
    del res

    # ---- e2e through the public API: host NumPy videos in, host int64 masks out, every step
    def stream_of_videos(k):
        return (videos_np[i % 2] for i in range(k))

    def run_pipelined(k):
        masks = 0
        for m in fuse_utils.segment_videos_with_fusion(stream_of_videos(k), net, batch_clips=bc):
            assert m.shape == (T_VIDEO, H, W) and m.dtype == np.int64
            masks += 1
        assert masks == k

    def run_single(k):
        for v in stream_of_videos(k):
            m = fuse_utils.segment_a_video_with_fusion(v, net, fuse_method="warp", batch_clips=bc)
        assert m.shape == (T_VIDEO, H, W) and m.dtype == np.int64

    for fn, key in ((run_pipelined, "e2e_brackets"), (run_single, "single_brackets")):
        fn(max(3, args.warmup))
        br = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            fn(args.steps)
            torch.cuda.synchronize()
            br.append(time.perf_counter() - t0)
        out[key] = br
    gc.enable()
    times = torch.tensor([out["ms_total"], median(out["e2e_brackets"]) * 1e3, median(out["single_brackets"]) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    out["ms_total"], out["e2e_ms"], out["single_ms"] = (float(x) for x in times)
    out["h2d_bytes"] = int(videos_np[0].nbytes)
    out["d2h_bytes"] = int(T_VIDEO * H * W * 8)
    return out


def e2e_block(m, args, world):
    fps = lambda ms: world * T_VIDEO * args.steps / (ms / 1e3)  # noqa: E731
    return {"value": fps(m["e2e_ms"]), "unit": "frames/s", "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": m["d2h_bytes"],
            "api": "fuse_utils.segment_videos_with_fusion(iterable of host NumPy (3,T,H,W) float32 videos, model) -> host int64 (T,H,W) masks; "
                   "every video is staged through pinned memory, uploaded, segmented and its mask copied back inside the timed region",
            "brackets_ms_per_step": [1e3 * t / args.steps for t in m["e2e_brackets"]], "reported": "median of three K-step brackets (max over ranks)",
            "single_call": {"value": fps(m["single_ms"]), "api": "fuse_utils.segment_a_video_with_fusion(numpy_video, model, fuse_method='warp'), one "
                            "synchronous call per step", "brackets_ms_per_step": [1e3 * t / args.steps for t in m["single_brackets"]]}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="video", choices=("video", "long-video"))
    ap.add_argument("--batch-clips", type=int, default=int(os.environ.get("CLASFV_BATCH_CLIPS", "192")))
    ap.add_argument("--precision", default="bf16", choices=("bf16", "fp16", "fp32"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the fp16 measurement")
    ap.add_argument("--per-clip", action="store_true", help="disable the dense-video schedule (every clip runs the whole trunk)")
    ap.add_argument("--frames", type=int, default=2000, help="long-video workload: frames")
    ap.add_argument("--size", type=int, default=224, help="long-video workload: frame height = width")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference_arm(args, rank)

    import torch
    import torch.distributed as dist
    import clasfv_b200  # noqa: F401

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device - clasfv_b200 has no CPU path (use --impl reference for the CPU arm)")
    if args.warmup < 3:
        print("bench.py: note - fewer than 3 warm-up steps requested", file=sys.stderr)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 or args.workload == "long-video":
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29631")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        else:
            dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "long-video":
        rc = run_long_video(args, dev, rank, world, barrier)
        dist.destroy_process_group()
        return rc

    m = measure_precision(args, args.precision, dev, rank, world, barrier, full=True)
    m16 = None
    if args.precision == "bf16" and not args.no_secondary:
        m16 = measure_precision(args, "fp16", dev, rank, world, barrier, full=False)

    if rank == 0:
        peaks = measured_peaks()
        bc = args.batch_clips
        ms_total, stage_ms, stage_gflop = m["ms_total"], m["stage_ms"], m["stage_gflop"]
        fuse_ms = sum(m["fuse_ms_each"]) / len(m["fuse_ms_each"])
        ms_per_step = ms_total / args.steps
        value = world * T_VIDEO * args.steps / (ms_total / 1e3)
        trunk_ms_per_step = stage_ms["trunk"] / args.steps
        # performed = MACs the tcgen05 kernel really executed (the dense-video schedule computes the stem and layer1
        # once for the frames overlapping windows share); algorithmic = SURVEY 8(d)'s per-clip figure x clips
        performed_tflops = stage_gflop["trunk"] / args.steps / trunk_ms_per_step          # GFLOP/ms == TFLOP/s
        algorithmic_tflops = TRUNK_GFLOP_PER_CLIP * N_CLIPS / trunk_ms_per_step
        tr = ncu_traffic() if (not args.per_clip and bc >= N_CLIPS and args.precision == "bf16") else {}
        conv_traffic = (tr["conv_umma_kernel"]["dram_read_bytes"] + tr["conv_umma_kernel"]["dram_write_bytes"]) if "conv_umma_kernel" in tr else None
        wf_key = next((k for k in tr if k.startswith("warp_fuse_staged")), None)
        fuse_traffic = (tr[wf_key]["dram_read_bytes"] + tr[wf_key]["dram_write_bytes"]) if wf_key else None
        elt = 4 if args.precision == "fp32" else 2
        # algorithmic bytes of F2: the LV plane + 4 flow planes of every clip frame in, class sums + mask out
        fuse_bytes = N_CLIPS * CLIP * H * W * 5 * elt + T_VIDEO * H * W * (8 + 1) + T_VIDEO * 8
        line = {
            "metric": "frames/sec segmented+tracked (fusion on)", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "fp16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": workload_config(),
            "run_config": {"batch_clips": bc, "parallelism": f"video-sharded x{world}, no collective"},
            "clip_frames_per_s": world * N_CLIPS * CLIP * args.steps / (ms_total / 1e3),
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()} | {"warp_fuse": fuse_ms},
            "warp_fuse_ms_each_step": [round(x, 3) for x in m["fuse_ms_each"]],
            "e2e": e2e_block(m, args, world),
            "gpu_launches": m["launches"],
            "gpu_launches_is": "kernels launched through the library handle inside the timed region of `value` (counted by the library: "
                               "clasfv_launch_count), %d per step" % (m["launches"] // max(1, args.steps)),
            "roofline": {"kernel": "conv_umma_kernel (trunk: stem 3x1x1 + layer1-4)",
                         "bound": "tensor", "achieved": performed_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": performed_tflops / peaks["bf16_tflops"], "traffic": conv_traffic,
                         "traffic_is": "DRAM read+write bytes of all conv_umma_kernel launches of one step (ncu launch list of this "
                                       "command and these kernels, profiles/" + TRAFFIC_FILE + "); null with --per-clip, another "
                                       "batch size or another precision",
                         "peak_source": peaks["source"] + " (sustained; burst %.1f)" % peaks["bf16_tflops_burst"],
                         "achieved_is": "performed FLOPs (2 x true MACs executed by the kernel, counted per launch by the library) / "
                                        "trunk time from CUDA events inside clasfv_forward",
                         "performed_gflop_per_step": stage_gflop["trunk"] / args.steps,
                         "algorithmic_gflop_per_clip": TRUNK_GFLOP_PER_CLIP,
                         "algorithmic_equivalent_tflops": algorithmic_tflops,
                         "schedule": "per-clip" if args.per_clip else
                                     "dense-video: stem+layer1 once per video frame, clip-edge frames per clip (bit-identical outputs)"},
            "roofline_warp_fuse": {"kernel": "warp_fuse_staged_kernel", "bound": "hbm", "achieved": fuse_bytes / (fuse_ms * 1e6),
                                   "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": fuse_bytes / (fuse_ms * 1e6) / peaks["hbm_gbs"],
                                   "traffic": fuse_traffic, "algorithmic_bytes": fuse_bytes},
            "clocks": m["clocks"], "mean_abs_flow_px": m["flow_px"],
        }
        if m16 is not None:
            v16 = world * T_VIDEO * args.steps / (m16["ms_total"] / 1e3)
            line["fp16"] = {"note": "the same workload in the fp16 tensor-core mode (same kernels; 11 significant bits instead of 8): the "
                                    "16-bit mode whose masks meet the north-star parity gates (tests/test_gpu_parity.py, DESIGN.md 5)",
                            "value": v16, "ms_per_step": m16["ms_total"] / args.steps, "e2e": e2e_block(m16, args, world)}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_sample(4)
            line["cpu_baseline"] = {"value": cb["frames_per_s"], "unit": "frames/s", "cores": cb["threads"], "kind": "port",
                                    "value_is": "EXTRAPOLATED: the sample's seconds per clip x 169 clips",
                                    "sample": "4 of 169 stride-1 clips (reference network restated on PyTorch CPU, "
                                              f"{cb['sec_per_clip_model']:.2f} s/clip) + oracle warp-fuse, scaled to the whole video"}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_long_video(args, dev, rank, world, barrier):
    """BASELINE configs[4]: ONE long video (default 2000 frames, 224 x 224: 1969 stride-1 clips) split by clip range across the
    ranks, neighbour halo exchange of the partial class sums over NCCL (sharding.segment_long_video).  A step is the whole
    video.  value: frames/s with the video resident on every GPU and the mask left on the device (gather=False);
    e2e: the public call with the HOST video in and the host uint8 mask out on rank 0.  Strong scaling."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from clasfv_b200 import sharding, synthetic
    from clasfv_b200.src.model.R2plus1D_18_MotionNet import R2plus1D_18_MotionNet
    t, size = args.frames, args.size
    net = R2plus1D_18_MotionNet(pretrained=False, precision=args.precision)
    net.load_state_dict(synthetic.random_state_dict(0))
    net = net.to(dev).eval()
    video = synthetic.synthetic_echo_video(t, size, size, seed=3)
    bc = min(args.batch_clips, 64 if size > 112 else 192)
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    video_dev = torch.from_numpy(video).to(dev)

    def timed(fn, k):
        ts = []
        for _ in range(k):
            barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        x = torch.tensor(ts, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return [float(v) for v in x]

    stage = {}
    resident = lambda: sharding.segment_long_video(video_dev, net, batch_clips=bc, gather=False)      # noqa: E731
    public = lambda: sharding.segment_long_video(video, net, batch_clips=bc, gather="rank0", timings=stage)  # noqa: E731
    timed(resident, warm)
    t_res = timed(resident, steps)
    timed(public, warm)
    stage.clear()
    t_pub = timed(public, steps)
    st = torch.tensor([stage.get(k, 0.0) / steps for k in ("upload", "forward", "fuse", "halo", "gather", "d2h")], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
    if rank == 0:
        n_clips = t - CLIP + 1
        line = {"metric": "frames/sec segmented+tracked (fusion on)", "value": t / median(t_res), "unit": "frames/s", "n_gpus": world,
                "steps": steps, "warmup": warm, "ms_per_step": 1e3 * median(t_res), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "fp16", "fp32": "f32"}[args.precision], "data": "synthetic",
                "config": {"workload": f"configs[4]: one {t}-frame {size}x{size} video, {n_clips} stride-1 clips split by clip range across "
                                       f"{world} rank(s), NCCL neighbour halo exchange of the partial class sums", "frames": t, "height": size,
                           "width": size, "clips": n_clips, "batch_clips": bc, "parallelism": f"clip-range split x{world}, P2P halo"},
                "e2e": {"value": t / median(t_pub), "unit": "frames/s", "h2d_bytes_per_step": int(video.nbytes), "d2h_bytes_per_step": int(t * size * size),
                        "api": "sharding.segment_long_video(host NumPy video, model, gather='rank0') -> host uint8 mask on rank 0",
                        "seconds_each": t_pub, "stage_seconds_max_over_ranks": dict(zip(("upload", "forward", "fuse", "halo", "gather", "d2h"), [float(x) for x in st]))},
                "seconds_each_resident": t_res}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
