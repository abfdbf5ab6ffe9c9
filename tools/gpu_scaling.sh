#!/bin/bash
# N-GPU evidence for profiles/: the video-sharded bench (configs[1] per rank, configs[3] shape of work) and the long-video
# clip-range split (configs[4]) under torchrun on N GPUs of one box.
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --workload long-video --steps 3 --warmup 2 > gpurun_out/bench_long_${N}gpu_$TAG.json 2> gpurun_out/bench_long_${N}gpu_$TAG.err
else
  timeout 600 $TR --master-port 29533 bench.py --gpus $N --workload long-video --steps 3 --warmup 2 > gpurun_out/bench_long_${N}gpu_$TAG.json 2> gpurun_out/bench_long_${N}gpu_$TAG.err
  echo "long-video x$N exit $?"
  timeout 600 $TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
  echo "bench x$N exit $?"
  timeout 300 $TR --master-port 29535 tools/long_video_check.py 300 224 224 fp16 > gpurun_out/long_video_check_${N}gpu_$TAG.jsonl 2>&1
  tail -1 gpurun_out/long_video_check_${N}gpu_$TAG.jsonl
fi
python - <<PY
import json
for f in ("gpurun_out/bench_long_${N}gpu_$TAG.json", "gpurun_out/bench_${N}gpu_$TAG.json"):
    try:
        d = json.load(open(f)); print(f, d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d.get("e2e", {}).get("stage_seconds_max_over_ranks"))
    except Exception as e:
        print(f, "missing", e)
PY
