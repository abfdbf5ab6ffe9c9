#!/bin/bash
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "cta_pair or dense_video" 2>&1 | grep -E "passed|failed|FAILED" | tail -3
for V in A CLASFV_UMMA_NO_PAIR; do
env $V=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_${V}_$TAG.json 2>gpurun_out/bench_${V}_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${V}_$TAG.json").read().strip().splitlines()[-1])
print("$V", round(d["value"]), d.get("stage_ms_per_step"), "roofline", d["roofline"]["frac"], "e2e", round(d["e2e"]["value"]), d["clocks"])
PY
done
timeout 200 python tools/conv_trace.py 200 bf16 > gpurun_out/conv_trace_pair_$TAG.txt 2>&1; head -1 gpurun_out/conv_trace_pair_$TAG.txt
