#!/bin/bash
# conv parity + schedule identity, one bench run, one per-convolution trace
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "cta_pair or tcgen05 or dense_video or config1" 2>&1 | grep -E "passed|failed|FAILED" | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print(round(d["value"]), d.get("stage_ms_per_step"), "roofline", d["roofline"]["frac"], "e2e", round(d["e2e"]["value"]), d["clocks"])
PY
timeout 200 python tools/conv_trace.py 200 bf16 > gpurun_out/conv_trace_$TAG.txt 2>&1; head -1 gpurun_out/conv_trace_$TAG.txt
