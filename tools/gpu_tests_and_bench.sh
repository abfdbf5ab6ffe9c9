#!/bin/bash
# head kernel bring-up: bf16 forward parity first, then bench
TAG=${1:-r01l}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -x -k "bf16 or dense_video" > gpurun_out/pytest_head_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_head_$TAG.log
tail -n 25 gpurun_out/pytest_head_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
if [ -z "$SKIP_FULL" ]; then
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|FAILED|^\[|Error" gpurun_out/pytest_gpu_$TAG.log | tail -20
fi
