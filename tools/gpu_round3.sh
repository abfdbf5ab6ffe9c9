#!/bin/bash
# dense-video schedule bring-up: the new parity tests first (fast fail), then the whole gpu suite, bench both schedules
TAG=${1:-r01k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -x -k "dense_video or tcgen05" > gpurun_out/pytest_dense_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dense_$TAG.log
tail -n 30 gpurun_out/pytest_dense_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --per-clip > gpurun_out/bench_perclip_$TAG.json 2> gpurun_out/bench_perclip_$TAG.err; echo "bench exit $?" >> gpurun_out/bench_perclip_$TAG.err
cat gpurun_out/bench_perclip_$TAG.json; tail -3 gpurun_out/bench_perclip_$TAG.err
if [ -z "$SKIP_FULL" ]; then
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|FAILED|^\[|Error" gpurun_out/pytest_gpu_$TAG.log | tail -20
fi
