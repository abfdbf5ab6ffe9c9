#!/bin/bash
# conv unit checks first (fast fail), then tests, bench, warp_fuse sweep, ncu launch list
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 300 python tests/gpu_diag.py conv_umma > gpurun_out/diag_conv_$TAG.log 2>&1; echo "diag exit $?" >> gpurun_out/diag_conv_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> gpurun_out/bench_$TAG.err
timeout 600 python tools/bench_warp_fuse.py > gpurun_out/warp_fuse_sweep_$TAG.jsonl 2>&1
if [ -z "$SKIP_NCU" ]; then
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 480 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_$TAG.log
fi
tail -n 18 gpurun_out/diag_conv_$TAG.log; grep -E "passed|failed|FAILED|^\[" gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/warp_fuse_sweep_$TAG.jsonl | cut -c1-200
