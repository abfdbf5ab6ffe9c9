#!/bin/bash
# ncu launch list (durations only) of one bench step
TAG=${1:-r01}
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 700 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu exit $?"
