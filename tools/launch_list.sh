#!/bin/bash
# ncu launch list of one bench step: duration and DRAM bytes per launch (no source, no sections)
TAG=${1:-r01}; SKIP=${2:-160}
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $SKIP -c 200 --csv \
    --log-file gpurun_out/launches_dram_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu exit $?"
