#!/bin/bash
# last refresh: 20-step bench + ncu launch list of one step
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_dram_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_$TAG.log 2>&1
echo "launch list exit $?"
cut -c1-160 gpurun_out/bench_$TAG.json
