#!/bin/bash
# final refresh after the last conv change: full GPU suite + smoke, bench, launch list, conv ncu captures, conv trace
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|FAILED|^ERROR|exit" gpurun_out/pytest_gpu_$TAG.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_dram_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_$TAG.log 2>&1
echo "launch list exit $?"
timeout 300 python tools/profile_forward.py 64 2 > gpurun_out/pf_plain.log 2>&1 &&
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --clock-control none \
    -k regex:conv_umma -s 55 -c 55 -o gpurun_out/prof_conv_all_$TAG -f python tools/profile_forward.py 64 2 > gpurun_out/pf_ncu_conv_all.log 2>&1
echo "conv sections exit $?"
for L in 22 29 30; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_umma -s $L -c 1 -o gpurun_out/prof_conv_full_${L}_$TAG -f python tools/profile_forward.py 64 2 > gpurun_out/pf_ncu_conv.log 2>&1
echo "conv full $L exit $?"
done
timeout 200 python tools/conv_trace.py 200 bf16 > gpurun_out/conv_trace_$TAG.txt 2>&1
cat gpurun_out/bench_$TAG.json | cut -c1-200
