#!/bin/bash
# Evidence run for profiles/: plain bench, ncu launch list of one bench step, ncu --set full of the three hot kernels.
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> gpurun_out/bench_$TAG.err
timeout 600 python tools/bench_warp_fuse.py > gpurun_out/warp_fuse_sweep_$TAG.jsonl 2>&1
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 192 -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
echo "launch list exit $?"
timeout 300 python tools/profile_forward.py 16 2 > gpurun_out/pf_plain.log 2>&1 &&
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --clock-control none \
    -k regex:conv_umma -s 54 -c 54 -o gpurun_out/prof_conv_all_$TAG -f python tools/profile_forward.py 16 2 > gpurun_out/pf_ncu_conv_all.log 2>&1
echo "conv sections exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_umma -s 55 -c 2 -o gpurun_out/prof_conv_$TAG -f python tools/profile_forward.py 16 2 > gpurun_out/pf_ncu_conv.log 2>&1
echo "conv full exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:head_umma -s 1 -c 1 -o gpurun_out/prof_head_$TAG -f python tools/profile_forward.py 16 2 > gpurun_out/pf_ncu_head.log 2>&1
echo "head full exit $?"
timeout 300 python tools/bench_warp_fuse.py --once > gpurun_out/wf_plain.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:warp_fuse -s 1 -c 3 -o gpurun_out/prof_wf_$TAG -f python tools/bench_warp_fuse.py --once > gpurun_out/pf_ncu_wf.log 2>&1
echo "wf full exit $?"
du -sh gpurun_out; ls -la gpurun_out | head -30; cat gpurun_out/bench_$TAG.json | cut -c1-400
