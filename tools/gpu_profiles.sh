#!/bin/bash
# Evidence run for profiles/: plain bench, ncu launch list (+ DRAM bytes) of bench steps, ncu --set full of the hot kernels.
# Every ncu command runs only after the same command has exited 0 without ncu.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
if [ -n "$BENCH_ONLY" ]; then exit 0; fi
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_dram_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_$TAG.log 2>&1
echo "launch list exit $?"
timeout 300 python tools/profile_forward.py 64 2 > gpurun_out/pf_plain.log 2>&1 &&
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --clock-control none \
    -k regex:conv_umma -s 55 -c 55 -o gpurun_out/prof_conv_all_$TAG -f python tools/profile_forward.py 64 2 > gpurun_out/pf_ncu_conv_all.log 2>&1
echo "conv sections exit $?"
# full captures (source-level) of three representative convolutions of the first forward: launch 22 = layer1 clip-edge
# temporal 3x1x1 (time-segmented, 64 columns), 29 = layer2 spatial 1x3x3 128->288, 30 = layer2 temporal 3x1x1 288->128 (N-split)
for L in 22 29 30; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_umma -s $L -c 1 -o gpurun_out/prof_conv_full_${L}_$TAG -f python tools/profile_forward.py 64 2 > gpurun_out/pf_ncu_conv.log 2>&1
echo "conv full $L exit $?"
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:head_umma -s 1 -c 1 -o gpurun_out/prof_head_$TAG -f python tools/profile_forward.py 64 2 > gpurun_out/pf_ncu_head.log 2>&1
echo "head full exit $?"
for DT in bf16 fp32; do
timeout 300 python tools/bench_warp_fuse.py --once --dtypes $DT > gpurun_out/wf_plain_$DT.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none -k regex:warp_fuse -s 1 -c 1 -o gpurun_out/prof_wf_${DT}_$TAG -f python tools/bench_warp_fuse.py --once --dtypes $DT > gpurun_out/pf_ncu_wf_$DT.log 2>&1
echo "wf full $DT exit $?"
done
timeout 300 python tools/bench_warp_fuse.py --clips 16 64 256 > gpurun_out/warp_fuse_sweep_$TAG.jsonl 2>&1
timeout 300 python tools/bench_warp_fuse.py --clips 64 --size 224 --flow-px 0 4 >> gpurun_out/warp_fuse_sweep_$TAG.jsonl 2>&1
timeout 200 python tools/conv_trace.py 200 bf16 > gpurun_out/conv_trace_$TAG.txt 2>&1
du -sh gpurun_out; cat gpurun_out/bench_$TAG.json | cut -c1-300
