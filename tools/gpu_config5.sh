#!/bin/bash
# BASELINE config 5 on N GPUs: one 2000-frame 224x224 video split by clip range with the NCCL halo exchange, checked
# against the single-GPU result on rank 0; then the video-sharded bench (config 4 shape: one video per rank per step).
N=${1:-8}; TAG=${2:-r01}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR tools/long_video_check.py 2000 224 224 bf16 > gpurun_out/config5_${N}gpu_$TAG.log 2>&1; echo "exit $?" >> gpurun_out/config5_${N}gpu_$TAG.log
grep -E "check|exit|Error" gpurun_out/config5_${N}gpu_$TAG.log | tail -3
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "bench exit $?" >> gpurun_out/bench_${N}gpu_$TAG.err
cat gpurun_out/bench_${N}gpu_$TAG.json | cut -c1-700; tail -2 gpurun_out/bench_${N}gpu_$TAG.err
