#!/bin/bash
# multi-GPU validation: long-video clip-range split (config 5) and the video-sharded bench (config 4 shape) on N GPUs
N=${1:-2}; TAG=${2:-r01}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR tools/long_video_check.py 400 112 112 bf16 > gpurun_out/long_video_${N}gpu_112_$TAG.log 2>&1; echo "exit $?" >> gpurun_out/long_video_${N}gpu_112_$TAG.log
grep -E "check|exit|Error" gpurun_out/long_video_${N}gpu_112_$TAG.log | tail -3
timeout 900 $TR tools/long_video_check.py 300 224 224 bf16 > gpurun_out/long_video_${N}gpu_224_$TAG.log 2>&1; echo "exit $?" >> gpurun_out/long_video_${N}gpu_224_$TAG.log
grep -E "check|exit|Error" gpurun_out/long_video_${N}gpu_224_$TAG.log | tail -3
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "bench exit $?" >> gpurun_out/bench_${N}gpu_$TAG.err
cat gpurun_out/bench_${N}gpu_$TAG.json | cut -c1-1500; tail -2 gpurun_out/bench_${N}gpu_$TAG.err
