#!/bin/bash
# multi-GPU validation: new head-geometry tests (1 GPU), long-video clip-range split and the video-sharded bench on N GPUs
N=${1:-2}; TAG=${2:-r01u}
timeout 900 python -m pytest tests -m gpu -q -s -x -k "head_geometries or dense_video" > gpurun_out/pytest_geom_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_geom_$TAG.log
grep -E "^\[|passed|failed|Error|exit" gpurun_out/pytest_geom_$TAG.log | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
true
true
timeout 900 $TR tools/long_video_check.py 300 224 224 bf16 > gpurun_out/long_video_${N}gpu_224_$TAG.log 2>&1; echo "exit $?" >> gpurun_out/long_video_${N}gpu_224_$TAG.log
grep -E "check|exit|Error" gpurun_out/long_video_${N}gpu_224_$TAG.log | tail -3
true
true
