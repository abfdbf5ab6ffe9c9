#!/bin/bash
TAG=${1:-r01s}
timeout 900 python -m pytest tests -m gpu -q -x -k "warp_fus or long_video or fusion" > gpurun_out/pytest_wf_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_wf_$TAG.log
tail -n 15 gpurun_out/pytest_wf_$TAG.log
timeout 600 python tools/bench_warp_fuse.py > gpurun_out/warp_fuse_sweep_$TAG.jsonl 2>&1
cut -c1-175 gpurun_out/warp_fuse_sweep_$TAG.jsonl
