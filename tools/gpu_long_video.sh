#!/bin/bash
# BASELINE configs[4] on N GPUs of one box: bench.py --workload long-video under torchrun; one JSON line per run in gpurun_out/
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --workload long-video --steps 3 --warmup 2 > gpurun_out/bench_long_${N}gpu_$TAG.json 2> gpurun_out/bench_long_${N}gpu_$TAG.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload long-video --steps 3 --warmup 2 \
    > gpurun_out/bench_long_${N}gpu_$TAG.json 2> gpurun_out/bench_long_${N}gpu_$TAG.err
fi
echo "long-video x$N exit $?"; tail -2 gpurun_out/bench_long_${N}gpu_$TAG.err; cat gpurun_out/bench_long_${N}gpu_$TAG.json
