#!/bin/bash
# the driver's scaling launch, replayed: torchrun over N GPUs of one box.   usage: gpu_scale.sh <N> <tag>
N=${1:-8}; tag=${2:-scale}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt 2>&1
free -g >> gpurun_out/${tag}_gpus.txt 2>&1
t0=$SECONDS
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
echo "exit $? after $((SECONDS - t0)) s" >> gpurun_out/${tag}_bench_${N}gpu.err
tail -c 600 gpurun_out/${tag}_bench_${N}gpu.json; grep -E "exit" gpurun_out/${tag}_bench_${N}gpu.err
