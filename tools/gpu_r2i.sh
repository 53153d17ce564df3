#!/bin/bash
# session call 7 (two GPUs): the second-device / replica tests and the torchrun bench line
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2i_gpus.txt 2>&1
python -m pytest tests/test_gpu_c64.py tests/test_gpu_f128.py -m gpu -q -k "second_device or replicas or clone" > gpurun_out/r2i_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err
python tools/multi_gpu_host_probe.py > gpurun_out/r2i_host_multi.txt 2>&1
tail -3 gpurun_out/r2i_pytest_2gpu.log; head -c 400 gpurun_out/r2i_bench_2gpu.json; cat gpurun_out/r2i_host_multi.txt
