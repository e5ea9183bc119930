#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/driver_like_n2.json 2> gpurun_out/driver_like_n2.err; echo "rc=$?" >> gpurun_out/driver_like_n2.err
