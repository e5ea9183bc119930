#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python bench.py --steps 50 --warmup 5 --scaling-only > gpurun_out/scale2_n1.json 2> gpurun_out/scale2_n1.err
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --scaling-only > gpurun_out/scale2_n2.json 2> gpurun_out/scale2_n2.err
