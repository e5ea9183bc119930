#!/bin/bash
# 2-GPU session: scaling of the sharded scans
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
python bench.py --steps 50 --warmup 5 --no-extra > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err; echo "rc=$?" >> gpurun_out/bench_s1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; echo "rc=$?" >> gpurun_out/bench_s2.err
