#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2p_bench_n2.json 2> gpurun_out/r2p_bench_n2.err
echo "rc=$?"; tail -8 gpurun_out/r2p_bench_n2.err; cat gpurun_out/r2p_bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2p_ref_n2.json 2> gpurun_out/r2p_ref_n2.err
echo "ref rc=$?"; cat gpurun_out/r2p_ref_n2.json | cut -c1-400
