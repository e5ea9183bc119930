#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests -m gpu -x -q -k "sharded or alignment" > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest2.log
MC_DEBUG_TIMING=1 timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/diag_n2_dbg.json 2> gpurun_out/diag_n2_dbg.err
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/diag_n2.json 2> gpurun_out/diag_n2.err
python tools/time_nw.py c2 3000 > gpurun_out/time_nw.log 2>&1
