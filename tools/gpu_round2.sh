#!/bin/bash
# 2-GPU session: scaling of the sharded scans + sharded Phase A in the CLI
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "sharded or accumulate or scan" > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; echo "rc=$?" >> gpurun_out/bench_s2.err
python bench.py --steps 50 --warmup 5 --no-extra > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err; echo "rc=$?" >> gpurun_out/bench_s1.err
python tools/gen_config.py c4 /tmp/c4.fa > gpurun_out/gen.log 2>&1
( time bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4_g1.clstr ) > gpurun_out/cli_c4_g1.log 2>&1
( time bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --gpus 2 --output /tmp/c4_g2.clstr ) > gpurun_out/cli_c4_g2.log 2>&1
( time MC_DEBUG_TIMING=1 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --gpus 2 --output /tmp/c4_g2b.clstr ) 2>&1 | grep -v "mc_align\|mc_ctx" > gpurun_out/cli_c4_g2_dbg.log
md5sum /tmp/c4_g1.clstr /tmp/c4_g2.clstr > gpurun_out/clstr_md5_g2.txt
