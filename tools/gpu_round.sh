#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/gen_config.py c4 /tmp/c4.fa > gpurun_out/gen.log 2>&1
( time MC_DEBUG_TIMING=1 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4_a.clstr ) 2>&1 | grep -v "mc_align\|mc_ctx_create" > gpurun_out/cli_c4_a.log
( time bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4_a.clstr ) > gpurun_out/cli_c4_a2.log 2>&1
( time MC_COMPACT_MIN_ROWS=999999999 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4_b.clstr ) > gpurun_out/cli_c4_b.log 2>&1
