#!/bin/bash
# one GPU-box session: parity tests, CLI stage timings on C2 and C4
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python tools/gen_config.py c2 /tmp/c2.fa > gpurun_out/gen.log 2>&1
( time bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr ) > gpurun_out/cli_c2.log 2>&1
( time bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2b.clstr ) > gpurun_out/cli_c2b.log 2>&1
python tools/gen_config.py c4 /tmp/c4.fa >> gpurun_out/gen.log 2>&1
( time timeout 900 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4.clstr ) > gpurun_out/cli_c4.log 2>&1
grep -c ">Cluster" /tmp/c4.clstr >> gpurun_out/cli_c4.log 2>&1
md5sum /tmp/c2.clstr /tmp/c4.clstr > gpurun_out/clstr_md5.txt
