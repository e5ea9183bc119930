#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > gpurun_out/smi_start.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python tools/gen_config.py c4 /tmp/c4.fa > gpurun_out/gen.log 2>&1
( time bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4_a.clstr ) > gpurun_out/cli_c4_a.log 2>&1
( time bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4_b.clstr ) > gpurun_out/cli_c4_b.log 2>&1
( time bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --gpus 2 --output /tmp/c4_c.clstr ) > gpurun_out/cli_c4_c.log 2>&1
md5sum /tmp/c4_a.clstr /tmp/c4_c.clstr > gpurun_out/clstr_md5.txt
python tools/prof_step.py --shape c4 --clusters 30 > gpurun_out/prof_step_c4.log 2>&1
python tools/prof_step.py --shape c2 --clusters 100 > gpurun_out/prof_step_c2.log 2>&1
python tools/gen_config.py c3 /tmp/c3.fa >> gpurun_out/gen.log 2>&1
( time timeout 400 bin/meshclust /tmp/c3.fa --id 0.70 --align --output /tmp/c3.clstr ) > gpurun_out/cli_c3.log 2>&1
grep -c ">Cluster" /tmp/c3.clstr >> gpurun_out/cli_c3.log 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > gpurun_out/smi_end.txt
