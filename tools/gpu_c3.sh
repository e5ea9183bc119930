#!/bin/bash
# --align path on one GPU: the two alignment golden cases and the C3 config through the CLI
set -u
mkdir -p gpurun_out
python -m pytest tests/test_host_logic.py -m gpu -x -q -k "reference_gpu and (G or H)" > gpurun_out/pytest_c3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_c3.log
python tools/gen_config.py c3 /tmp/c3.fa > gpurun_out/gen.log 2>&1
( time bin/meshclust /tmp/c3.fa --id 0.70 --align --output /tmp/c3.clstr ) > gpurun_out/cli_c3_final.log 2>&1
md5sum /tmp/c3.clstr >> gpurun_out/cli_c3_final.log
