#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "sharded or scan" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
for m in A B C; do python tools/time_sharded.py $m; done > gpurun_out/time_sharded.log 2>&1
