#!/bin/bash
# quick one-GPU check: parity tests + the default bench line
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --no-extra > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "rc=$?" >> gpurun_out/bench_check.err
