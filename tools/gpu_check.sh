#!/bin/bash
# quick one-GPU check: parity tests (optionally a -k expression) + the default bench line
set -u
mkdir -p gpurun_out
if [ "${1:-}" != "" ]; then
  python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
  exit 0
fi
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --no-extra > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; echo "rc=$?" >> gpurun_out/bench_check.err
