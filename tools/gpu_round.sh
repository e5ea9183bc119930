#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python tools/time_scan.py > gpurun_out/time_scan_pre.log 2>&1
python tools/trace_scan.py c2 > gpurun_out/trace_pre_c2.log 2>&1
python bench.py --steps 50 --warmup 5 --no-extra > gpurun_out/bench_pre.json 2> gpurun_out/bench_pre.err
