#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err; echo "rc=$?" >> gpurun_out/bench_g1.err
python tools/time_scan.py > gpurun_out/time_scan_batch.log 2>&1
