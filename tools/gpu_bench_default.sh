#!/bin/bash
set -u
mkdir -p gpurun_out
python bench.py --no-extra > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?" >> gpurun_out/bench_default.err
