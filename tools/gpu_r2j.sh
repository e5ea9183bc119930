#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "batch_launch or abi" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
echo "rc=$?"; tail -5 gpurun_out/r2j_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench.json')); print({k:d[k] for k in ['value','ms_per_step','clocks','gpu_launches']}); print(d['roofline']); print(d['independent_scans'])"
