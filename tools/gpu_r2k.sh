#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "alignment" 2>&1 | tail -1
for cfg in c2 c3; do python tools/time_nw.py $cfg 4000 2>&1 | tail -3; done
python tools/time_nw.py c5 600 2>&1 | sed -n 2,3p
