#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hist or encode or ragged or invalid or alignment or iupac or digits" 2>&1 | tail -3
MC_DEBUG_TIMING=1 python tools/prof_k1.py c2 2>&1 | tail -8
MC_DEBUG_TIMING=1 python tools/prof_k1.py c4 2>&1 | tail -3; MC_DEBUG_TIMING=1 python tools/prof_k1.py c5 2>&1 | tail -3
