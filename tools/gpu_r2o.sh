#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hist or encode or ragged or invalid or golden or keys or iupac" 2>&1 | tail -2
MC_DEBUG_TIMING=1 python tools/prof_k1.py c2 2>&1 | grep -E "kmer_count|c2:" | tail -4
MC_DEBUG_TIMING=1 python tools/prof_k1.py c4 2>&1 | grep -E "kmer_count|c4:" | tail -2
python tools/prof_keys.py c4 | tail -2
python tools/prof_keys.py c5 | tail -1
nsys --version 2>/dev/null | head -1
