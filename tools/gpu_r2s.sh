#!/bin/bash
mkdir -p gpurun_out
python tools/gen_config.py c2 /tmp/c2s.fa 3000 > /dev/null
MC_PA_TRACE=10 MC_PA_COMPACT_MIN=64 timeout 60 bin/meshclust /tmp/c2s.fa --id 0.97 --kmer 4 --output /tmp/c2s.clstr > gpurun_out/r2s.log 2>&1
echo "rc=$?"; grep -E "phase_a|meshclust:|Accum" gpurun_out/r2s.log | cut -c1-220 | tail -40
