#!/bin/bash
mkdir -p gpurun_out
python tools/gen_config.py c2 /tmp/c2.fa > /dev/null
for i in 1 2; do
MC_PA_TRACE=4000 timeout 300 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr > gpurun_out/r2g_c2_run$i.log 2>&1
echo "c2 rc=$? $(md5sum < /tmp/c2.clstr)"
done
grep -E "Accumulation|trace|Total|Pairs" gpurun_out/r2g_c2_run2.log
