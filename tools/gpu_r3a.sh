#!/bin/bash
mkdir -p gpurun_out
for cfg in c2 c4 c5 c1; do python tools/gen_config.py $cfg /tmp/$cfg.fa > /dev/null; done
timeout 600 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4.clstr > gpurun_out/r3a_c4.log 2>&1; echo "c4 rc=$? $(md5sum < /tmp/c4.clstr) want f0917a7a"
grep -E "\[|Total|Read|Accum" gpurun_out/r3a_c4.log | grep -v "^bounds"
MC_SPLIT_FULL_SORT=1 timeout 600 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4.clstr > gpurun_out/r3a_c4_full.log 2>&1; echo "c4 full sorts rc=$? $(md5sum < /tmp/c4.clstr)"
grep -E "split" gpurun_out/r3a_c4_full.log
timeout 600 bin/meshclust /tmp/c5.fa --kmer 6 --output /tmp/c5.clstr > gpurun_out/r3a_c5.log 2>&1; echo "c5 rc=$? $(md5sum < /tmp/c5.clstr) want 36aebc3b"
grep -E "\[|Total" gpurun_out/r3a_c5.log | grep -v "^bounds" | head -12
timeout 600 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr > gpurun_out/r3a_c2.log 2>&1; echo "c2 rc=$? $(md5sum < /tmp/c2.clstr) want 83cffd7e"
grep -E "split|Total|ahead" gpurun_out/r3a_c2.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
