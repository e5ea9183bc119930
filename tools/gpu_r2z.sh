#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ingest" 2>&1 | tail -3
for cfg in c2 c4 c5; do python tools/gen_config.py $cfg /tmp/$cfg.fa > /dev/null; done
for i in 1 2; do
timeout 600 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4.clstr > gpurun_out/r2z_c4_$i.log 2>&1; echo "c4 rc=$? $(md5sum < /tmp/c4.clstr) want f0917a7a"
grep -E "\[|Total|Read|Accum" gpurun_out/r2z_c4_$i.log | grep -v "^bounds"
done
timeout 600 bin/meshclust /tmp/c5.fa --kmer 6 --output /tmp/c5.clstr > gpurun_out/r2z_c5.log 2>&1; echo "c5 rc=$? $(md5sum < /tmp/c5.clstr) want 36aebc3b"
grep -E "\[|Total|Read|Accum" gpurun_out/r2z_c5.log | grep -v "^bounds"
timeout 600 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr > gpurun_out/r2z_c2.log 2>&1; echo "c2 rc=$? $(md5sum < /tmp/c2.clstr) want 83cffd7e"
grep -E "\[|Total|Read|Accum" gpurun_out/r2z_c2.log | grep -v "^bounds"
