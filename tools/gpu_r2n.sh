#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_tests.log 2>&1; tail -3 gpurun_out/r2n_tests.log
for cfg in c2 c4 c5; do
python tools/gen_config.py $cfg /tmp/$cfg.fa > /dev/null
done
timeout 300 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr > gpurun_out/r2n_c2.log 2>&1; echo "c2 $(md5sum < /tmp/c2.clstr) $(grep -E 'distance keys' gpurun_out/r2n_c2.log) $(grep Total gpurun_out/r2n_c2.log)"
timeout 600 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4.clstr > gpurun_out/r2n_c4.log 2>&1; echo "c4 $(md5sum < /tmp/c4.clstr) $(grep -E 'distance keys' gpurun_out/r2n_c4.log) $(grep Total gpurun_out/r2n_c4.log)"
MC_KEYS_NO_TILE=1 timeout 600 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4b.clstr > gpurun_out/r2n_c4b.log 2>&1; echo "c4 no tile $(md5sum < /tmp/c4b.clstr) $(grep -E 'distance keys' gpurun_out/r2n_c4b.log)"
timeout 600 bin/meshclust /tmp/c5.fa --kmer 6 --output /tmp/c5.clstr > gpurun_out/r2n_c5.log 2>&1; echo "c5 $(md5sum < /tmp/c5.clstr) $(grep -E 'distance keys' gpurun_out/r2n_c5.log) $(grep Total gpurun_out/r2n_c5.log)"
grep -E "\[|Total" gpurun_out/r2n_c4.log | head -20
