#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_host_logic.py -x -q -m gpu -k "split_alignments" 2>&1 | tail -5
for cfg in c5 c3; do python tools/gen_config.py $cfg /tmp/$cfg.fa > /dev/null; done
for g in 2 1; do
timeout 600 bin/meshclust /tmp/c5.fa --kmer 6 --gpus $g --output /tmp/c5.clstr > gpurun_out/r2y_c5_g$g.log 2>&1; echo "c5 gpus $g rc=$? $(md5sum < /tmp/c5.clstr) want 36aebc3b"
grep -E "split|labels|Total|copied" gpurun_out/r2y_c5_g$g.log
timeout 600 bin/meshclust /tmp/c3.fa --id 0.70 --align --gpus $g --output /tmp/c3.clstr > gpurun_out/r2y_c3_g$g.log 2>&1; echo "c3 gpus $g rc=$? $(md5sum < /tmp/c3.clstr) want 70c4ae74"
grep -E "copied|Total|Accum|Update" gpurun_out/r2y_c3_g$g.log
done
