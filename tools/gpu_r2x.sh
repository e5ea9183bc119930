#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 1200 python -m pytest tests/test_host_logic.py -x -q -m gpu -k "split_alignments or sharded" 2>&1 | tail -5
for cfg in c2 c5 c3; do python tools/gen_config.py $cfg /tmp/$cfg.fa > /dev/null; done
for g in 1 2; do
timeout 600 bin/meshclust /tmp/c5.fa --kmer 6 --gpus $g --output /tmp/c5.clstr > gpurun_out/r2x_c5_g$g.log 2>&1; echo "c5 gpus $g rc=$? $(md5sum < /tmp/c5.clstr) want 36aebc3b"
grep -E "\[|Total|Read|Accum" gpurun_out/r2x_c5_g$g.log | grep -v "^bounds"
done
for g in 1 2; do
timeout 600 bin/meshclust /tmp/c3.fa --id 0.70 --align --gpus $g --output /tmp/c3.clstr > gpurun_out/r2x_c3_g$g.log 2>&1; echo "c3 gpus $g rc=$? $(md5sum < /tmp/c3.clstr)"
grep -E "copied|Total|Accum|Update|split" gpurun_out/r2x_c3_g$g.log
done
timeout 600 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --gpus 2 --output /tmp/c2.clstr > gpurun_out/r2x_c2_g2.log 2>&1; echo "c2 gpus 2 rc=$? $(md5sum < /tmp/c2.clstr) want 83cffd7e"
grep -E "gpus|Total" gpurun_out/r2x_c2_g2.log
