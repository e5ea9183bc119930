#!/bin/bash
# 8-GPU session: bench.py weak scaling at N = 1, 2, 4, 8 on one box, and the CLI on C4 at 1 and 8 GPUs
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
python tools/gen_config.py c4 /tmp/c4.fa > gpurun_out/gen.log 2>&1 &
GEN=$!
for N in 1 2 4 8; do
  if [ "$N" -gt "$NG" ]; then continue; fi
  if [ "$N" -eq 1 ]; then
    python bench.py --steps 50 --warmup 5 --scaling-only > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 50 --warmup 5 --scaling-only > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "rc=$?" >> gpurun_out/scale_n$N.err
done
wait $GEN
for G in 1 8; do
  ( time bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --gpus $G --output /tmp/c4_g$G.clstr ) > gpurun_out/cli_c4_scale_g$G.log 2>&1
done
md5sum /tmp/c4_g*.clstr > gpurun_out/clstr_md5_scale.txt
