#!/bin/bash
# 8-GPU session: bench.py at N = 8 and the CLI on the 1M-sequence config at 1/2/4/8 GPUs
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "gpus=$NG" > gpurun_out/scale8_info.txt
nvidia-smi topo -m >> gpurun_out/scale8_info.txt 2>&1
python tools/gen_config.py c4 /tmp/c4.fa > gpurun_out/gen.log 2>&1 &
GEN=$!
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $NG --steps 50 --warmup 5 > gpurun_out/scale_n$NG.json 2> gpurun_out/scale_n$NG.err
echo "rc=$?" >> gpurun_out/scale_n$NG.err
wait $GEN
for G in 1 2 4 8; do
  if [ "$G" -gt "$NG" ]; then continue; fi
  ( time MC_DEBUG_TIMING=1 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --gpus $G --output /tmp/c4_g$G.clstr ) 2>&1 | grep -v "mc_align\|mc_ctx_create" > gpurun_out/cli_c4_scale_g$G.log
done
md5sum /tmp/c4_g*.clstr > gpurun_out/clstr_md5_scale.txt
