#!/bin/bash
# multi-GPU session: weak scaling of bench.py at N = 1, 2, 4 (and 8 when the box has them)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "scan or sharded or accumulate" > gpurun_out/pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest4.log
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  if [ "$N" -gt "$NG" ]; then continue; fi
  if [ "$N" -eq 1 ]; then
    python bench.py --steps 50 --warmup 5 --no-extra > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "rc=$?" >> gpurun_out/scale_n$N.err
done
