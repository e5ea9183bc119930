#!/bin/bash
# bench.py at N GPUs the way the driver launches it: bash tools/gpu_scale.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r3_scale_n$N.json 2> gpurun_out/r3_scale_n$N.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/r3_scale_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","n_gpus")}); print(d["roofline"]["frac"], d["clocks"], d["config"]["exchange"]); print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"]); print(d["extra"]["c4_shape_scan"].get("dependent_chain")); print(d["seqs_clustered"]["wall_s"], d["seqs_clustered"]["value"])
PY
