#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for cfg in c2 c4 c5; do python tools/gen_config.py $cfg /tmp/$cfg.fa > /dev/null; done
run() { # name file args
  local name=$1 f=$2; shift 2
  for sh in 3 2 1; do
  MC_PA_COMPACT_SHIFT=$sh timeout 600 bin/meshclust $f "$@" --output /tmp/$name.clstr > gpurun_out/r2u_${name}_$sh.log 2>&1; echo "$name shift $sh rc=$? $(md5sum < /tmp/$name.clstr)"; grep -E "Accumulation|Total" gpurun_out/r2u_${name}_$sh.log
  done
}
run c2 /tmp/c2.fa --id 0.97 --kmer 4
run c4 /tmp/c4.fa --id 0.90 --kmer 5
run c5 /tmp/c5.fa --kmer 6
