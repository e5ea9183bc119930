#!/bin/bash
mkdir -p gpurun_out
python tools/gen_config.py c2 /tmp/c2s.fa 3000 > /dev/null
MC_PA_COMPACT_MIN=64 timeout 60 bin/meshclust /tmp/c2s.fa --id 0.97 --kmer 4 --output /tmp/c2s.clstr > gpurun_out/r2t_s.log 2>&1; echo "small compact rc=$? $(md5sum < /tmp/c2s.clstr)"; grep Accum gpurun_out/r2t_s.log
MC_PA_NO_COMPACT=1 timeout 60 bin/meshclust /tmp/c2s.fa --id 0.97 --kmer 4 --output /tmp/c2s.clstr > gpurun_out/r2t_sn.log 2>&1; echo "small nocompact rc=$? $(md5sum < /tmp/c2s.clstr)"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "accumulate_run" 2>&1 | tail -5
for cfg in c2 c4 c5; do python tools/gen_config.py $cfg /tmp/$cfg.fa > /dev/null; done
run() { # name file args
  local name=$1 f=$2; shift 2
  timeout 600 bin/meshclust $f "$@" --output /tmp/$name.clstr > gpurun_out/r2t_$name.log 2>&1; echo "$name rc=$? $(md5sum < /tmp/$name.clstr)"; grep -E "Accumulation|Total" gpurun_out/r2t_$name.log
  MC_PA_NO_COMPACT=1 timeout 600 bin/meshclust $f "$@" --output /tmp/$name.n.clstr > gpurun_out/r2t_${name}_n.log 2>&1; echo "$name nocompact rc=$? $(md5sum < /tmp/$name.n.clstr)"; grep -E "Accumulation|Total" gpurun_out/r2t_${name}_n.log
}
run c2 /tmp/c2.fa --id 0.97 --kmer 4
run c4 /tmp/c4.fa --id 0.90 --kmer 5
run c5 /tmp/c5.fa --kmer 6
