#!/bin/bash
# round 2 profiling session (B200_PROFILING.md recipe): every command first runs plain and must exit 0, then once under ncu
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
: > gpurun_out/r02_prof_status.txt
run() { local name=$1; shift; "$@" > gpurun_out/r02_$name.plain.log 2>&1 || { echo "$name: plain run failed" >> gpurun_out/r02_prof_status.txt; return 1; }; return 0; }
if run bench python bench.py --steps 1 --warmup 3 --scaling-only; then
  $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 3 --scaling-only > gpurun_out/r02_bench.ncu.log 2>&1
  echo "bench launches rc=$?" >> gpurun_out/r02_prof_status.txt
fi
if run scan_c2 python tools/prof_scan.py --shape c2 --chain --launches 12; then
  $NCU --set full --import-source on -k regex:scan_tma -s 6 -c 4 -f -o gpurun_out/r02_scan_c2_chain python tools/prof_scan.py --shape c2 --chain --launches 12 > gpurun_out/r02_scan_c2.ncu.log 2>&1
  echo "scan c2 rc=$?" >> gpurun_out/r02_prof_status.txt
fi
if run scan_c4 python tools/prof_scan.py --shape c4 --chain --launches 3; then
  $NCU --set full --import-source on -k regex:scan_tma -s 1 -c 2 -f -o gpurun_out/r02_scan_c4_chain python tools/prof_scan.py --shape c4 --chain --launches 3 > gpurun_out/r02_scan_c4.ncu.log 2>&1
  echo "scan c4 rc=$?" >> gpurun_out/r02_prof_status.txt
fi
if run k1_c2 python tools/prof_k1.py c2; then
  $NCU --set full --import-source on -k regex:"kmer_count|validate_kernel" -c 2 -f -o gpurun_out/r02_k1_c2 python tools/prof_k1.py c2 > gpurun_out/r02_k1_c2.ncu.log 2>&1
  echo "k1 c2 rc=$?" >> gpurun_out/r02_prof_status.txt
fi
if run keys_c4 python tools/prof_keys.py c4; then
  $NCU --set full --import-source on -k regex:dist_keys -c 1 -f -o gpurun_out/r02_keys_c4 python tools/prof_keys.py c4 > gpurun_out/r02_keys_c4.ncu.log 2>&1
  echo "keys c4 rc=$?" >> gpurun_out/r02_prof_status.txt
fi
if run nw_c2 python tools/time_nw.py c2 3000; then
  $NCU --set full --import-source on -k regex:nw_kernel -s 1 -c 1 -f -o gpurun_out/r02_nw_c2 python tools/time_nw.py c2 3000 > gpurun_out/r02_nw_c2.ncu.log 2>&1
  echo "nw c2 rc=$?" >> gpurun_out/r02_prof_status.txt
fi
python tools/gen_config.py c2 /tmp/c2.fa > /dev/null
if run pa_c2 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr; then
  $NCU --set full --import-source on -k regex:phase_a_kernel -c 1 -f -o gpurun_out/r02_phase_a_c2 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2n.clstr > gpurun_out/r02_pa_c2.ncu.log 2>&1
  echo "phase_a c2 rc=$? $(md5sum < /tmp/c2n.clstr)" >> gpurun_out/r02_prof_status.txt
fi
cat gpurun_out/r02_prof_status.txt
ls -la gpurun_out/*.ncu-rep | grep r02
