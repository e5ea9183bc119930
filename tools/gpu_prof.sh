#!/bin/bash
# 1-GPU profiling session (B200_PROFILING.md recipe): every command first runs plain and must exit 0,
# then once under ncu.  Reports land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.md
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
run() { # name, then the command
  local name=$1; shift
  "$@" > gpurun_out/prof_$name.plain.log 2>&1 || { echo "$name: plain run failed" >> gpurun_out/prof_status.txt; return 1; }
  return 0
}
: > gpurun_out/prof_status.txt
# launch list of the bench command
if run bench python bench.py --steps 5 --warmup 3 --no-extra; then
  $NCU --metrics gpu__time_duration.sum -c 800 --csv --log-file gpurun_out/r01b_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-extra > gpurun_out/prof_bench.ncu.log 2>&1
  echo "bench launches rc=$?" >> gpurun_out/prof_status.txt
fi
if run scan_c2 python tools/prof_scan.py --shape c2; then
  $NCU --set full --import-source on -k regex:scan_tma -c 6 -f -o gpurun_out/r01b_scan_c2 python tools/prof_scan.py --shape c2 > gpurun_out/prof_scan_c2.ncu.log 2>&1
  echo "scan c2 rc=$?" >> gpurun_out/prof_status.txt
fi
if run scan_c4 python tools/prof_scan.py --shape c4 --launches 3; then
  $NCU --set full --import-source on -k regex:scan_tma -c 3 -f -o gpurun_out/r01b_scan_c4 python tools/prof_scan.py --shape c4 --launches 3 > gpurun_out/prof_scan_c4.ncu.log 2>&1
  echo "scan c4 rc=$?" >> gpurun_out/prof_status.txt
fi
if run step_c4 python tools/prof_step.py --shape c4 --clusters 3; then
  $NCU --set full --import-source on -k regex:accumulate_tail -c 6 -f -o gpurun_out/r01_tail_c4 python tools/prof_step.py --shape c4 --clusters 3 > gpurun_out/prof_step_c4.ncu.log 2>&1
  echo "tail c4 rc=$?" >> gpurun_out/prof_status.txt
fi
if run k1_c2 python tools/prof_k1.py c2; then
  $NCU --set full --import-source on -k regex:"kmer_hist|encode_kernel" -c 2 -f -o gpurun_out/r01_k1_c2 python tools/prof_k1.py c2 > gpurun_out/prof_k1_c2.ncu.log 2>&1
  echo "k1 c2 rc=$?" >> gpurun_out/prof_status.txt
fi
if run nw_c2 python tools/time_nw.py c2 3000; then
  $NCU --set full --import-source on -k regex:nw_kernel -s 1 -c 1 -f -o gpurun_out/r01_nw_c2 python tools/time_nw.py c2 3000 > gpurun_out/prof_nw_c2.ncu.log 2>&1
  echo "nw c2 rc=$?" >> gpurun_out/prof_status.txt
fi
if run upd_c4 python tools/prof_step.py --shape c2 --clusters 2; then :; fi
