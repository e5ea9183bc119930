#!/bin/bash
# end-of-round validation on one B200: smoke, parity tests, bench (both arms), the CLI on C1..C5
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "rc=$?" >> gpurun_out/final_bench_ref.err
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?" >> gpurun_out/final_bench.err
for cfg in c1 c2 c3 c4 c5; do
  python tools/gen_config.py $cfg /tmp/$cfg.fa >> gpurun_out/final_gen.log 2>&1
done
( time bin/meshclust /tmp/c1.fa --id 0.90 --kmer 3 --output /tmp/c1.clstr ) > gpurun_out/final_cli_c1.log 2>&1
( time bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr ) > gpurun_out/final_cli_c2.log 2>&1
( time timeout 600 bin/meshclust /tmp/c3.fa --id 0.70 --align --output /tmp/c3.clstr ) > gpurun_out/final_cli_c3.log 2>&1
( time bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4.clstr ) > gpurun_out/final_cli_c4.log 2>&1
( time timeout 900 bin/meshclust /tmp/c5.fa --kmer 6 --output /tmp/c5.clstr ) > gpurun_out/final_cli_c5.log 2>&1
for cfg in c1 c2 c3 c4 c5; do echo "$cfg $(grep -c '>Cluster' /tmp/$cfg.clstr) clusters $(md5sum < /tmp/$cfg.clstr)"; done > gpurun_out/final_clstr.txt 2>&1
# ncu evidence for the final kernels (each command ran plain above or runs plain first)
NCU="ncu --clock-control none"
python bench.py --steps 5 --warmup 3 --no-extra > gpurun_out/final_prof_bench.plain.log 2>&1 && \
  $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r01c_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-extra > gpurun_out/final_prof_bench.ncu.log 2>&1
python tools/prof_scan.py --shape c2 > gpurun_out/final_prof_scan_c2.plain.log 2>&1 && \
  $NCU --set full --import-source on -k regex:scan_tma -c 2 -f -o gpurun_out/r01c_scan_c2_batch python tools/prof_scan.py --shape c2 > gpurun_out/final_prof_scan_c2.ncu.log 2>&1
MC_SCAN_NO_BATCH=1 python tools/prof_scan.py --shape c2 > gpurun_out/final_prof_scan_c2s.plain.log 2>&1 && \
  MC_SCAN_NO_BATCH=1 $NCU --set full --import-source on -k regex:scan_tma -c 6 -f -o gpurun_out/r01c_scan_c2_single python tools/prof_scan.py --shape c2 > gpurun_out/final_prof_scan_c2s.ncu.log 2>&1
MC_SCAN_NO_BATCH=1 python tools/prof_scan.py --shape c4 --launches 3 > gpurun_out/final_prof_scan_c4.plain.log 2>&1 && \
  MC_SCAN_NO_BATCH=1 $NCU --set full --import-source on -k regex:scan_tma -c 3 -f -o gpurun_out/r01c_scan_c4_single python tools/prof_scan.py --shape c4 --launches 3 > gpurun_out/final_prof_scan_c4.ncu.log 2>&1
python tools/time_nw.py c2 3000 > gpurun_out/final_time_nw.log 2>&1
python tools/time_scan.py > gpurun_out/final_time_scan.log 2>&1
