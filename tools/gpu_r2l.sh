#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_host_logic.py -q -m gpu 2>&1 | tail -3
python tools/gen_config.py c5 /tmp/c5.fa > /dev/null
for spec in 0 2 3 4; do
if [ $spec = 0 ]; then unset MC_SPLIT_SPEC; else export MC_SPLIT_SPEC=$spec; fi
timeout 600 bin/meshclust /tmp/c5.fa --kmer 6 --output /tmp/c5.clstr > gpurun_out/r2l_c5_$spec.log 2>&1
echo "c5 spec=$spec rc=$? $(md5sum < /tmp/c5.clstr) $(grep -E 'alignment rounds' gpurun_out/r2l_c5_$spec.log) $(grep -E 'labels|Total' gpurun_out/r2l_c5_$spec.log | tr '\n' ' ')"
done
unset MC_SPLIT_SPEC
python tools/gen_config.py c3 /tmp/c3.fa > /dev/null
timeout 600 bin/meshclust /tmp/c3.fa --id 0.70 --align --output /tmp/c3.clstr > gpurun_out/r2l_c3.log 2>&1
echo "c3 rc=$? $(md5sum < /tmp/c3.clstr)"; grep -E "\[|Total" gpurun_out/r2l_c3.log | tail -12
