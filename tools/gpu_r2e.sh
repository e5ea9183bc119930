#!/bin/bash
# round 2: tagged-record exchange in the persistent Phase-A kernel (parity + step timing)
mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "accumulate_run" 2>&1 | tail -3
done
timeout 1500 python -m pytest tests/test_host_logic.py -q -m gpu -k "identical_to_reference_gpu or host_driven" > gpurun_out/r2e_cli.log 2>&1
tail -5 gpurun_out/r2e_cli.log
python tools/gen_config.py c2 /tmp/c2.fa > /dev/null
for i in 1 2 3; do
MC_PA_TRACE=4000 timeout 300 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr > gpurun_out/r2e_c2_run$i.log 2>&1
echo "c2 rc=$? $(md5sum < /tmp/c2.clstr)"
done
grep -E "Accumulation|trace|Total|Pairs" gpurun_out/r2e_c2_run3.log
python tools/gen_config.py c4 /tmp/c4.fa > /dev/null
MC_PA_TRACE=4000 timeout 600 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4.clstr > gpurun_out/r2e_c4.log 2>&1
echo "c4 rc=$? $(md5sum < /tmp/c4.clstr)"
grep -E "Accumulation|trace|Total|Pairs" gpurun_out/r2e_c4.log
