#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "accumulate_run" 2>&1 | tail -3
python tools/gen_config.py c2 /tmp/c2.fa > /dev/null
for i in 1 2 3; do
MC_PA_TRACE=4000 timeout 300 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr > gpurun_out/r2f_c2_run$i.log 2>&1
echo "c2 rc=$? $(md5sum < /tmp/c2.clstr)"
done
grep -E "Accumulation|trace|Total|Pairs" gpurun_out/r2f_c2_run3.log
timeout 1500 python -m pytest tests/test_host_logic.py -q -m gpu -k "identical_to_reference_gpu or host_driven" > gpurun_out/r2f_cli.log 2>&1
tail -3 gpurun_out/r2f_cli.log
