#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "accumulate_run" 2>&1 | tail -2
python tools/gen_config.py c2 /tmp/c2.fa > /dev/null
MC_PA_TRACE=4000 timeout 300 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr > gpurun_out/r2i_c2.log 2>&1
echo "c2 rc=$? $(md5sum < /tmp/c2.clstr)"
grep -E "Accumulation|trace" gpurun_out/r2i_c2.log
python tools/gen_config.py c4 /tmp/c4.fa > /dev/null
MC_PA_TRACE=4000 timeout 600 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4.clstr > gpurun_out/r2i_c4.log 2>&1
echo "c4 rc=$? $(md5sum < /tmp/c4.clstr)"
grep -E "Accumulation|trace|Total|\[" gpurun_out/r2i_c4.log
python tools/gen_config.py c5 /tmp/c5.fa > /dev/null
MC_PA_TRACE=4000 timeout 600 bin/meshclust /tmp/c5.fa --kmer 6 --output /tmp/c5.clstr > gpurun_out/r2i_c5.log 2>&1
echo "c5 rc=$? $(md5sum < /tmp/c5.clstr)"
grep -E "Accumulation|trace|Total|\[" gpurun_out/r2i_c5.log
