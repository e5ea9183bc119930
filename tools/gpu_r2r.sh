#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "accumulate_run" 2>&1 | tail -12
timeout 1500 python -m pytest tests/test_host_logic.py -q -m gpu -k "identical_to_reference_gpu" 2>&1 | tail -3
for cfg in c2 c4 c5; do python tools/gen_config.py $cfg /tmp/$cfg.fa > /dev/null; done
for i in 1 2; do
MC_PA_TRACE=4000 timeout 300 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr > gpurun_out/r2r_c2.log 2>&1; echo "c2 rc=$? $(md5sum < /tmp/c2.clstr)"; grep -E "Accumulation|trace, 2" gpurun_out/r2r_c2.log
done
MC_PA_NO_COMPACT=1 timeout 300 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr > gpurun_out/r2r_c2n.log 2>&1; echo "c2 nocompact rc=$? $(md5sum < /tmp/c2.clstr)"; grep -E "Accumulation" gpurun_out/r2r_c2n.log
timeout 600 bin/meshclust /tmp/c4.fa --id 0.90 --kmer 5 --output /tmp/c4.clstr > gpurun_out/r2r_c4.log 2>&1; echo "c4 rc=$? $(md5sum < /tmp/c4.clstr)"; grep -E "Accumulation|Total" gpurun_out/r2r_c4.log
timeout 600 bin/meshclust /tmp/c5.fa --kmer 6 --output /tmp/c5.clstr > gpurun_out/r2r_c5.log 2>&1; echo "c5 rc=$? $(md5sum < /tmp/c5.clstr)"; grep -E "Accumulation|Total" gpurun_out/r2r_c5.log
