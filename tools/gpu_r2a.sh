#!/bin/bash
# round 2, first GPU pass: the persistent Phase-A kernel (parity + step timing)
mkdir -p gpurun_out
export MC_DEBUG_TIMING=
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "accumulate_run" > gpurun_out/r2a_tests.log 2>&1
echo "pytest accumulate_run rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout 1200 python -m pytest tests/test_host_logic.py -x -q -m gpu -k "identical_to_reference_gpu" > gpurun_out/r2a_cli.log 2>&1
echo "pytest cli rc=$?" >> gpurun_out/r2a_cli.log
tail -5 gpurun_out/r2a_cli.log
python tools/gen_config.py c2 /tmp/c2.fa > /dev/null
for i in 1 2; do
MC_PA_TRACE=4000 timeout 300 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2.clstr > gpurun_out/r2a_c2_run$i.log 2>&1
echo "c2 rc=$?" >> gpurun_out/r2a_c2_run$i.log
done
grep -E "Accumulation|trace|Total|rc=" gpurun_out/r2a_c2_run2.log
md5sum /tmp/c2.clstr
MC_PHASE_A_STEPS=1 timeout 300 bin/meshclust /tmp/c2.fa --id 0.97 --kmer 4 --output /tmp/c2s.clstr > gpurun_out/r2a_c2_steps.log 2>&1
grep -E "Accumulation|Total" gpurun_out/r2a_c2_steps.log
md5sum /tmp/c2s.clstr
