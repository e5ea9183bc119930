#!/bin/bash
mkdir -p gpurun_out /tmp/e
python - <<'PY'
import sys
sys.path.insert(0,'tests')
import _hostcases as H
paths,args=H.make_inputs('E','/tmp/e')
open('/tmp/e/args','w').write(' '.join(paths+args))
open('/tmp/e/golden.clstr','wb').write(H.read_golden('E'))
PY
md5sum /tmp/e/golden.clstr
for i in 1 2 3 4 5 6 7 8; do
  timeout 120 bin/meshclust $(cat /tmp/e/args) --output /tmp/e/out$i.clstr > /tmp/e/log$i 2>&1
  echo "run $i rc=$? $(md5sum < /tmp/e/out$i.clstr) $(grep Accumulation /tmp/e/log$i)"
done
MC_PHASE_A_STEPS=1 timeout 120 bin/meshclust $(cat /tmp/e/args) --output /tmp/e/outs.clstr > /tmp/e/logs 2>&1
echo "steps rc=$? $(md5sum < /tmp/e/outs.clstr) $(grep Accumulation /tmp/e/logs)"
for g in 1 2 8 37; do
MC_PA_GRID=$g timeout 120 bin/meshclust $(cat /tmp/e/args) --output /tmp/e/outg.clstr > /tmp/e/logg 2>&1
echo "grid $g rc=$? $(md5sum < /tmp/e/outg.clstr) $(grep Accumulation /tmp/e/logg)"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "accumulate_run" 2>&1 | tail -5
