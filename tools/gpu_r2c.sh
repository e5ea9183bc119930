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
for i in 1 2 3 4 5 6; do
  rm -f /tmp/e/out.clstr
  timeout 120 bin/meshclust $(cat /tmp/e/args) --output /tmp/e/out.clstr > /tmp/e/stdout 2> /tmp/e/stderr
  echo "run $i rc=$? md5=$(md5sum < /tmp/e/out.clstr 2>/dev/null) | $(grep Accumulation /tmp/e/stdout) | stderr: $(tail -2 /tmp/e/stderr)"
done
for g in 38 37 30 20 12; do
  rm -f /tmp/e/out.clstr
  MC_PA_GRID=$g timeout 120 bin/meshclust $(cat /tmp/e/args) --output /tmp/e/out.clstr > /tmp/e/stdout 2> /tmp/e/stderr
  echo "grid $g rc=$? md5=$(md5sum < /tmp/e/out.clstr 2>/dev/null) | $(grep Accumulation /tmp/e/stdout) | stderr: $(tail -2 /tmp/e/stderr)"
done
