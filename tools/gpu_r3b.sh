#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3b_ref.json 2> gpurun_out/r3b_ref.err; echo "ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r3b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']['frac'], d['clocks']); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['one_upload_per_10_scans']['value']); print(d.get('phase_a_in_product')); print(d['seqs_clustered']['wall_s'], d['seqs_clustered']['value'], d['seqs_clustered']['stages'])
r=json.loads(open('gpurun_out/r3b_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r.get('cpu_baseline'))
"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
