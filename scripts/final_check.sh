#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2z_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r2z_pytest.log | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2z_smoke.log
( time timeout 1200 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'steps', 'warmup', 'gpu_launches', 'clocks', 'launch_detail')})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['traffic'], 'cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('kind'))
print('parity', d.get('parity'))
for x in d.get('density_optimization', []):
    print('  denopt', x['grid'], x['seconds'], x.get('cpu_reference', {}).get('seconds'), x.get('dE_eV_per_atom_vs_reference'))
PY
( time timeout 1200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err ) 2>&1 | grep real
cat gpurun_out/r2z_ref.json | cut -c1-600
