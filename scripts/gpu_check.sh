# one GPU call: full GPU test suite, both bench arms, launch list, ncu --set full of one whole evaluation
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 400 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('evals/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'], 'frac', round(d['roofline']['frac'],3), d.get('cpu_baseline'))
for k in d['roofline']['kernels']: print('   %-45s x%-3d %8.1f us  %s GB/s' % (k['stage'], k['launches_per_eval'], k['ms_per_eval']*1e3, round(k['GBps']) if k['GBps'] else None))
print(d.get('density_optimization'))
print(open('gpurun_out/bench_ref.json').read()[:300])
PY
if [ "$1" = "ncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-denopt > gpurun_out/launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches.csv | head -20
ncu --set full --clock-control none --import-source on -k regex:'zinv_kernel|zfwd_kernel|xmix_kernel|spass_kernel' -s 44 -c 11 -f -o gpurun_out/prof_eval python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-denopt > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_eval.ncu-rep --page raw --csv > gpurun_out/prof_eval_raw.csv
python profiles/ncu_summary.py gpurun_out/prof_eval_raw.csv
fi
