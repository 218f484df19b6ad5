mkdir -p gpurun_out
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json; tail -4 gpurun_out/bench_ref.err
( time python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1200 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
print('evals/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'], d['clocks'], 'frac', round(d['roofline']['frac'],3))
print(d.get('cpu_baseline')); print(d.get('density_optimization')); print(d['roofline']['dominant_kernel'])
PY
