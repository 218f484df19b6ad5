mkdir -p gpurun_out
for v in 0 1 2 3 4; do
  PAD_FAST_FFT=1 PAD_XMIX_VARIANT=$v python bench.py --steps 60 --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_v$v.json'))
    print('variant', $v, 'evals/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3))
    for k in d['roofline']['kernels']:
        if k['stage'].startswith('x-fwd * kernel'): print('   %-45s x%-3d %8.1f us  %s GB/s' % (k['stage'], k['launches_per_eval'], k['ms_per_eval']*1e3, round(k['GBps']) if k['GBps'] else None))
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_v$v.err').read()[-2000:])
PY
done
