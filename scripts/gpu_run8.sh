mkdir -p gpurun_out
python -m pytest tests/test_gpu_fastfft.py -m gpu -x -q > gpurun_out/pytest_fast.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fast.log
tail -3 gpurun_out/pytest_fast.log
for zc in 8 16; do
  PAD_FAST_FFT=1 PAD_SPASS_ZC=$zc python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_fast_zc$zc.json 2> gpurun_out/bench_fast_zc$zc.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_fast_zc$zc.json'))
    print('spass_zc', $zc, 'evals/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'])
    for k in d['roofline']['kernels']: print('   %-45s x%-3d %8.1f us  %s GB/s' % (k['stage'], k['launches_per_eval'], k['ms_per_eval']*1e3, round(k['GBps']) if k['GBps'] else None))
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench_fast_zc$zc.err').read()[-2000:])
PY
done
