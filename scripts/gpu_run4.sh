mkdir -p gpurun_out
python -m pytest tests/test_gpu_fastfft.py -m gpu -x -q > gpurun_out/pytest_fast.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fast.log
tail -5 gpurun_out/pytest_fast.log
for zg in 0 1 2 4; do
  PAD_FAST_FFT=1 PAD_ZGROUP=$zg python bench.py --no-cpu-baseline > gpurun_out/bench_fast_zg$zg.json 2> gpurun_out/bench_fast_zg$zg.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_fast_zg$zg.json'))
    print('zgroup', $zg, 'evals/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
except Exception as e:
    print('zgroup', $zg, 'failed', e); print(open('gpurun_out/bench_fast_zg$zg.err').read()[-2000:])
PY
done
PAD_FAST_FFT=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_own.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_own.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_own.csv | head -24
