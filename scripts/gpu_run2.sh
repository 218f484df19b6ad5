mkdir -p gpurun_out
python -m pytest tests/test_gpu_fastfft.py -m gpu -x -q > gpurun_out/pytest_fast.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fast.log
tail -5 gpurun_out/pytest_fast.log
for zg in 0 1 2 4; do
  PAD_FAST_FFT=1 PAD_ZGROUP=$zg python bench.py --no-cpu-baseline > gpurun_out/bench_fast_zg$zg.json 2> gpurun_out/bench_fast_zg$zg.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_fast_zg$zg.json'))
print('zgroup', $zg, 'evals/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
PY
done
PAD_FAST_FFT=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_own.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_own.log 2>&1
PAD_FAST_FFT=1 ncu --set full --clock-control none --import-source on -k regex:'zinv_kernel|zfwd_kernel|spass_kernel|xmix_kernel' -s 20 -c 12 -o gpurun_out/prof_own python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_own_full.log 2>&1
ls -la gpurun_out
