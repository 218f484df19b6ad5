mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python bench.py > gpurun_out/bench_cufft.json 2> gpurun_out/bench_cufft.err
PAD_FAST_FFT=1 python bench.py --no-cpu-baseline > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
cat gpurun_out/bench_cufft.json gpurun_out/bench_fast.json
PAD_FAST_FFT=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fast.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fast.log 2>&1
PAD_FAST_FFT=1 ncu --set full --clock-control none --import-source on -k regex:'zinv_kernel|zfwd_kernel|spass_kernel|xmix_kernel' -s 40 -c 10 -o gpurun_out/prof_own python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_own_full.log 2>&1
ls -la gpurun_out
