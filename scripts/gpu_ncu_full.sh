mkdir -p gpurun_out
PAD_FAST_FFT=1 ncu --set full --clock-control none --import-source on -k regex:'zinv_kernel|zfwd_kernel|xmix_kernel' -s 21 -c 7 -f -o gpurun_out/prof_own2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_own_full.log 2>&1
tail -3 gpurun_out/ncu_own_full.log
ncu -i gpurun_out/prof_own2.ncu-rep --page raw --csv > gpurun_out/prof_own2_raw.csv
python profiles/ncu_summary.py gpurun_out/prof_own2_raw.csv
