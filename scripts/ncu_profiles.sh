#!/bin/bash
# Round-2 final profiles: (1) ncu launch list of the bench command, (2) ncu --set full of the 12 kernels of one WGC99 evaluation,
# (3) launch lists of the fused term list / density optimisation and of revHC at 128^3.  CUDA graphs off (one launch per kernel).
mkdir -p gpurun_out
export PAD_GRAPHS=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-denopt > gpurun_out/r02_ncu_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_bench_summary.txt 2>&1
head -30 gpurun_out/r02_launches_bench_summary.txt
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:'zinv_kernel|zfwd_kernel|xmix_kernel|spass_kernel|wgc_sum_scalars' -s 36 -c 12 \
   -o /tmp/r02_eval_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-denopt > gpurun_out/r02_ncu_full.log 2>&1
ls -la /tmp/r02_eval_full.ncu-rep
ncu -i /tmp/r02_eval_full.ncu-rep --page raw --csv > gpurun_out/r02_eval_full_raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/r02_eval_full_raw.csv > gpurun_out/r02_eval_full_summary.md 2>&1
cat gpurun_out/r02_eval_full_summary.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_denopt.csv \
    python scripts/denopt_profile.py 256 6 > gpurun_out/r02_ncu_denopt.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_launches_denopt.csv > gpurun_out/r02_launches_denopt_summary.txt 2>&1
head -30 gpurun_out/r02_launches_denopt_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_revhc.csv \
    python scripts/hc_profile.py 128 > gpurun_out/r02_ncu_revhc.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_launches_revhc.csv > gpurun_out/r02_launches_revhc_summary.txt 2>&1
head -30 gpurun_out/r02_launches_revhc_summary.txt
gzip -f gpurun_out/r02_launches_denopt.csv
