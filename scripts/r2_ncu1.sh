#!/bin/bash
# ncu --set full with source counters for the four kernel families of one WGC99 evaluation (batch zinv, unfused mid)
mkdir -p gpurun_out
export PAD_ZINV_STREAM=0 PAD_FUSE_MID=0
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'zinv_kernel|zfwd_kernel|xmix_kernel|spass_kernel' -s 44 -c 11 \
   -o gpurun_out/r2_eval_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-denopt > gpurun_out/r2_ncu1.log 2>&1
ls -la gpurun_out/r2_eval_full.ncu-rep
ncu -i gpurun_out/r2_eval_full.ncu-rep --page raw --csv > gpurun_out/r2_eval_full_raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/r2_eval_full_raw.csv 2>/dev/null | head -30
