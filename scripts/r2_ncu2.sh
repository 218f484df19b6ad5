#!/bin/bash
# ncu --set full with source counters: forward z kernel (4 fields) and the mid inverse z kernel only (report < 64 MiB)
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'zinv_kernel|zfwd_kernel' -s 8 -c 2 \
   -o gpurun_out/r2_z_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-denopt > gpurun_out/r2_ncu2.log 2>&1
ls -la gpurun_out/r2_z_full.ncu-rep
ncu -i gpurun_out/r2_z_full.ncu-rep --page raw --csv > gpurun_out/r2_z_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_z_full.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_z_full_sass.csv 2>/dev/null
ls -la gpurun_out/
