#!/bin/bash
# ncu --set full with source counters: forward z kernel (4 fields) and the mid inverse z kernel only; the report itself is too
# large to travel back (> 64 MiB with the embedded cubin), so the raw and SASS-level pages are exported here and it is removed
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'zinv_kernel|zfwd_kernel' -s 8 -c 2 \
   -o /tmp/r2_z_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-denopt > gpurun_out/r2_ncu2.log 2>&1
ncu -i /tmp/r2_z_full.ncu-rep --page raw --csv > gpurun_out/r2_z_full_raw.csv 2>/dev/null
ncu -i /tmp/r2_z_full.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_z_full_sass.csv 2>/dev/null
gzip -f gpurun_out/r2_z_full_sass.csv
ls -la gpurun_out/
