#!/bin/bash
# round 2, first GPU contact of the pipelined (z, y) kernels: parity, then timing against the per-pass path
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_fastfft.py -x -q -m gpu > gpurun_out/r2_pytest_fastfft.log 2>&1
echo "pytest fastfft rc=$?" >> gpurun_out/r2_pytest_fastfft.log
tail -5 gpurun_out/r2_pytest_fastfft.log
for cfg in "1 0 0" "0 0 0" "1 8 2" "1 8 4" "1 32 4" "1 16 8" "1 32 17" "1 64 8"; do
  set -- $cfg
  PAD_PIPE=$1 PAD_PIPE_LPI=$2 PAD_PIPE_TPI=$3 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-denopt \
     > gpurun_out/r2_bench_pipe$1_lpi$2_tpi$3.json 2> gpurun_out/r2_bench_pipe$1_lpi$2_tpi$3.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_bench_pipe$1_lpi$2_tpi$3.json'))
    print('pipe=$1 lpi=$2 tpi=$3', 'ms/step', round(d['ms_per_step'], 4), 'E', d.get('config_detail', d['config']).get('energy_Ha'))
    for k in d['roofline']['kernels']:
        print('    %-70s %8.1f us' % (k['stage'], 1e3 * k['ms_per_eval']))
except Exception as e:
    print('pipe=$1 lpi=$2 tpi=$3 FAILED', e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none \
   -k regex:'zy_fwd|yz_inv|xmix|spass|zfwd|zinv' -s 30 -c 10 --csv --log-file gpurun_out/r2_ncu_traffic_pipe.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-denopt > gpurun_out/r2_ncu_traffic.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_ncu_traffic_pipe.csv')) if len(r) > 10]
hdr = rows[0]
i_k, i_m, i_v = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
i_id = hdr.index('ID')
agg = {}
for r in rows[1:]:
    agg.setdefault((r[i_id], r[i_k][:60]), {})[r[i_m]] = float(r[i_v].replace(',', ''))
tot = 0
for (i, k), m in agg.items():
    b = m.get('dram__bytes_read.sum', 0) + m.get('dram__bytes_write.sum', 0)
    tot += b
    print('%-62s %8.1f us  rd %7.1f MB  wr %7.1f MB  L2hit %5.1f' % (k, m.get('gpu__time_duration.sum', 0) / 1e3,
          m.get('dram__bytes_read.sum', 0) / 1e6, m.get('dram__bytes_write.sum', 0) / 1e6, m.get('lts__t_sector_hit_rate.pct', 0)))
print('total dram MB', tot / 1e6)
PY
