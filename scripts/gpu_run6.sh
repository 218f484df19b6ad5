mkdir -p gpurun_out
for zg in 1 2; do
PAD_FAST_FFT=1 PAD_ZGROUP=$zg ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -k regex:'spass|xmix' -s 300 -c 80 --csv --log-file gpurun_out/l2_zg$zg.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
done
python - <<'PY'
import csv
for zg in (1,2):
    rows=list(csv.reader(open(f'gpurun_out/l2_zg{zg}.csv')))
    for i,r in enumerate(rows):
        if 'Kernel Name' in r: h=i;break
    hdr=rows[h]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
    d={}
    for r in rows[h+1:]:
        d.setdefault((int(r[ii]), r[ki][:40]),{})[r[mi]]=float(r[vi].replace(',',''))
    print('zgroup',zg)
    for (i,k),m in sorted(d.items())[:16]:
        print(f"  {k:42s} {m['gpu__time_duration.sum']/1000:8.1f} us  rd {m['dram__bytes_read.sum']/1e6:8.1f} MB  wr {m['dram__bytes_write.sum']/1e6:8.1f} MB  L2hit {m['lts__t_sector_hit_rate.pct']:5.1f}%")
PY
