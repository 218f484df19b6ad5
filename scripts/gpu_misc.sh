mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/strain_scan.py 2>&1 | tail -2
for f in revhc pbe; do
python bench.py --gpus 1 --slab-grid 128 --slab-functional $f --steps 5 --warmup 3 2> gpurun_out/slab_$f.err | tail -1 > gpurun_out/slab128_$f.json
python -c "import json;d=json.load(open('gpurun_out/slab128_$f.json'));print('$f 128^3 slab world 1:', round(d['ms_per_step'],3),'ms', d['config']['n_fft'], repr(d['config']['energy_Ha']))" || tail -3 gpurun_out/slab_$f.err
done
