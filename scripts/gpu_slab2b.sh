mkdir -p gpurun_out
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29611 scripts/slab_check.py 64 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -14
for ov in 0 1; do
  PAD_SLAB_OVERLAP=$ov run 2962$ov bench.py --gpus 2 --slab-grid 256 --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/slab256_n2_ov$ov.json
  python -c "import json;d=json.load(open('gpurun_out/slab256_n2_ov$ov.json'));print('overlap $ov: 256^3 on 2 GPUs', round(d['ms_per_step'],3),'ms', repr(d['config']['energy_Ha']))"
done
