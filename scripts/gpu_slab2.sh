# 2 GPUs: slab strong scaling with and without the pipelined exchange, independent systems, slab density optimisation over NCCL
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for ov in 0 1; do
  PAD_SLAB_OVERLAP=$ov run 29601 bench.py --gpus 2 --slab-grid 512 --steps 10 --warmup 3 > gpurun_out/slab512_n2_ov$ov.json 2> gpurun_out/slab512_n2_ov$ov.err
  python -c "import json;d=json.load(open('gpurun_out/slab512_n2_ov$ov.json'));print('overlap $ov: 512^3 on 2 GPUs', round(d['ms_per_step'],2),'ms', d['config']['energy_Ha'])" || tail -5 gpurun_out/slab512_n2_ov$ov.err
done
PAD_SLAB_OVERLAP=1 run 29603 bench.py --gpus 2 --slab-grid 256 --steps 20 --warmup 3 > gpurun_out/slab256_n2.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/slab256_n2.json'));print('256^3 on 2 GPUs', round(d['ms_per_step'],2),'ms', d['config']['energy_Ha'])"
run 29605 scripts/slab_denopt.py 128 > gpurun_out/slab_denopt_n2.log 2>&1; tail -4 gpurun_out/slab_denopt_n2.log
run 29607 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_n2.json'));print('independent systems on 2 GPUs', round(d['value'],1),'evals/s')"
