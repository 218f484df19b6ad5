mkdir -p gpurun_out
nvidia-smi -L
for grid in 256 512; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 --slab-grid $grid > gpurun_out/slab_${grid}_n2.json 2> gpurun_out/slab_${grid}_n2.err
tail -c 800 gpurun_out/slab_${grid}_n2.err | tail -5
cat gpurun_out/slab_${grid}_n2.json
done
python bench.py --gpus 1 --steps 10 --warmup 3 --slab-grid 512 > gpurun_out/slab_512_n1.json 2> gpurun_out/slab_512_n1.err; tail -3 gpurun_out/slab_512_n1.err; cat gpurun_out/slab_512_n1.json
