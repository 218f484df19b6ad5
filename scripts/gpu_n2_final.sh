mkdir -p gpurun_out
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29701 scripts/strain_scan.py 2>/dev/null | tail -1 | tee gpurun_out/strain_scan_n2.json
run 29702 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.out 2> gpurun_out/bench_n2.err
echo "stdout lines: $(wc -l < gpurun_out/bench_n2.out)"; head -c 250 gpurun_out/bench_n2.out; echo
run 29703 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.out 2>/dev/null; echo "ref stdout lines: $(wc -l < gpurun_out/bench_ref_n2.out)"; head -c 200 gpurun_out/bench_ref_n2.out; echo
