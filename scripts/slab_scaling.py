"""Several strong-scaling measurements in ONE torchrun job (one NCCL initialisation):
    torchrun --nproc-per-node 8 scripts/slab_scaling.py 512:wgc99:0 512:wgc99:1 1024:wgc99:1 512:revhc:1 1024:pbe:1
Each argument is grid:functional:overlap[:steps]; rank 0 prints one JSON line per measurement."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch.distributed as dist

rank, world = bench.slab_init()
for spec in sys.argv[1:]:
    parts = spec.split(':')
    n, fun, ov = int(parts[0]), parts[1], parts[2] != '0'
    steps = int(parts[3]) if len(parts) > 3 else 6
    try:
        line = bench.slab_measure(n, fun, steps, 3, overlap=ov)
    except Exception as e:      # noqa: BLE001 -- keep going with the remaining measurements
        line = {'spec': spec, 'error': repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    import torch
    from profess_ad_b200 import _native
    torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
