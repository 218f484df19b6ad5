"""Stage times (CUDA events on rank 0's stream) of one slab-decomposed WGC99 evaluation.
    torchrun --nproc-per-node N scripts/slab_stage_profile.py [grid]"""
import ctypes, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import profess_ad_b200.functionals as F
from profess_ad_b200 import parallel, _native
from profess_ad_b200.synthetic import smooth_supercell
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lo, hi = parallel.slab_bounds(n, rank, world)
box, den = smooth_supercell(n, max(1, n // 64), device=dev, x_range=(lo, hi))
lib = _native.load_library()
wgc = F.WangGovindCarter99()
with parallel.slab((n, n, n)):
    def step():
        d = den.requires_grad_(True)
        E = wgc.forward(box, d)
        (g,) = torch.autograd.grad(E, d)
        den.requires_grad_(False)
        return E
    for _ in range(3):
        step()
    torch.cuda.synchronize(); dist.barrier()
    lib.pad_profile_begin()
    for _ in range(5):
        E = step()
    torch.cuda.synchronize()
    names = ctypes.create_string_buffer(48 * 256)
    ms = (ctypes.c_double * 256)()
    n_st, n_ev = ctypes.c_int(0), ctypes.c_int(0)
    lib.pad_profile_end(names, ms, 256, ctypes.byref(n_st), ctypes.byref(n_ev))
    if rank == 0:
        tot = 0.0
        print('WGC99 E+V', n, 'on', world, 'GPUs, E =', E.item())
        for i in range(n_st.value):
            nm = names.raw[48 * i:48 * i + 48].split(b'\0')[0].decode()
            print('    %-62s %8.1f us' % (nm, 1e3 * ms[i]))
            tot += ms[i]
        print('    %-62s %8.1f us' % ('sum', 1e3 * tot))
dist.destroy_process_group()
