"""Density optimisation of ONE grid spread over the ranks (torchrun, NCCL): 256-atom Al supercell, IonElectron +
Hartree + WGC99 + PZ, slab-decomposed; compared with the same optimisation on one GPU (rank 0)."""
import json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from denopt_bench import supercell
import profess_ad_b200.functionals as F
from profess_ad_b200 import parallel, ion_utils as IU, _density_opt as D

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
os.environ['NCCL_DEBUG'] = 'WARN'
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
box, frac = supercell(4)
box = box.to(dev)
pot = os.path.join(ROOT, 'tests', 'potentials', 'al.gga.recpot')
shape = (grid,) * 3
n_elec = 3.0 * frac.shape[0]
terms = [F.IonElectron, F.Hartree, F.WangGovindCarter99().forward, F.PerdewZunger]
vol = abs(torch.linalg.det(box).item())
with parallel.slab(shape) as ctx:
    t0 = time.perf_counter()
    v_loc = IU.ionic_potential(box, ctx.local_shape, [(pot, frac.to(dev))])
    torch.cuda.synchronize()
    t_vext = time.perf_counter() - t0
    out = None
    for rep in range(2):
        den = torch.full(ctx.local_shape, n_elec / vol, dtype=torch.double, device=dev)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        res, trace = parallel.optimize_density(box, den, v_loc, terms, n_elec, ntol=1e-7)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
    forces = IU.ion_electron_forces(box, den, [(pot, frac.to(dev))])
    out = {'grid': grid, 'world': world, 'seconds': dt, 'vext_s': t_vext, 'iterations': res['iterations'], 'closures': res['closures'],
           'converged': res['converged'], 'energy_Ha': res['energy'], 'max_force': forces.abs().max().item()}
if rank == 0:
    print(json.dumps(out), flush=True)
    # the same optimisation on one GPU
    v = IU.ionic_potential(box, shape, [(pot, frac.to(dev))])
    den1 = torch.full(shape, n_elec / vol, dtype=torch.double, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r1, _ = D.run(box, den1, v, D.describe_terms(terms), n_elec, 1e-7, 3, 'LBFGS', 0.1, 1000, 'dE')
    torch.cuda.synchronize()
    print(json.dumps({'single_gpu_seconds': time.perf_counter() - t0, 'energy_Ha': r1['energy'], 'iterations': r1['iterations'],
                      'dE_vs_slab_Ha': r1['energy'] - res['energy']}), flush=True)
dist.barrier()
dist.destroy_process_group()
