"""Correctness of the slab path over real NCCL (torchrun, >= 2 GPUs): energies / potentials of several functionals
and one density optimisation against the same computation on a single GPU.  Short timeouts everywhere."""
import datetime, json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from denopt_bench import supercell
import profess_ad_b200.functionals as F
from profess_ad_b200 import parallel, ion_utils as IU, _density_opt as D
from profess_ad_b200.synthetic import smooth_supercell

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
os.environ.setdefault('NCCL_DEBUG_FILE', '/tmp/nccl.%h.%p.log')
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=90))
# rough density on a skewed cell (same seed on every rank): no symmetry between the slabs
gen = torch.Generator().manual_seed(5)
box_h = 9.0 * torch.eye(3, dtype=torch.double) + 0.3 * torch.rand(3, 3, dtype=torch.double, generator=gen)
den_h = 0.03 * (1 + 0.5 * torch.rand(grid, grid, grid, dtype=torch.double, generator=gen))
box, den_full = box_h.to(dev), den_h.to(dev)
lo, hi = parallel.slab_bounds(grid, rank, world)
funcs = [('WGC99', F.WangGovindCarter99().forward), ('PBE', F.PerdewBurkeErnzerhof), ('Hartree', F.Hartree), ('WT', F.WangTeter)]
ok = True
for overlap in (False, True):
    with parallel.slab((grid,) * 3, overlap=overlap):
        for name, f in funcs:
            d = den_full[lo:hi].contiguous().requires_grad_(True)
            E = f(box, d)
            (g,) = torch.autograd.grad(E, d)
            d1 = den_full.clone().requires_grad_(True)
            saved, parallel._state.ctx = parallel._state.ctx, None      # plain single-GPU plan (3-D cuFFT)
            try:
                E1 = f(box, d1)
                (g1,) = torch.autograd.grad(E1, d1)
            finally:
                parallel._state.ctx = saved
            dE = abs(E.item() - E1.item()) / abs(E1.item())
            dv = ((g - g1[lo:hi]).abs().max() / g1.abs().max()).item()
            good = dE < 1e-12 and dv < 1e-11
            ok &= good
            if rank == 0:
                print(f'overlap={overlap} {name}: E {E.item():.12f} rel dE {dE:.1e} rel dv {dv:.1e} {"ok" if good else "MISMATCH"}', flush=True)
# density optimisation on slabs vs one GPU
boxs, frac = supercell(2)
frac = (frac + 0.03 * torch.rand(frac.shape, dtype=torch.double, generator=gen)) % 1.0      # break the symmetry
boxs = boxs.to(dev)
pot = os.path.join(ROOT, 'tests', 'potentials', 'al.gga.recpot')
shape = (grid,) * 3
n_elec = 3.0 * frac.shape[0]
terms = [F.IonElectron, F.Hartree, F.WangGovindCarter99().forward, F.PerdewZunger]
vol = abs(torch.linalg.det(boxs).item())
with parallel.slab(shape) as ctx:
    v_loc = IU.ionic_potential(boxs, ctx.local_shape, [(pot, frac.to(dev))])
    den = torch.full(ctx.local_shape, n_elec / vol, dtype=torch.double, device=dev)
    res, trace = parallel.optimize_density(boxs, den, v_loc, terms, n_elec, ntol=1e-7)
    forces = IU.ion_electron_forces(boxs, den, [(pot, frac.to(dev))])
v = IU.ionic_potential(boxs, shape, [(pot, frac.to(dev))])
den1 = torch.full(shape, n_elec / vol, dtype=torch.double, device=dev)
r1, _ = D.run(boxs, den1, v, D.describe_terms(terms), n_elec, 1e-7, 3, 'LBFGS', 0.1, 1000, 'dE')
dv = ((v_loc - v[lo:hi]).abs().max() / v.abs().max()).item()
dn = (den - den1[lo:hi]).abs().max().item()
good = abs(res['energy'] - r1['energy']) < 1e-8 and dv < 1e-11 and res['converged']
ok &= good
if rank == 0:
    print(f'denopt slab: E {res["energy"]:.10f} it {res["iterations"]} closures {res["closures"]} | single: E {r1["energy"]:.10f} it {r1["iterations"]} '
          f'| v_ext rel {dv:.1e} max|dn| {dn:.1e} max|F| {forces.abs().max().item():.2e} {"ok" if good else "MISMATCH"}', flush=True)
    print('ALL OK' if ok else 'FAILED', flush=True)
dist.barrier()
dist.destroy_process_group()
