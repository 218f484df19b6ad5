"""Short density optimisation (n_maxiter outer iterations) for a launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/denopt_profile.py 256 6"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from denopt_bench import supercell
from profess_ad_b200.system import System
import profess_ad_b200.functionals as F

grid, maxiter = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device('cuda:0')
box, frac = supercell(4)
pot = os.path.join(ROOT, 'tests', 'potentials', 'al.gga.recpot')
terms = [F.IonElectron, F.Hartree, F.WangGovindCarter99().forward, F.PerdewZunger]
t0 = time.perf_counter()
s = System(box, (grid,) * 3, [['Al', pot, frac]], terms, units='b', coord_type='fractional', device=dev)
torch.cuda.synchronize()
print('setup', time.perf_counter() - t0)
t0 = time.perf_counter()
s.optimize_density(ntol=1e-7, n_method='LBFGS', from_uniform=True, n_maxiter=maxiter)
torch.cuda.synchronize()
print('denopt', time.perf_counter() - t0, s.last_optimization.get('closures'))
t0 = time.perf_counter()
f = s.forces()
torch.cuda.synchronize()
print('forces', time.perf_counter() - t0, f.abs().max().item())
