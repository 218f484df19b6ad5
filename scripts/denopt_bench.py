"""Seconds per density optimisation (BASELINE.json metric 2) on the B200: Al fcc supercells, WGC99 stack."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from profess_ad_b200.system import System
import profess_ad_b200.functionals as F

A = 4.05 / 0.529177210903
FCC = [[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]]


def supercell(side):
    box = side * A * torch.eye(3, dtype=torch.double)
    frac = []
    for i in range(side):
        for j in range(side):
            for k in range(side):
                for b in FCC:
                    frac.append([(i + b[0]) / side, (j + b[1]) / side, (k + b[2]) / side])
    return box, torch.tensor(frac, dtype=torch.double)


def run(side, grid):
    dev = torch.device('cuda:0')
    box, frac = supercell(side)
    pot = os.path.join(ROOT, 'tests', 'potentials', 'al.gga.recpot')
    terms = [F.IonElectron, F.Hartree, F.WangGovindCarter99().forward, F.PerdewZunger]
    t0 = time.perf_counter()
    s = System(box, (grid,) * 3, [['Al', pot, frac]], terms, units='b', coord_type='fractional', device=dev)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    out = None
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.optimize_density(ntol=1e-7, n_method='LBFGS', from_uniform=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out = {'atoms': 4 * side ** 3, 'grid': grid, 'seconds': dt, 'setup_s': t_setup,
               'energy_eV_per_atom': s.energy('eV') / (4 * side ** 3), 'info': {k: v for k, v in s.last_optimization.items() if isinstance(v, (int, float, bool, str))}}
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    for side, grid in ((1, 32), (2, 64), (4, 128), (4, 256)):
        run(side, grid)
