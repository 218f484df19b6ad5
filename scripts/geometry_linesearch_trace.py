"""Line-search internals of the Li2 (ions only) geometry optimisation: native ion-ion sum vs the torch pair list."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import profess_ad_b200.functionals as F
from profess_ad_b200.system import System
from profess_ad_b200 import ion_utils as IU
import profess_ad_b200.system as SYS
from profess_ad_b200._optimizers.lbfgs import lbfgsnew as LB

mode = sys.argv[1]
g = np.load('tests/golden/geometry_li2_ions.npz')
terms = [F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof]

def torch_path(box, coords, charges, Rc, Rd):
    mi, mj, shifts = IU._pair_list(box, coords, Rc)
    rho = torch.sum(charges) / torch.abs(torch.linalg.det(box))
    Zi, Zj = charges[mi], charges[mj]
    Qi = torch.scatter_add(charges, 0, mi, Zj)
    aux = (0.75 / np.pi) * Qi / rho
    Ra = aux.sign() * aux.abs().pow(1 / 3)
    r_ij = (coords[mj] + shifts @ box - coords[mi]).norm(p=2, dim=1)
    E_local = torch.sum(0.5 * Zi * Zj * torch.erfc(r_ij / Rd) / r_ij)
    E_corr = torch.sum(-np.pi * charges * rho * Ra.square() + np.pi * charges * rho * (Ra.square() - 0.5 * Rd * Rd) * torch.erf(Ra / Rd)
                       + np.sqrt(np.pi) * charges * rho * Ra * Rd * torch.exp(-Ra.square() / (Rd * Rd)) - charges.square() / np.sqrt(np.pi) / Rd)
    return E_local + E_corr
if mode == 'torch':
    SYS.ion_interaction_sum = torch_path

orig_phi, orig_run = LB._WolfeSearch.phi, LB._WolfeSearch.run
def phi(self, alpha):
    v = orig_phi(self, alpha)
    print('    phi(%.12g) = %.17g' % (alpha, v), flush=True)
    return v
def run(self, h):
    print('  line search: |d|_1 = %.6g' % float(self.d.abs().sum()), flush=True)
    t = orig_run(self, h)
    print('  -> t = %.12g  evals %d' % (t, self.evals), flush=True)
    return t
LB._WolfeSearch.phi, LB._WolfeSearch.run = phi, run
ions = [['Li', 'tests/potentials/li.gga.recpot', torch.from_numpy(g['frac0'])]]
s = System(torch.from_numpy(g['box0_A']), tuple(int(n) for n in g['shape']), ions, terms, units='a', coord_type='fractional')
ok = s.optimize_geometry(g_maxiter=12, ntol=1e-9, ftol=0.02, stol=None, g_verbose=True)
print('ok', ok, s.last_geometry_optimization)
