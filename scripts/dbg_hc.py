import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import profess_ad_b200.functionals as F
from profess_ad_b200.system import System
from profess_ad_b200 import _density_opt as D
from oracle import ofdft_oracle as orc
EV = 27.211386245988
g = np.load(os.path.join(ROOT, 'tests/golden/denopt_al_fcc4_config1.npz'))
tab = np.load(os.path.join(ROOT, 'tests/golden/hc_table.npz'))
t_rev = torch.from_numpy(tab['revhc'])
box = torch.from_numpy(g['box_bohr']); n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
shape = (n, n, n); frac = torch.from_numpy(g['frac'])
pot = os.path.join(ROOT, 'tests/potentials/al.gga.recpot')
dev = torch.device('cuda:0')
ohc = orc.RevisedHuangCarter(0.45, 0.10, 2 / 3, 1.15, kernel=t_rev)
oterms = [orc.IonElectron, orc.Hartree, ohc, orc.PerdewZunger]
for native in (True, False):
    System.use_native_optimizer = native
    hc = F.RevisedHuangCarter((0.45, 0.10, 2.0 / 3.0, 1.15), kernel=t_rev.clone())
    terms = [F.IonElectron, F.Hartree, hc.forward, F.PerdewZunger]
    s = System(box, shape, [['Al', pot, frac]], terms, units='b', coord_type='fractional')
    s.optimize_density(ntol=1e-7, from_uniform=True)
    lo = s.last_optimization
    print('native' if native else 'host', s.energy('Ha'), lo.get('iterations'), lo.get('closures'), lo.get('converged'))
    if native:
        print(' trace E', [round(r[0], 7) for r in lo['trace'].tolist()])
    den = s.density().cpu(); v_ext = s.ionic_potential().cpu()
    print('  oracle energy at this density', float(orc.total_energy(box, den, oterms, v_ext)), ' n_nodes', hc.last_n_nodes if hasattr(hc, 'last_n_nodes') else None)
n_elec = 12.0
den0 = torch.full(shape, n_elec / abs(torch.linalg.det(box).item()), dtype=torch.double)
_step = orc.LbfgsState.step
import math
def dbg_closure_wrap(closure, st):
    def c(x):
        E, g = closure(x)
        print('[oracle] closures %d E %.12f |g|1 %.6e gg %.6e k %d H %.6e' % (st.closures + 1, float(E), float(g.abs().sum()), float(g.dot(g)), len(getattr(st, 'Y', [])), getattr(st, 'H', 1.0)), file=sys.stderr)
        return E, g
    return c
def step_dbg(self, closure):
    return _step(self, dbg_closure_wrap(closure, self))
ref = orc.optimize_density(box, den0, n_elec, oterms, v_ext=v_ext, ntol=1e-7)
print('oracle', ref['energy'], ref['iterations'], ref['closures'])
print(' trace E', [round(e, 7) for e in ref['trace']])
# fused evaluator vs oracle at the oracle's optimum and a perturbed density
for d in (ref['den'], ref['den'] * (1 + 0.05 * torch.rand(shape, dtype=torch.double))):
    T = D.describe_terms(terms)
    E, v = D.eval_total(box.to(dev), d.to(dev).contiguous(), v_ext.to(dev), T)
    Eo = orc.total_energy(box, d, oterms, v_ext)
    dd = d.clone().requires_grad_(True)
    (go,) = torch.autograd.grad(orc.total_energy(box, dd, oterms, v_ext), dd)
    dV = abs(torch.linalg.det(box).item()) / d.numel()
    print('eval_total dE', E.item() - float(Eo), 'dv rel', ((v.cpu() - go / dV).abs().max() / (go / dV).abs().max()).item())
