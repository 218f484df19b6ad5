"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference (PROFESS-AD v1.0.1, /root/reference/src) is imported with ``sys.modules`` stubs for
the three packages that are not installed here (xitorch, torch_nl, matplotlib).  torch_nl is
replaced by a brute-force periodic pair list with the same (mapping, batch, shifts) convention;
xitorch.solve_ivp is never called because the Huang-Carter omega(eta) table is injected
(``hc.kernel`` is a plain attribute, functionals.py:1230) -- the HC rows are therefore pinned
*given the table*, not at the ODE boundary.

Outputs (all fp64):
  functionals_<case>.npz : box, den, v_ext, and for every functional  E_<name>, V_<name>
  hc_table.npz           : the injected omega(eta) table
  denopt_<case>.npz      : v_ext, final density/energy/iterations of System.optimize_density
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
POT = os.path.join(ROOT, 'tests', 'potentials')


def brute_force_neighborlist(cutoff, pos, cell, pbc, batch, self_interaction=False):
    """Same return convention as torch_nl.compute_neighborlist: r_ij = pos[j] + shifts @ cell - pos[i]."""
    cutoff = float(cutoff)
    inv = torch.linalg.inv(cell)
    heights = 1.0 / torch.sqrt(torch.sum(inv.T.pow(2), 1))     # interplanar spacings
    reps = [int(np.ceil(cutoff / h.item())) + 1 for h in heights]
    rng = [torch.arange(-r, r + 1, dtype=torch.double) for r in reps]
    S = torch.stack(torch.meshgrid(*rng, indexing='ij'), -1).reshape(-1, 3)
    n = pos.shape[0]
    ii, jj = torch.meshgrid(torch.arange(n), torch.arange(n), indexing='ij')
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    disp = (pos[jj] - pos[ii]).unsqueeze(1) + (S @ cell).unsqueeze(0)      # (n*n, nS, 3)
    dist = disp.norm(dim=2)
    ok = dist < cutoff
    if not self_interaction:
        zero_shift = (S.abs().sum(1) == 0).unsqueeze(0)
        same = (ii == jj).unsqueeze(1)
        ok &= ~(zero_shift & same)
    pair, sh = torch.nonzero(ok, as_tuple=True)
    mapping = torch.stack([ii[pair], jj[pair]])
    return mapping, torch.zeros(mapping.shape[1], dtype=torch.long), S[sh]


def import_reference():
    for name in ('xitorch', 'xitorch.integrate', 'xitorch.optimize', 'torch_nl', 'matplotlib', 'matplotlib.pyplot'):
        sys.modules[name] = types.ModuleType(name)

    def _no(*a, **k):
        raise NotImplementedError('xitorch is not installed')
    sys.modules['xitorch.integrate'].solve_ivp = _no
    sys.modules['xitorch.optimize'].minimize = _no
    sys.modules['torch_nl'].compute_neighborlist = brute_force_neighborlist
    sys.path.insert(0, '/root/reference/src')
    import professad.functionals as F
    import professad.functional_tools as T
    import professad.system as S
    import professad.crystal_tools as C
    return F, T, S, C


def main():
    sys.path.insert(0, ROOT)
    from oracle import ofdft_oracle as orc           # only for the deterministic input generators + HC table
    F, T, S, C = import_reference()
    torch.set_num_threads(8)

    table = orc.hc_kernel_table(0.7143, n_eta=2001)
    table_rev = orc.hc_kernel_table(2.0 / 3.0, n_eta=2001)
    np.savez_compressed(os.path.join(HERE, 'hc_table.npz'), hc=table.numpy(), revhc=table_rev.numpy())

    def hc_with_table(cls, args, tab):
        orig = cls.generate_kernel
        cls.generate_kernel = lambda self, *a, **k: None
        obj = cls(args)
        cls.generate_kernel = orig
        obj.kernel = tab.clone()
        obj.debug = False
        return obj

    cases = {
        'rough_even': orc.synth_rough((12, 10, 14), seed=0),
        'rough_odd': orc.synth_rough((9, 11, 7), seed=1),
        'rough_mixed': orc.synth_rough((8, 9, 6), seed=2),
        'smooth16': orc.synth_smooth(16, 1),
    }
    for cname, (box, den) in cases.items():
        gen = torch.Generator().manual_seed(7)
        v_ext = -0.5 + 0.2 * torch.rand(*den.shape, dtype=torch.double, generator=gen)
        wts = F.WangTeterStyleFunctional((0.8, 0.7, lambda x: 1 + x + 0.1 * x * x))
        funcs = {
            'IonElectron': lambda b, n: F.IonElectron(b, n, v_ext),
            'Hartree': F.Hartree, 'ThomasFermi': F.ThomasFermi, 'Weizsaecker': F.Weizsaecker,
            'WangTeter': F.WangTeter, 'Perrot': F.Perrot, 'SmargiassiMadden': F.SmargiassiMadden,
            'WangGovindCarter98': F.WangGovindCarter98,
            'WangGovindCarter99': F.WangGovindCarter99().forward,
            'WangGovindCarter99_g3k12': F.WangGovindCarter99((0.9, 0.7, 3.0, 1.2)).forward,
            'WangTeterStyle': wts.forward,
            'lda_exchange': F.lda_exchange, 'perdew_zunger_correlation': F.perdew_zunger_correlation,
            'PerdewZunger': F.PerdewZunger,
            'pbe_exchange': F.pbe_exchange, 'pbe_correlation': F.pbe_correlation,
            'PerdewBurkeErnzerhof': F.PerdewBurkeErnzerhof,
            'HuangCarter': hc_with_table(F.HuangCarter, (0.01177, 0.7143, 1.2), table).forward,
            'RevisedHuangCarter': hc_with_table(F.RevisedHuangCarter, (0.45, 0.10, 2.0 / 3.0, 1.15), table_rev).forward,
        }
        out = dict(box=box.numpy(), den=den.numpy(), v_ext=v_ext.numpy())
        for name, f in funcs.items():
            d = den.clone()
            E = f(box, d).detach().reshape(()).item()
            V = T.get_functional_derivative(box, d, f).detach().numpy()
            out['E_' + name], out['V_' + name] = np.float64(E), V
            print(f'{cname:12s} {name:28s} E = {E:+.15e}  max|V| = {np.abs(V).max():.6e}')
        np.savez_compressed(os.path.join(HERE, f'functionals_{cname}.npz'), **out)

    # --- density optimisations through the reference's System ---------------------------------
    os.chdir(os.path.join(ROOT, 'tests'))
    runs = {}
    bv = 4.050 * torch.tensor([[0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5]], dtype=torch.double)
    runs['al_fcc18_wt_pbe'] = dict(box=bv, shape=(18, 18, 18), units='a',
                                   ions=[['Al', 'potentials/al.gga.recpot', torch.zeros(1, 3, dtype=torch.double)]],
                                   terms=[F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof],
                                   kwargs=dict(ntol=1e-7))
    runs['li_bcc18_sm_pbe'] = dict(box=3.48 * torch.eye(3, dtype=torch.double), shape=(18, 18, 18), units='a',
                                   ions=[['Li', 'potentials/li.gga.recpot',
                                          torch.tensor([[0, 0, 0], [0.5, 0.5, 0.5]], dtype=torch.double)]],
                                   terms=[F.IonIon, F.IonElectron, F.Hartree, F.SmargiassiMadden, F.PerdewBurkeErnzerhof],
                                   kwargs=dict(ntol=1e-7))
    bvc, fc = C.get_cell('fcc-c', vol_per_atom=16.8, coord_type='fractional')
    shp = S.System.ecut2shape(1600, bvc)
    runs['al_fcc4_config1'] = dict(box=bvc, shape=shp, units='a',
                                   ions=[['Al', 'potentials/al.gga.recpot', fc]],
                                   terms=[F.IonElectron, F.Hartree, F.ThomasFermi, F.Weizsaecker, F.PerdewZunger],
                                   kwargs=dict(ntol=1e-7, from_uniform=True))
    runs['al_fcc4_tpgd'] = dict(box=bvc, shape=(20, 20, 20), units='a',
                                ions=[['Al', 'potentials/al.gga.recpot', fc]],
                                terms=[F.IonElectron, F.Hartree, F.WangTeter, F.PerdewZunger],
                                kwargs=dict(ntol=1e-6, n_method='TPGD', n_conv_cond_count=5))
    runs['al_fcc4_wgc99'] = dict(box=bvc, shape=(24, 24, 24), units='a',
                                 ions=[['Al', 'potentials/al.gga.recpot', fc]],
                                 terms=[F.IonElectron, F.Hartree, F.WangGovindCarter99().forward, F.PerdewZunger],
                                 kwargs=dict(ntol=1e-7))
    for rname, r in runs.items():
        sysm = S.System(r['box'], r['shape'], r['ions'], r['terms'], units=r['units'], coord_type='fractional')
        n_closure = [0]
        sysm.optimize_density(**r['kwargs'])
        out = dict(box_bohr=sysm.lattice_vectors('b').numpy(), shape=np.array(r['shape']),
                   v_ext=sysm.ionic_potential().numpy(), den=sysm.density().numpy(),
                   energy_Ha=np.float64(sysm.energy('Ha')), energy_eV=np.float64(sysm.energy('eV')),
                   n_elec=np.float64(sysm.electron_count()),
                   frac=sysm.fractional_ionic_coordinates().numpy())
        if any(t.__qualname__ == 'IonIon' for t in r['terms']):
            out['E_ion_Ha'] = np.float64(sysm._System__Eion_cache)
        print(f'{rname:20s} E = {out["energy_eV"]:.10f} eV  shape {tuple(r["shape"])}')
        np.savez_compressed(os.path.join(HERE, f'denopt_{rname}.npz'), **out)


if __name__ == '__main__':
    main()
