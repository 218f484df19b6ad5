"""GPU parity of the System facade and the density-optimisation loop against the unmodified
reference (golden runs in tests/golden/denopt_*.npz) and the reference's own known answers.

Gate (BASELINE.json): optimised energies within 1e-6 eV/atom."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

EV = 27.211386245988


def _system(case, golden_dir, potentials_dir, terms, pot, **kw):
    from profess_ad_b200.system import System
    g = np.load(os.path.join(golden_dir, f'denopt_{case}.npz'))
    box = torch.from_numpy(g['box_bohr'])
    shape = tuple(int(s) for s in g['shape'])
    ions = [[pot[:2].capitalize(), os.path.join(potentials_dir, pot), torch.from_numpy(g['frac'])]]
    return System(box, shape, ions, terms, units='b', coord_type='fractional', **kw), g


def test_vext_and_ion_energy_match_reference(golden_dir, potentials_dir):
    import profess_ad_b200.functionals as F
    terms = [F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof]
    s, g = _system('al_fcc18_wt_pbe', golden_dir, potentials_dir, terms, 'al.gga.recpot')
    v = s.ionic_potential().cpu().numpy()
    assert np.abs(v - g['v_ext']).max() <= 1e-11 * np.abs(g['v_ext']).max()
    assert abs(s._System__Eion_cache - float(g['E_ion_Ha'])) < 1e-10
    assert s.electron_count() == 3 and s.ion_count() == 1


CASES = {
    'al_fcc18_wt_pbe': ('al.gga.recpot', lambda F: [F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof],
                        dict(ntol=1e-7), 1),
    'li_bcc18_sm_pbe': ('li.gga.recpot', lambda F: [F.IonIon, F.IonElectron, F.Hartree, F.SmargiassiMadden, F.PerdewBurkeErnzerhof],
                        dict(ntol=1e-7), 2),
    'al_fcc4_config1': ('al.gga.recpot', lambda F: [F.IonElectron, F.Hartree, F.ThomasFermi, F.Weizsaecker, F.PerdewZunger],
                        dict(ntol=1e-7, from_uniform=True), 4),
    'al_fcc4_tpgd': ('al.gga.recpot', lambda F: [F.IonElectron, F.Hartree, F.WangTeter, F.PerdewZunger],
                     dict(ntol=1e-6, n_method='TPGD', n_conv_cond_count=5), 4),
    'al_fcc4_wgc99': ('al.gga.recpot', lambda F: [F.IonElectron, F.Hartree, F.WangGovindCarter99().forward, F.PerdewZunger],
                      dict(ntol=1e-7), 4),
}


@pytest.mark.parametrize('native', [True, False], ids=['device_resident', 'host_driven'])
@pytest.mark.parametrize('case', list(CASES))
def test_optimize_density_matches_reference(case, native, golden_dir, potentials_dir, monkeypatch):
    import profess_ad_b200.functionals as F
    from profess_ad_b200.system import System
    monkeypatch.setattr(System, 'use_native_optimizer', native)
    pot, terms, kw, natoms = CASES[case]
    s, g = _system(case, golden_dir, potentials_dir, terms(F), pot)
    s.optimize_density(**kw)
    if native:
        assert s.last_optimization.get('native') and s.last_optimization['converged']
    dE_eV_per_atom = abs(s.energy('eV') - float(g['energy_eV'])) / natoms
    assert dE_eV_per_atom < 1e-6, f'{case}: {dE_eV_per_atom:.3e} eV/atom'
    den = s.density().cpu().numpy()
    assert np.abs(den - g['den']).max() < 1e-5


def test_known_answers_profess4(golden_dir, potentials_dir):
    """tests/test_match_profess4.py:23,35 (atol 1e-4 eV)."""
    import profess_ad_b200.functionals as F
    s, _ = _system('al_fcc18_wt_pbe', golden_dir, potentials_dir, CASES['al_fcc18_wt_pbe'][1](F), 'al.gga.recpot')
    s.optimize_density(ntol=1e-7)
    assert abs(s.energy('eV') - (-57.183329401794985)) < 1e-4
    s, _ = _system('li_bcc18_sm_pbe', golden_dir, potentials_dir, CASES['li_bcc18_sm_pbe'][1](F), 'li.gga.recpot')
    s.optimize_density(ntol=1e-7)
    assert abs(s.energy('eV') - (-14.741886997024537)) < 1e-4


def test_potentials_hook_and_convergence_measures(golden_dir, potentials_dir):
    """tests/test_functional_derivative.py:120-139 and tests/test_den_opt.py:58-75."""
    import profess_ad_b200.functionals as F
    from profess_ad_b200.functional_tools import get_functional_derivative
    terms = [F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof]
    s, _ = _system('al_fcc4_tpgd', golden_dir, potentials_dir, terms, 'al.gga.recpot')
    s.optimize_density(ntol=1e-7)
    E1, den1 = s.energy(), s.density().clone()

    def dEdn(bv, n):
        return s.ionic_potential() + sum(get_functional_derivative(bv, n, f) for f in terms[1:])
    s.initialize_density()
    s.optimize_density(ntol=1e-7, potentials=dEdn)
    assert abs(E1 - s.energy()) <= 1e-7 * abs(E1)
    assert (den1 - s.density()).abs().max().item() < 1e-5

    dEdchi = s.check_density_convergence()
    v = s.functional_derivative('density')
    chi = torch.sqrt(s.density())
    n_tilde = torch.mean(chi.pow(2)) * s.volume()
    mu = torch.mean(v * s.density()) * s.volume() / s.electron_count()
    proj = (s.electron_count() / n_tilde) * 2 * chi * (v - mu)
    assert abs(dEdchi - proj.abs().max().item()) <= 1e-10 * dEdchi
    assert abs(s.chemical_potential() - mu.item()) < 1e-12
    assert s.check_density_convergence('euler') < 1e-3


def test_setters_and_errors(golden_dir, potentials_dir):
    import profess_ad_b200.functionals as F
    s, g = _system('al_fcc4_tpgd', golden_dir, potentials_dir, [F.IonElectron, F.ThomasFermi], 'al.gga.recpot')
    e0 = s.energy()
    s.set_lattice(1.02 * s.lattice_vectors('b'), units='b')
    assert abs(s.density().mean().item() * s.volume() - s.electron_count()) < 1e-10      # N conserved
    assert s.energy() != e0
    with pytest.raises(ValueError):
        s.set_lattice(s.lattice_vectors('b'), units='x')
    with pytest.raises(ValueError):
        s.optimize_density(n_method='nope')
    with pytest.raises(NotImplementedError):
        s.elastic_constants()
    assert s.stress().shape == (3, 3)
    with pytest.raises(AssertionError):
        s.set_density(torch.ones(3, 3, 3, dtype=torch.double))


def test_fused_evaluator_and_projection_match_oracle():
    """pad_eval_total (whole term list in one call) and pad_chi_project vs the autograd oracle."""
    from oracle import ofdft_oracle as orc
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import _density_opt as D
    dev = torch.device('cuda:0')
    for shape, seed in (((16, 18, 20), 3), ((15, 17, 13), 4)):
        box, den = orc.synth_rough(shape, seed=seed)
        gen = torch.Generator().manual_seed(seed)
        v_ext = -0.5 + 0.2 * torch.rand(*shape, dtype=torch.double, generator=gen)
        chi = torch.sqrt(den) * (1 + 0.1 * torch.rand(*shape, dtype=torch.double, generator=gen))
        n_elec = 11.0
        combos = [
            ([F.IonElectron, F.Hartree, F.WangGovindCarter99().forward, F.PerdewZunger],
             [orc.IonElectron, orc.Hartree, orc.WangGovindCarter99(), orc.PerdewZunger]),
            ([F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof],
             [orc.IonElectron, orc.Hartree, orc.WangTeter, orc.PerdewBurkeErnzerhof]),
            ([F.IonElectron, F.Hartree, F.ThomasFermi, F.Weizsaecker, F.PerdewZunger],
             [orc.IonElectron, orc.Hartree, orc.ThomasFermi, orc.Weizsaecker, orc.PerdewZunger]),
        ]
        for native_terms, oracle_terms in combos:
            T = D.describe_terms(native_terms)
            assert T is not None
            E_ref, g_ref, n_ref = orc.chi_gradient(box, chi, n_elec, oracle_terms, v_ext)
            b, c = box.to(dev), chi.to(dev)
            n = D.chi_to_density(b, c, n_elec)
            assert ((n.cpu() - n_ref).abs().max() / n_ref.abs().max()).item() < 1e-13
            E, v = D.eval_total(b, n, v_ext.to(dev), T)
            assert abs(E.item() - E_ref.item()) <= 1e-10 * abs(E_ref.item())
            g, stats = D.chi_project(b, c, n, v, n_elec)
            assert ((g.cpu() - g_ref).abs().max() / g_ref.abs().max()).item() < 1e-9
            stats = stats.cpu()
            dV = abs(torch.linalg.det(box).item()) / chi.numel()
            assert abs(stats[0].item() - g_ref.abs().sum().item()) <= 1e-9 * g_ref.abs().sum().item()
            assert abs(stats[1].item() - (g_ref * g_ref).sum().item()) <= 1e-9 * (g_ref * g_ref).sum().item()
            assert abs(stats[2].item() - (g_ref / dV).abs().max().item()) <= 1e-9 * (g_ref / dV).abs().max().item()
    assert D.describe_terms([F.Hartree, lambda b, n: F.ThomasFermi(b, n)]) is None      # user term -> generic path
    assert D.describe_terms([F.ThomasFermi, F.ThomasFermi]) is None


def test_huang_carter_density_optimisation_native_vs_oracle(golden_dir, potentials_dir):
    """Huang-Carter family in the fused evaluator and the device-resident optimiser (pad_terms.kinetic == 3):
    same optimised energy as the CPU oracle's restatement of the reference loop, with the same omega(eta) table
    injected on both sides."""
    import profess_ad_b200.functionals as F
    from profess_ad_b200.system import System
    from profess_ad_b200 import _density_opt as D
    from oracle import ofdft_oracle as orc
    g = np.load(os.path.join(golden_dir, 'denopt_al_fcc4_config1.npz'))
    with np.load(os.path.join(golden_dir, 'hc_table.npz')) as tab:
        t_rev = torch.from_numpy(tab['revhc'])
    box = torch.from_numpy(g['box_bohr'])
    # 16^3: on a 12^3 grid this cell drives the fixed-step L-BFGS through a negative-curvature region where rounding
    # differences grow 15x per iteration and every implementation (the reference on another BLAS included) stops
    # somewhere else; from 16^3 on the optimisation is well conditioned
    shape = (16, 16, 16)
    frac = torch.from_numpy(g['frac'])
    pot = os.path.join(potentials_dir, 'al.gga.recpot')
    hc = F.RevisedHuangCarter((0.45, 0.10, 2.0 / 3.0, 1.15), kernel=t_rev.clone())
    terms = [F.IonElectron, F.Hartree, hc.forward, F.PerdewZunger]
    assert D.describe_terms(terms) is not None and D.describe_terms(terms).kinetic == 3
    s = System(box, shape, [['Al', pot, frac]], terms, units='b', coord_type='fractional')
    s.optimize_density(ntol=1e-7, from_uniform=True)
    assert s.last_optimization.get('native') and s.last_optimization['converged']
    v_ext = s.ionic_potential().cpu()
    n_elec = s.electron_count()
    den0 = torch.full(shape, n_elec / s.volume(), dtype=torch.double)
    ohc = orc.RevisedHuangCarter(0.45, 0.10, 2 / 3, 1.15, kernel=t_rev)
    ref = orc.optimize_density(box, den0, n_elec, [orc.IonElectron, orc.Hartree, ohc, orc.PerdewZunger], v_ext=v_ext, ntol=1e-7)
    assert abs(s.energy('Ha') - ref['energy']) * EV / frac.shape[0] < 1e-6, (s.energy('Ha'), ref['energy'])


def test_eos_scan_reproduces_published_table(potentials_dir):
    """docs/source/example_elastic.rst:81-86 (fcc Al, WT + PBE, 2000 eV): V0 = 16.76389 A^3, E0 = -57.18370 eV,
    K0 = 78.80961 GPa -- through parallel.eos_fit (independent systems per GPU; world 1 here)."""
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import parallel
    from profess_ad_b200.crystal_tools import get_cell
    from profess_ad_b200.system import System
    pot = os.path.join(potentials_dir, 'al.gga.recpot')

    def make_system():
        box, frac = get_cell('fcc', vol_per_atom=16.9, coord_type='fractional')
        terms = [F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof]
        s = System(box, System.ecut2shape(2000, box), [['Al', pot, frac]], terms, units='a', coord_type='fractional')
        s.optimize_density(ntol=1e-10)
        return s
    params, err = parallel.eos_fit(make_system, f=0.05, N=9, eos='bm')
    assert abs(params[3] - 16.76389) < 2e-4
    assert abs(params[2] - (-57.18370)) < 2e-5
    assert abs(params[0] - 78.80961) < 0.02


def test_exact_single_orbital_limits(potentials_dir):
    """The reference's own known answers (tests/test_den_opt.py:13-41): with Weizsaecker as the only kinetic term a
    one-electron system is exact -- hydrogen atom in a 20 bohr box: -0.5 Ha (2 places); harmonic oscillator with
    k = 10: 3/2 sqrt(k) Ha (5 places; docs/source/example_density_optimization.rst:175-221 quotes 4.74341650)."""
    import profess_ad_b200.functionals as F
    from profess_ad_b200.system import System
    L = 20.0
    box = L * torch.eye(3, dtype=torch.double)
    shape = System.ecut2shape(250, box)
    ions = [['H', os.path.join(potentials_dir, 'H.coulomb-kcut-15.recpot'), torch.tensor([[0.5, 0.5, 0.5]]).double()]]
    s = System(box, shape, ions, [F.IonElectron, F.Weizsaecker], units='b', coord_type='fractional')
    s.set_electron_number(1)
    s.optimize_density(ntol=1e-4)
    assert abs(s.energy('Ha') - (-0.5)) < 5e-3
    k = 10
    f = [torch.arange(n, dtype=torch.double) / n for n in shape]
    x, y, z = torch.meshgrid(*[L * fi for fi in f], indexing='ij')
    qho = 0.5 * k * ((x - L / 2).pow(2) + (y - L / 2).pow(2) + (z - L / 2).pow(2))
    s.set_potential(qho)
    s.initialize_density()
    s.optimize_density(ntol=1e-4)
    assert abs(s.energy('Ha') - 1.5 * np.sqrt(k)) < 5e-6
