"""Native ionic potential and ion-electron forces (csrc/ions.cu) against the unmodified reference's vectors
(tests/golden/ions_*.npz), the CPU oracle and the product's own torch lattice sum."""
import os
import threading

import numpy as np
import pytest
import torch

from test_oracle_ions import CASES, load_case

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('case', CASES)
def test_native_vext_and_forces_match_reference(case, golden_dir, potentials_dir):
    from profess_ad_b200 import ion_utils as IU
    g, box, den, species = load_case(case, golden_dir, potentials_dir)
    b = box.to(DEV)
    sp = [(p, f.to(DEV)) for p, f in species]
    v = IU.ionic_potential(b, den.shape, sp).cpu().numpy()
    assert np.abs(v - g['v_ext']).max() <= 1e-11 * np.abs(g['v_ext']).max()
    F = IU.ion_electron_forces(b, den.to(DEV), sp).cpu().numpy()
    assert np.abs(F - g['forces_IonElectron']).max() <= 1e-10 * np.abs(g['forces_IonElectron']).max()


@pytest.mark.parametrize('case', CASES)
def test_system_forces_match_reference(case, golden_dir, potentials_dir):
    import profess_ad_b200.functionals as F
    from profess_ad_b200.system import System
    g, box, den, species = load_case(case, golden_dir, potentials_dir)
    ions = [[os.path.basename(p)[:2].capitalize(), p, f] for p, f in species]
    s = System(box, tuple(den.shape), ions, [F.IonIon, F.IonElectron, F.Hartree, F.ThomasFermi], units='b',
               coord_type='fractional')
    assert np.abs(s.ionic_potential().cpu().numpy() - g['v_ext']).max() <= 1e-11 * np.abs(g['v_ext']).max()
    s.set_density(torch.from_numpy(g['den']))
    forces = s.forces('Ha/b').cpu().numpy()
    assert np.abs(forces - g['forces_Ha_b']).max() <= 1e-9 * np.abs(g['forces_Ha_b']).max()
    ev_a = s.forces('eV/a').cpu().numpy()
    assert np.allclose(ev_a, forces * System.eV_per_Ha / System.A_per_b, rtol=1e-14)
    with pytest.raises(ValueError):
        s.forces('N')


def test_many_ions_vs_torch_lattice_sum_and_oracle(potentials_dir):
    """More ions than one shared-memory chunk (128), all-even grid on a skewed cell: the native kernel against
    the O(N_k N_ion) torch path of the product and the autograd forces of the CPU oracle."""
    from oracle import ofdft_oracle as orc
    from profess_ad_b200 import ion_utils as IU
    from profess_ad_b200.functional_tools import wavevecs
    shape = (20, 16, 18)
    box, den = orc.synth_rough(shape, seed=9, L=14.0)
    gen = torch.Generator().manual_seed(4)
    f_al = torch.rand(150, 3, dtype=torch.double, generator=gen)
    f_li = torch.rand(3, 3, dtype=torch.double, generator=gen)
    pa, pl = os.path.join(potentials_dir, 'al.gga.recpot'), os.path.join(potentials_dir, 'li.gga.recpot')
    b = box.to(DEV)
    v = IU.ionic_potential(b, shape, [(pa, f_al.to(DEV)), (pl, f_li.to(DEV))])
    k = torch.sqrt(wavevecs(b, shape)[3])
    v_t = (IU.lattice_sum(b, shape, (f_al @ box).to(DEV), IU.interpolate_recpot(pa, k))
           + IU.lattice_sum(b, shape, (f_li @ box).to(DEV), IU.interpolate_recpot(pl, k)))
    assert ((v - v_t).abs().max() / v_t.abs().max()).item() < 1e-11
    v_o = orc.ionic_potential(box, shape, [(pa, f_al @ box), (pl, f_li @ box)])
    assert ((v.cpu() - v_o).abs().max() / v_o.abs().max()).item() < 1e-11
    F_o = orc.ion_electron_forces(box, den, [(pa, f_al @ box), (pl, f_li @ box)])
    F = IU.ion_electron_forces(b, den.to(DEV), [(pa, f_al.to(DEV)), (pl, f_li.to(DEV))]).cpu()
    assert ((F - F_o).abs().max() / F_o.abs().max()).item() < 1e-10


def _slab_rank(comm, shape, box, den, species, out, idx, errors):
    from profess_ad_b200 import parallel, ion_utils as IU
    try:
        dev = torch.device(DEV)
        with torch.cuda.stream(torch.cuda.Stream(dev)):
            with parallel.slab(shape, comm=comm) as ctx:
                sp = [(p, f.to(dev)) for p, f in species]
                v = IU.ionic_potential(box.to(dev), ctx.local_shape, sp)
                d = parallel.local_slab(den.to(dev))
                F = IU.ion_electron_forces(box.to(dev), d, sp)
                torch.cuda.current_stream(dev).synchronize()
                out[idx] = (v.cpu(), F.cpu())
    except BaseException as e:      # noqa: BLE001
        errors.append(e)
        try:
            comm.shared.barrier.abort()
        except Exception:
            pass


def test_slab_vext_and_forces(golden_dir, potentials_dir):
    from profess_ad_b200 import parallel
    world = 2
    g, box, den, species = load_case('li2_even', golden_dir, potentials_dir)      # (12, 14, 12): n0, n1 multiples of 2
    shape = tuple(den.shape)
    out, errors = [None] * world, []
    shared = parallel.ThreadComm.Shared(world)
    threads = [threading.Thread(target=_slab_rank, args=(parallel.ThreadComm(shared, r), shape, box, den, species, out, r, errors))
               for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    if errors:
        raise errors[0]
    v = torch.cat([o[0] for o in out], dim=0).numpy()
    assert np.abs(v - g['v_ext']).max() <= 1e-11 * np.abs(g['v_ext']).max()
    for o in out:
        assert np.abs(o[1].numpy() - g['forces_IonElectron']).max() <= 1e-10 * np.abs(g['forces_IonElectron']).max()


# ------------------------------------------------------------------------------------------------
#  stress: native analytic kernels vs the reference's autograd vectors
# ------------------------------------------------------------------------------------------------
def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize('case', CASES)
def test_functional_stresses_match_reference(case, golden_dir, potentials_dir):
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import ion_utils as IU
    from profess_ad_b200.functional_tools import get_stress, get_pressure
    g, box, den, species = load_case(case, golden_dir, potentials_dir)
    b, d = box.to(DEV), den.to(DEV)
    n = 0
    for key in g.files:
        if key.startswith('stress_') and key[7:] not in ('Ha_b3', 'IonElectron', 'IonIon'):
            f = getattr(F, key[7:])
            st = get_stress(b, d, f).cpu().numpy()
            assert _rel(st, g[key]) < 1e-10, (key, _rel(st, g[key]))
            assert abs(get_pressure(b, d, f).item() + np.trace(g[key]) / 3) <= 1e-10 * np.abs(g[key]).max()
            n += 1
    assert n >= 3
    st = IU.ion_electron_stress(b, d, [(p, f.to(DEV)) for p, f in species]).cpu().numpy()
    assert _rel(st, g['stress_IonElectron']) < 1e-10


@pytest.mark.parametrize('case', CASES)
def test_system_stress_and_pressure_match_reference(case, golden_dir, potentials_dir):
    import profess_ad_b200.functionals as F
    from profess_ad_b200.system import System
    g, box, den, species = load_case(case, golden_dir, potentials_dir)
    ions = [[os.path.basename(p)[:2].capitalize(), p, f] for p, f in species]
    if case == 'alli_mixed':
        terms = [F.IonIon, F.IonElectron, F.Hartree, F.ThomasFermi, F.Weizsaecker, F.PerdewZunger]
    else:
        terms = [F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof]
    s = System(box, tuple(den.shape), ions, terms, units='b', coord_type='fractional')
    s.set_density(torch.from_numpy(g['den']))
    assert abs(s.energy('Ha') - float(g['energy_Ha'])) < 1e-10
    st = s.stress('Ha/b3').cpu().numpy()
    assert np.abs(st - g['stress_Ha_b3']).max() <= 1e-9 * np.abs(g['stress_Ha_b3']).max(), np.abs(st - g['stress_Ha_b3']).max()
    assert abs(s.pressure('Ha/b3') - float(g['pressure_Ha_b3'])) <= 1e-9 * np.abs(g['stress_Ha_b3']).max()
    assert np.allclose(s.stress('GPa').cpu().numpy(), st * System.GPa_per_atomic, rtol=1e-14)
    assert abs(s.enthalpy('Ha') - (s.energy('Ha') + s.pressure() * s.volume('b3'))) < 1e-14
    with pytest.raises(ValueError):
        s.stress('bar')


@pytest.mark.parametrize('case', CASES)
def test_wgc99_stress_matches_reference_fresh_kernel(case, golden_dir, potentials_dir):
    """WGC99 stress = the reference's autograd with a freshly generated kernel (all three branches of the kernel's
    homogeneous solution: default gamma -> complex exponents, gamma = 0.5 -> real exponents; kappa != 1)."""
    import profess_ad_b200.functionals as F
    from profess_ad_b200.functional_tools import get_stress
    g, box, den, _ = load_case(case, golden_dir, potentials_dir)
    b, d = box.to(DEV), den.to(DEV)
    A98, B98 = (5 + 5 ** 0.5) / 6, (5 - 5 ** 0.5) / 6
    for key, f in (('stressfresh_WGC99', F.WangGovindCarter99()), ('stressfresh_WGC99_gamma05', F.WangGovindCarter99((A98, B98, 0.5, 1.0))),
                   ('stressfresh_WGC99_kappa12', F.WangGovindCarter99((A98, B98, 2.7, 1.2)))):
        st = get_stress(b, d, f.forward).cpu().numpy()
        assert _rel(st, g[key]) < 1e-9, (key, _rel(st, g[key]))


def test_stress_vs_oracle_rough_grids_and_unsupported():
    from oracle import ofdft_oracle as orc
    import profess_ad_b200.functionals as F
    from profess_ad_b200.functional_tools import get_stress
    pairs = [(F.ThomasFermi, orc.ThomasFermi), (F.Weizsaecker, orc.Weizsaecker), (F.Hartree, orc.Hartree),
             (F.WangTeter, orc.WangTeter), (F.WangGovindCarter98, orc.WangGovindCarter98), (F.SmargiassiMadden, orc.SmargiassiMadden),
             (F.PerdewZunger, orc.PerdewZunger), (F.PerdewBurkeErnzerhof, orc.PerdewBurkeErnzerhof),
             (F.pbe_exchange, orc.pbe_exchange), (F.lda_exchange, orc.lda_exchange)]
    for shape, seed in (((16, 18, 20), 3), ((15, 17, 13), 4), ((12, 15, 16), 5)):
        box, den = orc.synth_rough(shape, seed=seed)
        for f, fo in pairs:
            ref = orc.stress(box, den, fo).numpy()
            st = get_stress(box.to(DEV), den.to(DEV), f).cpu().numpy()
            assert _rel(st, ref) < 1e-10, (shape, f.__name__, _rel(st, ref))
    with pytest.raises(NotImplementedError):
        get_stress(box.to(DEV), den.to(DEV), lambda b, n: F.ThomasFermi(b, n))


def test_slab_stress(golden_dir, potentials_dir):
    """Stress on slab plans: the 7 sums of every piece are all-reduced (world 2, ranks as threads)."""
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import parallel, ion_utils as IU
    from profess_ad_b200.functional_tools import get_stress
    g, box, den, species = load_case('li2_even', golden_dir, potentials_dir)
    shape, world = tuple(den.shape), 2
    out, errors = [None] * world, []

    def rank(comm, idx):
        try:
            dev = torch.device(DEV)
            with torch.cuda.stream(torch.cuda.Stream(dev)):
                with parallel.slab(shape, comm=comm):
                    d = parallel.local_slab(den.to(dev))
                    b = box.to(dev)
                    st = get_stress(b, d, F.WangTeter) + get_stress(b, d, F.PerdewBurkeErnzerhof) + get_stress(b, d, F.Hartree)
                    st = st + IU.ion_electron_stress(b, d, [(p, f.to(dev)) for p, f in species])
                    torch.cuda.current_stream(dev).synchronize()
                    out[idx] = st.cpu().numpy()
        except BaseException as e:      # noqa: BLE001
            errors.append(e)
            try:
                comm.shared.barrier.abort()
            except Exception:
                pass
    shared = parallel.ThreadComm.Shared(world)
    threads = [threading.Thread(target=rank, args=(parallel.ThreadComm(shared, r), r)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    if errors:
        raise errors[0]
    ref = g['stress_WangTeter'] + g['stress_PerdewBurkeErnzerhof'] + g['stress_Hartree'] + g['stress_IonElectron']
    for st in out:
        assert _rel(st, ref) < 1e-10


@pytest.mark.parametrize('case', CASES)
def test_huang_carter_stress_matches_reference(case, golden_dir, potentials_dir):
    """Stress of HuangCarter / RevisedHuangCarter (TF + vW + non-local term, csrc/hc.cu:pad_stress_hc_nl) against the
    unmodified reference's autograd through box_vecs (tests/golden/hc_stress.npz, make_golden_hc_stress.py; the omega(eta)
    table injected on both sides), and the same through System.stress()."""
    import profess_ad_b200.functionals as F
    from profess_ad_b200.functional_tools import get_stress
    g, box, den, species = load_case(case, golden_dir, potentials_dir)
    hs = np.load(os.path.join(golden_dir, 'hc_stress.npz'))
    tab = np.load(os.path.join(golden_dir, 'hc_table.npz'))
    b, d = box.to(DEV), den.to(DEV)
    for name, f in (('HC', F.HuangCarter(tuple(hs['hc_args']), kernel=torch.from_numpy(tab['hc']))),
                    ('revHC', F.RevisedHuangCarter(tuple(hs['revhc_args']), kernel=torch.from_numpy(tab['revhc'])))):
        E = f.forward(b, d).item()
        assert abs(E - float(hs[f'{case}_{name}_E'])) <= 1e-10 * abs(E)
        st = get_stress(b, d, f.forward).cpu().numpy()
        ref = hs[f'{case}_{name}']
        assert _rel(st, ref) < 1e-9, (name, st, ref)
        assert np.abs(st - st.T).max() <= 1e-12 * np.abs(st).max()
