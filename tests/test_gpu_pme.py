"""Particle-mesh Ewald structure factor on the device (csrc/ions.cu: k_pme_spread / k_pme_spectrum) against the exact
structure factor (the reference's tests/test_particle_mesh_ewald.py:46-63), against the torch restatement, and through
System(pme_order=...) (test4 of that file: same energy, density, forces as with exact structure factors)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_structure_factor_low_k_matches_exact():
    from profess_ad_b200 import ion_utils as IU
    shape = (35, 36, 37)
    box = torch.tensor([[4.9, 0.1, 0.2], [-0.2, 5.0, 0.3], [0.3, -0.1, 5.1]], dtype=torch.double, device=DEV)
    cart = torch.tensor([[0, 0, 0], [2, 0.1, 0.2], [0.3, 1, 2]], dtype=torch.double, device=DEV)
    S = IU.structure_factor(box, shape, cart).cpu().numpy()
    Sp = IU.structure_factor_spline(box, shape, cart, 20).cpu().numpy()
    t = 10
    for sl in ((slice(None, t), slice(None, t)), (slice(None, t), slice(-t, None)), (slice(-t, None), slice(None, t)),
               (slice(-t, None), slice(-t, None))):
        assert np.allclose(S[sl[0], sl[1], :t], Sp[sl[0], sl[1], :t])


@pytest.mark.parametrize('shape,order', [((20, 21, 22), 4), ((20, 21, 22), 12), ((16, 12, 18), 8), ((35, 36, 37), 20), ((24, 24, 24), 2)])
def test_native_spreading_equals_torch_restatement(shape, order):
    from profess_ad_b200 import ion_utils as IU
    gen = torch.Generator().manual_seed(order)
    box = (7.5 * torch.eye(3, dtype=torch.double) + 0.1 * torch.rand(3, 3, dtype=torch.double, generator=gen)).to(DEV)
    frac = torch.rand(7, 3, dtype=torch.double, generator=gen) * 3 - 1           # also outside [0, 1): wrapped
    cart = frac.to(DEV) @ box
    a = IU.structure_factor_spline(box, shape, cart, order)
    b = IU._structure_factor_spline_torch(box, shape, cart, order)
    # the exponential-spline factors b(m) grow like (pi / 2)^order towards the Nyquist frequencies and amplify the rounding of the
    # spread charges (fused multiply-adds in the kernel's Cox-de Boor recursion, another FFT) by the same factor
    # (three axes: up to (pi / 2)^(3 order) at the Nyquist corner -- 6e11 for order 20; the low-frequency check below is the sharp one)
    tol = 1e-12 if order <= 8 else 2e-15 * (np.pi / 2) ** (3 * order)
    err = ((a - b).abs().max() / b.abs().max()).item()
    assert err < tol, err
    t = 6                      # the low-frequency corner, where b = O(1): rounding level
    err_low = ((a[:t, :t, :t] - b[:t, :t, :t]).abs().max() / b[:t, :t, :t].abs().max()).item()
    assert err_low < 1e-12, err_low


def test_pme_ionic_potential_and_system(potentials_dir):
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import ion_utils as IU
    from profess_ad_b200.functional_tools import wavevecs
    from profess_ad_b200.system import System
    pot = os.path.join(potentials_dir, 'li.gga.recpot')
    shape, L = (25, 25, 25), 6.96
    box = L * torch.eye(3, dtype=torch.double)
    frac = torch.tensor([[0, 0, 0], [0.5, 0.5, 0.5]], dtype=torch.double)
    # native PME v_ext == torch lattice_sum with the spline structure factor (reference semantics)
    b = box.to(DEV)
    v = IU.ionic_potential(b, shape, [(pot, frac.to(DEV))], pme_order=20)
    k = torch.sqrt(wavevecs(b, shape)[3])
    v_t = IU.lattice_sum(b, shape, (frac @ box).to(DEV), IU.interpolate_recpot(pot, k), 20)
    assert ((v - v_t).abs().max() / v_t.abs().max()).item() < 1e-11
    # a skewed, even grid: the Hermitian part on the self-conjugate planes matters
    gen = torch.Generator().manual_seed(2)
    box2 = (L * torch.eye(3, dtype=torch.double) + 0.3 * torch.rand(3, 3, dtype=torch.double, generator=gen)).to(DEV)
    shape2 = (24, 20, 22)
    v = IU.ionic_potential(box2, shape2, [(pot, frac.to(DEV))], pme_order=8)
    k = torch.sqrt(wavevecs(box2, shape2)[3])
    v_t = IU.lattice_sum(box2, shape2, frac.to(DEV) @ box2, IU.interpolate_recpot(pot, k), 8)
    assert ((v - v_t).abs().max() / v_t.abs().max()).item() < 1e-11
    # tests/test_particle_mesh_ewald.py:65-89
    terms = [F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof]
    res = []
    for order in (None, 20):
        s = System(box, shape, [['Li', pot, frac]], terms, units='b', coord_type='fractional', pme_order=order)
        s.optimize_density()
        res.append((s.energy('eV'), s.density().cpu().numpy(), s.forces().cpu().numpy(), s.stress().cpu().numpy()))
    assert np.allclose(res[0][0], res[1][0])
    assert np.allclose(res[0][1], res[1][1])
    assert np.allclose(res[0][2], res[1][2], atol=1e-8)
    assert np.allclose(res[0][3], res[1][3])


@pytest.mark.parametrize('case', ['li2_odd', 'li2_even', 'alli_mixed'])
@pytest.mark.parametrize('order', [4, 8])
def test_pme_forces_and_stress_match_reference_autograd(case, order, golden_dir, potentials_dir):
    """IonElectron with the particle-mesh structure factor: v_ext, forces (one c2r + gather of B-spline derivative weights,
    pad_ion_forces_pme) and stress (pad_ion_stress_pme) against the UNMODIFIED reference's values / autograd through
    structure_factor_spline (System(pme_order=n), tests/golden/ions_pme.npz from make_golden_pme_forces.py) -- and through
    System.forces() / System.stress()."""
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import ion_utils as IU
    from profess_ad_b200.system import System
    from test_oracle_ions import load_case
    g, box, den, species = load_case(case, golden_dir, potentials_dir)
    ref = np.load(os.path.join(golden_dir, 'ions_pme.npz'))
    key = f'{case}_o{order}_'
    b, d = box.to(DEV), den.to(DEV)
    sp = [(p, f.to(DEV)) for p, f in species]
    v = IU.ionic_potential(b, tuple(den.shape), sp, pme_order=order).cpu().numpy()
    assert np.abs(v - ref[key + 'vext']).max() <= 1e-10 * np.abs(ref[key + 'vext']).max()
    Fp = IU.ion_electron_forces(b, d, sp, pme_order=order).cpu().numpy()
    assert np.abs(Fp - ref[key + 'forces']).max() <= 1e-9 * np.abs(ref[key + 'forces']).max(), (Fp, ref[key + 'forces'])
    # (and they are NOT the forces of the exact structure factor: the mesh error is differentiated, as in the reference)
    assert np.abs(Fp - g['forces_IonElectron']).max() > 1e-6 * np.abs(g['forces_IonElectron']).max()
    st = IU.ion_electron_stress(b, d, sp, pme_order=order).cpu().numpy()
    assert np.abs(st - ref[key + 'stress']).max() <= 1e-9 * np.abs(ref[key + 'stress']).max(), (st, ref[key + 'stress'])
    ions = [[os.path.basename(p)[:2].capitalize(), p, f] for p, f in species]
    s = System(box, tuple(den.shape), ions, [F.IonElectron], units='b', coord_type='fractional', pme_order=order)
    s.set_density(den)
    assert abs(s.energy('Ha') - float(ref[key + 'energy'])) <= 1e-10 * abs(float(ref[key + 'energy']))
    assert np.abs(s.forces('Ha/b').cpu().numpy() - ref[key + 'forces']).max() <= 1e-9 * np.abs(ref[key + 'forces']).max()
    assert np.abs(s.stress('Ha/b3').cpu().numpy() - ref[key + 'stress']).max() <= 1e-9 * np.abs(ref[key + 'stress']).max()


@pytest.mark.parametrize('case,world', [('li2_even', 2), ('alli_mixed', 1)])
def test_slab_pme_vext_forces_stress(case, world, golden_dir, potentials_dir):
    """Particle-mesh Ewald on slab plans: every rank spreads ALL ions onto its own x-planes, the mesh transform is the slab FFT,
    the partner values of the special points come from the Hermitian symmetry of the transforms (never from another rank), the
    force gather runs over the rank's planes and the partial forces are added -- against the reference's vectors."""
    import threading
    from profess_ad_b200 import parallel, ion_utils as IU
    from test_oracle_ions import load_case
    g, box, den, species = load_case(case, golden_dir, potentials_dir)
    ref = np.load(os.path.join(golden_dir, 'ions_pme.npz'))
    key = f'{case}_o8_'
    shape = tuple(den.shape)
    out, errors = [None] * world, []

    def rank(comm, idx):
        try:
            dev = torch.device(DEV)
            with torch.cuda.stream(torch.cuda.Stream(dev)):
                with parallel.slab(shape, comm=comm) as ctx:
                    d = parallel.local_slab(den.to(dev))
                    b = box.to(dev)
                    sp = [(p, f.to(dev)) for p, f in species]
                    v = IU.ionic_potential(b, ctx.local_shape, sp, pme_order=8)
                    F_ = IU.ion_electron_forces(b, d, sp, pme_order=8)
                    st = IU.ion_electron_stress(b, d, sp, pme_order=8)
                    torch.cuda.current_stream(dev).synchronize()
                    out[idx] = (v.cpu().numpy(), F_.cpu().numpy(), st.cpu().numpy())
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
    v = np.concatenate([o[0] for o in out], axis=0)
    assert np.abs(v - ref[key + 'vext']).max() <= 1e-10 * np.abs(ref[key + 'vext']).max()
    for _, F_, st in out:
        assert np.abs(F_ - ref[key + 'forces']).max() <= 1e-9 * np.abs(ref[key + 'forces']).max()
        assert np.abs(st - ref[key + 'stress']).max() <= 1e-9 * np.abs(ref[key + 'stress']).max()
