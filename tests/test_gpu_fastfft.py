"""The hand-written fused z-pass FFT pipeline (csrc/fftz.cu) vs library FFTs and vs the oracle."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _plan(box, den):
    from profess_ad_b200 import _native
    return _native.get_plan(box, den), _native


@pytest.mark.parametrize('shape', [(6, 10, 128), (12, 9, 256), (5, 4, 256), (16, 16, 128), (6, 10, 512), (3, 4, 512)])
def test_fast_rfft3_matches_library(shape):
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(sum(shape))
    f = torch.rand(*shape, dtype=torch.double, generator=gen).to(dev)
    box = torch.eye(3, dtype=torch.double, device=dev) * 7.0
    plan, nat = _plan(box, f)
    assert plan.lib.pad_fast_fft_supported(plan.handle) == 1
    nzp = shape[2] // 2 + 8
    spec = torch.zeros(shape[0], shape[1], nzp, dtype=torch.complex128, device=dev)
    nzp_out = ctypes.c_int(0)
    nat.check(plan.lib.pad_rfft3_fast(plan.handle, nat.ptr(f), nat.ptr(spec), ctypes.byref(nzp_out), nat.stream_ptr(dev)))
    assert nzp_out.value == nzp
    ref = torch.fft.rfftn(f)
    nzh = shape[2] // 2 + 1
    err = (spec[:, :, :nzh] - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-14, err
    assert spec[:, :, nzh:].abs().max().item() == 0.0
    out = torch.empty_like(f)
    nat.check(plan.lib.pad_irfft3_fast(plan.handle, nat.ptr(spec), nat.ptr(out), nat.stream_ptr(dev)))
    err = (out / f.numel() - f).abs().max().item()
    assert err < 1e-14, err


@pytest.mark.parametrize('shape,axis', [((64, 128, 128), 0), ((64, 128, 128), 1), ((256, 64, 128), 0), ((128, 256, 256), 1),
                                        ((128, 64, 256), 0), ((64, 64, 128), 1), ((512, 64, 128), 0), ((64, 512, 128), 1)])
def test_strided_axis_pass_matches_library(shape, axis):
    """Own strided x / y pass (csrc/fft_strided.cuh) against torch.fft on the live columns; padding untouched."""
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(7 + sum(shape) + axis)
    nzh, nzp = shape[2] // 2 + 1, shape[2] // 2 + 8
    spec = torch.zeros(shape[0], shape[1], nzp, dtype=torch.complex128, device=dev)
    live = torch.view_as_complex(torch.rand(shape[0], shape[1], nzh, 2, dtype=torch.double, generator=gen) - 0.5).to(dev)
    spec[:, :, :nzh] = live
    spec[:, :, nzh:] = 123.0            # sentinel: the pass must not touch padding columns
    box = torch.eye(3, dtype=torch.double, device=dev) * 7.0
    f = torch.empty(*shape, dtype=torch.double, device=dev)
    plan, nat = _plan(box, f)
    for direction, ref in ((-1, torch.fft.fft(live, dim=axis)), (+1, torch.fft.ifft(live, dim=axis) * shape[axis])):
        work = spec.clone()
        nat.check(plan.lib.pad_fft_axis_fast(plan.handle, nat.ptr(work), axis, direction, nat.stream_ptr(dev)))
        err = (work[:, :, :nzh] - ref).abs().max().item() / ref.abs().max().item()
        assert err < 1e-14, (direction, err)
        assert (work[:, :, nzh:] == 123.0).all()


@pytest.mark.parametrize('shape', [(64, 64, 128), (128, 64, 256), (64, 128, 128), (256, 128, 128), (64, 256, 256),
                                   (64, 512, 128), (512, 128, 128), (64, 64, 512), (128, 128, 512)])
def test_fast_rfft3_own_xy(shape):
    """Full own 3-D transform (z pass + strided y and x passes, Nyquist column packed 8 lines per tile)."""
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(sum(shape))
    f = torch.rand(*shape, dtype=torch.double, generator=gen).to(dev)
    box = torch.eye(3, dtype=torch.double, device=dev) * 7.0
    plan, nat = _plan(box, f)
    nzh, nzp = shape[2] // 2 + 1, shape[2] // 2 + 8
    spec = torch.zeros(shape[0], shape[1], nzp, dtype=torch.complex128, device=dev)
    nat.check(plan.lib.pad_rfft3_fast(plan.handle, nat.ptr(f), nat.ptr(spec), None, nat.stream_ptr(dev)))
    ref = torch.fft.rfftn(f)
    err = (spec[:, :, :nzh] - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-14, err
    assert spec[:, :, nzh:].abs().max().item() == 0.0
    out = torch.empty_like(f)
    nat.check(plan.lib.pad_irfft3_fast(plan.handle, nat.ptr(spec), nat.ptr(out), nat.stream_ptr(dev)))
    err = (out / f.numel() - f).abs().max().item()
    assert err < 1e-14, err


@pytest.mark.parametrize('shape,seed', [((10, 12, 128), 31), ((9, 8, 256), 32), ((4, 6, 256), 33),
                                        ((64, 64, 128), 34), ((64, 128, 256), 35), ((128, 64, 128), 36),
                                        ((64, 64, 512), 37), ((512, 64, 128), 38), ((64, 512, 128), 39)])
def test_wgc99_fused_pipeline_matches_oracle_and_plain_path(shape, seed):
    from oracle import ofdft_oracle as orc
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import _native
    lib = _native.load_library()
    dev = torch.device('cuda:0')
    box, den = orc.synth_rough(shape, seed=seed, L=9.0)
    E_ref, V_ref = orc.energy_and_potential(box, den, orc.WangGovindCarter99())
    b, d = box.to(dev), den.to(dev)
    results = {}
    for fast in (1, 0):
        old = lib.pad_set_fast_fft(fast)
        try:
            E, V = F.energy_and_potential(b, d, F.WangGovindCarter99().forward)
            e_only = F.WangGovindCarter99().forward(b, d).item()
        finally:
            lib.pad_set_fast_fft(old)
        results[fast] = (E.item(), V.cpu())
        assert abs(E.item() - E_ref.item()) <= 1e-10 * abs(E_ref.item()), (fast, E.item(), E_ref.item())
        assert ((V.cpu() - V_ref).abs().max() / V_ref.abs().max()).item() < 1e-9, fast
        assert abs(e_only - E.item()) <= 1e-13 * abs(E.item())
    assert abs(results[1][0] - results[0][0]) <= 1e-12 * abs(results[0][0])


@pytest.mark.parametrize('shape,seed', [((64, 64, 128), 61), ((64, 128, 256), 62), ((128, 64, 128), 63)])
def test_pbe_fused_pipeline_matches_oracle_and_plain_path(shape, seed):
    """PerdewBurkeErnzerhof, pbe_exchange, pbe_correlation on the fused passes (gradient and divergence as one-in / three-out
    and three-in / one-out forms of the fused x pass, PBE point math + forward z transform of w inside the inverse z pass)
    against the cuFFT route and the oracle; energy-only and accumulating calls."""
    from oracle import ofdft_oracle as orc
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import _native as nat
    lib = nat.load_library()
    dev = torch.device('cuda:0')
    box, den = orc.synth_rough(shape, seed=seed, L=9.0)
    b, d = box.to(dev), den.to(dev)
    plan = nat.get_plan(b, d)
    for which, f, fo in ((3, F.PerdewBurkeErnzerhof, orc.PerdewBurkeErnzerhof), (1, F.pbe_exchange, orc.pbe_exchange),
                         (2, F.pbe_correlation, orc.pbe_correlation)):
        E_ref, V_ref = orc.energy_and_potential(box, den, fo)
        res = {}
        for fast in (1, 0):
            old = lib.pad_set_option(b'pbe_fast', fast)
            try:
                f0 = lib.pad_fft_exec_count()
                E, V = F.energy_and_potential(b, d, f)
                used_cufft = lib.pad_fft_exec_count() - f0
                e_only = f(b, d).item()
                Eacc = torch.full((), -0.5, dtype=torch.double, device=dev)
                vacc = torch.full_like(d, 0.125)
                nat.check(lib.pad_eval_pbe(plan.handle, nat.ptr(d), which, nat.ptr(Eacc), nat.ptr(vacc), 1, nat.stream_ptr(dev)))
            finally:
                lib.pad_set_option(b'pbe_fast', old)
            assert (used_cufft == 0) == bool(fast), (which, fast, used_cufft)
            assert abs(E.item() - E_ref.item()) <= 1e-10 * abs(E_ref.item()), (which, fast)
            assert ((V.cpu() - V_ref).abs().max() / V_ref.abs().max()).item() < 1e-9, (which, fast)
            assert abs(e_only - E.item()) <= 1e-13 * abs(E.item())
            assert abs(Eacc.item() + 0.5 - E.item()) <= 1e-12 * abs(E.item()) + 1e-15
            assert ((vacc - 0.125 - V).abs().max() / V.abs().max()).item() < 1e-12
            res[fast] = (E.item(), V)
        assert abs(res[1][0] - res[0][0]) <= 1e-12 * abs(res[0][0])
        assert ((res[1][1] - res[0][1]).abs().max() / res[0][1].abs().max()).item() < 1e-11


@pytest.mark.parametrize('shape,seed', [((64, 64, 128), 51), ((128, 64, 256), 52), ((64, 64, 512), 53)])
def test_hartree_fused_pipeline_matches_oracle_and_plain_path(shape, seed):
    """Hartree alone on the fused passes (z r2c -> y -> x . 4 pi / k^2 . x^-1 -> y^-1 -> z c2r + energy + potential), also
    accumulating into an existing potential, against the cuFFT route and the oracle."""
    from oracle import ofdft_oracle as orc
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import _native as nat
    lib = nat.load_library()
    dev = torch.device('cuda:0')
    box, den = orc.synth_rough(shape, seed=seed, L=9.0)
    E_ref, V_ref = orc.energy_and_potential(box, den, orc.Hartree)
    b, d = box.to(dev), den.to(dev)
    res = {}
    for fast in (1, 0):
        old = lib.pad_set_fast_fft(fast)
        try:
            f0 = lib.pad_fft_exec_count()
            E, V = F.energy_and_potential(b, d, F.Hartree)
            used_cufft = lib.pad_fft_exec_count() - f0
            plan = nat.get_plan(b, d)
            Eacc = torch.full((), 1.5, dtype=torch.double, device=dev)
            vacc = torch.full_like(d, 0.25)
            nat.check(lib.pad_eval_hartree(plan.handle, nat.ptr(d), nat.ptr(Eacc), nat.ptr(vacc), 1, nat.stream_ptr(dev)))
        finally:
            lib.pad_set_fast_fft(old)
        assert (used_cufft == 0) == bool(fast)
        assert abs(E.item() - E_ref.item()) <= 1e-10 * abs(E_ref.item()), fast
        assert ((V.cpu() - V_ref).abs().max() / V_ref.abs().max()).item() < 1e-9, fast
        assert abs(Eacc.item() - 1.5 - E.item()) <= 1e-12 * abs(E.item()) + 1e-15          # (1.5 + E rounds at 2e-16)
        assert ((vacc - 0.25 - V).abs().max() / V.abs().max()).item() < 1e-12
        res[fast] = E.item()
    assert abs(res[1] - res[0]) <= 1e-12 * abs(res[0])


@pytest.mark.parametrize('expo', [(5 + 5 ** 0.5) / 6, (5 - 5 ** 0.5) / 6, 2.0 / 3.0, -0.5393])
def test_fast_math_accuracy(expo):
    """Table-driven pow / sqrt / reciprocal (csrc/fastmath.cuh) against torch fp64 over 60 decades."""
    from profess_ad_b200 import _native as nat
    lib = nat.load_library()
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(11)
    n = 1 << 20
    x = torch.pow(10.0, 60.0 * torch.rand(n, dtype=torch.double, generator=gen) - 40.0).to(dev)
    x[:8] = torch.tensor([1.0, 2.0, 0.5, 1.0 - 2 ** -53, 1.0 + 2 ** -52, 0.0301, 1e-30, 7.3e19], dtype=torch.double)
    out = torch.empty(3 * n, dtype=torch.double, device=dev)
    ref = torch.empty(3 * n, dtype=torch.double, device=dev)
    nat.check(lib.pad_dbg_fastmath(nat.ptr(x), n, expo, nat.ptr(out), nat.ptr(ref), nat.stream_ptr(dev)))
    torch.cuda.synchronize()
    exact = torch.cat([torch.pow(x, expo), torch.sqrt(x), 1.0 / x])
    # exp(e log x) in plain fp64 (what the kernels replaced) is conditioned by |e log x|: weigh the pow error by it
    cond = torch.cat([1.0 + (expo * torch.log(x)).abs(), torch.ones(2 * n, dtype=torch.double, device=dev)])
    err = ((out - exact).abs() / exact.abs() / cond).view(3, n).max(dim=1).values.tolist()
    err_lib = ((ref - exact).abs() / exact.abs()).view(3, n).max(dim=1).values.tolist()
    assert err[0] < 4e-16, (err, err_lib)
    assert err[1] < 2.5e-16, (err, err_lib)        # sqrt: correctly rounded or 1 ulp
    assert err[2] < 7e-16, (err, err_lib)          # 1/x as rsqrt^2


@pytest.mark.parametrize('shape,seed', [((64, 64, 128), 41), ((64, 128, 256), 42), ((128, 64, 128), 43)])
def test_wt_family_fused_pipeline_matches_oracle_and_plain_path(shape, seed):
    """Wang-Teter family (alpha == beta: WT, Perrot, SM; alpha != beta: WGC98) on the fused FFT pipeline (one round trip,
    Lindhard kernel evaluated inside the x pass) against the oracle and against the cuFFT path."""
    from oracle import ofdft_oracle as orc
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import _native
    lib = _native.load_library()
    dev = torch.device('cuda:0')
    box, den = orc.synth_rough(shape, seed=seed, L=9.0)
    b, d = box.to(dev), den.to(dev)
    v0 = torch.rand(*shape, dtype=torch.double, generator=torch.Generator().manual_seed(seed)).to(dev)
    for name, f, fo in (('WT', F.WangTeter, orc.WangTeter), ('SM', F.SmargiassiMadden, orc.SmargiassiMadden),
                        ('WGC98', F.WangGovindCarter98, orc.WangGovindCarter98)):
        E_ref, V_ref = orc.energy_and_potential(box, den, fo)
        res = {}
        for fast in (1, 0):
            old = lib.pad_set_fast_fft(fast)
            try:
                E, V = F.energy_and_potential(b, d, f)
                e_only = f(b, d).item()
            finally:
                lib.pad_set_fast_fft(old)
            res[fast] = E.item()
            assert abs(E.item() - E_ref.item()) <= 1e-10 * abs(E_ref.item()), (name, fast, E.item(), E_ref.item())
            assert ((V.cpu() - V_ref).abs().max() / V_ref.abs().max()).item() < 1e-9, (name, fast)
            assert abs(e_only - E.item()) <= 1e-13 * abs(E.item())
        assert abs(res[1] - res[0]) <= 1e-12 * abs(res[0])
    # accumulate path (fused term list): E += , v +=
    from profess_ad_b200 import _density_opt as D
    T = D.describe_terms([F.Hartree, F.WangGovindCarter98, F.PerdewZunger])
    out = {}
    for fast in (1, 0):
        old = lib.pad_set_fast_fft(fast)
        try:
            E, v = D.eval_total(b, d, None, T)
        finally:
            lib.pad_set_fast_fft(old)
        out[fast] = (E.item(), v.clone())
    assert abs(out[1][0] - out[0][0]) <= 1e-12 * abs(out[0][0])
    assert ((out[1][1] - out[0][1]).abs().max() / out[0][1].abs().max()).item() < 1e-11


def test_full_size_properties_256():
    """BASELINE.json's headline size (256^3), where the CPU oracle takes tens of seconds per evaluation: properties
    that do not depend on the size.  (1) supercell consistency: tiling a 128^3 cell twice per axis multiplies every
    energy by 8 and tiles the potential; (2) translation by whole grid points leaves E unchanged and translates v;
    (3) the potential is the derivative of the energy along a random direction (central difference)."""
    import profess_ad_b200.functionals as F
    from profess_ad_b200.synthetic import smooth_supercell
    dev = torch.device('cuda:0')
    box1, den1 = smooth_supercell(128, 2, device=dev)
    gen = torch.Generator().manual_seed(3)
    den1 = den1 * (1 + 0.05 * torch.rand(128, 128, 128, dtype=torch.double, generator=gen).to(dev))   # no symmetry left
    # WGC99 rounds the electron number to an integer (functionals.py:952): keep it integral in both cells (96 -> 768)
    den1 = den1 * (96.0 / (den1.sum().item() * abs(torch.linalg.det(box1).item()) / den1.numel()))
    box2, den2 = 2 * box1, den1.repeat(2, 2, 2).contiguous()
    dV = abs(torch.linalg.det(box2).item()) / den2.numel()
    for name, f in (('WGC99', F.WangGovindCarter99().forward), ('WT', F.WangTeter), ('WGC98', F.WangGovindCarter98),
                    ('Hartree', F.Hartree), ('PBE', F.PerdewBurkeErnzerhof)):
        E1, V1 = F.energy_and_potential(box1, den1, f)
        E2, V2 = F.energy_and_potential(box2, den2, f)
        assert abs(E2.item() - 8 * E1.item()) <= 2e-12 * abs(E2.item()), (name, E2.item(), 8 * E1.item())
        assert ((V2 - V1.repeat(2, 2, 2)).abs().max() / V1.abs().max()).item() < 1e-10, name
        shift = (37, 101, 6)
        E3, V3 = F.energy_and_potential(box2, torch.roll(den2, shift, (0, 1, 2)).contiguous(), f)
        assert abs(E3.item() - E2.item()) <= 2e-12 * abs(E2.item()), name
        assert ((V3 - torch.roll(V2, shift, (0, 1, 2))).abs().max() / V2.abs().max()).item() < 1e-10, name
        if name in ('WGC99', 'WT', 'PBE'):
            # a direction that conserves the electron number: n0 = N / vol is a detached constant in the reference's
            # potentials (functionals.py:634-647), so only such directions see v as the full derivative; it is
            # correlated with v so that the derivative is not a small difference of large sums
            W = V2 / V2.abs().max()
            delta = 0.01 * den2.mean() * (W - W.mean())
            eps = 1e-3
            Ep = f(box2, (den2 + eps * delta).contiguous()).item()
            Em = f(box2, (den2 - eps * delta).contiguous()).item()
            lhs = (Ep - Em) / (2 * eps)
            rhs = (V2 * delta).sum().item() * dV
            assert abs(lhs - rhs) <= 1e-7 * abs(rhs) + 1e-12, (name, lhs, rhs)


@pytest.mark.parametrize('shape,lpi,tpi', [((64, 128, 128), 0, 0), ((128, 256, 256), 0, 0), ((64, 256, 128), 16, 1),
                                           ((64, 128, 256), 32, 7), ((256, 128, 128), 64, 3)])
def test_pipelined_zy_kernels_match_per_pass_kernels(shape, lpi, tpi):
    """Software-pipelined (z, y) kernels (csrc/zy_pipe.cuh: z items and y items of an x-plane in ONE persistent kernel,
    plane handed over through the L2) against the one-kernel-per-pass path, for the plain transforms, WGC99 and the
    Wang-Teter family, with several item sizes; the watchdog word of the control block must stay 0."""
    from oracle import ofdft_oracle as orc
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import _native as nat
    lib = nat.load_library()
    dev = torch.device('cuda:0')
    box, den = orc.synth_rough(shape, seed=sum(shape) + lpi, L=9.0)
    b, d = box.to(dev), den.to(dev)
    plan = nat.get_plan(b, d)
    nzh, nzp = shape[2] // 2 + 1, shape[2] // 2 + 8
    res = {}
    for pipe in (1, 0):
        olds = [lib.pad_set_option(b'pipe', pipe), lib.pad_set_option(b'pipe_lpi', lpi), lib.pad_set_option(b'pipe_tpi', tpi)]
        try:
            assert lib.pad_pipe_supported(plan.handle) == pipe
            spec = torch.zeros(shape[0], shape[1], nzp, dtype=torch.complex128, device=dev)
            nat.check(lib.pad_rfft3_fast(plan.handle, nat.ptr(d), nat.ptr(spec), None, nat.stream_ptr(dev)))
            ref = torch.fft.rfftn(d)
            assert ((spec[:, :, :nzh] - ref).abs().max() / ref.abs().max()).item() < 1e-14, pipe
            assert spec[:, :, nzh:].abs().max().item() == 0.0
            back = torch.empty_like(d)
            nat.check(lib.pad_irfft3_fast(plan.handle, nat.ptr(spec.clone()), nat.ptr(back), nat.stream_ptr(dev)))
            assert ((back / d.numel() - d).abs().max() / d.abs().max()).item() < 1e-14, pipe
            out = {}
            for name, f in (('WGC99', F.WangGovindCarter99().forward), ('WT', F.WangTeter), ('WGC98', F.WangGovindCarter98)):
                for rep in range(3):        # the control block must come back clean launch after launch
                    E, V = F.energy_and_potential(b, d, f)
                out[name] = (E.item(), V.clone())
            res[pipe] = out
            if pipe:
                assert lib.pad_pipe_status(plan.handle, nat.stream_ptr(dev)) == 0
        finally:
            for nm, o in zip((b'pipe', b'pipe_lpi', b'pipe_tpi'), olds):
                lib.pad_set_option(nm, o)
    for name in res[1]:
        e1, v1 = res[1][name]
        e0, v0 = res[0][name]
        assert abs(e1 - e0) <= 2e-14 * abs(e0), (name, e1, e0)
        assert ((v1 - v0).abs().max() / v0.abs().max()).item() < 1e-13, name
    E_ref, V_ref = orc.energy_and_potential(box, den, orc.WangGovindCarter99())
    assert abs(res[1]['WGC99'][0] - E_ref.item()) <= 1e-10 * abs(E_ref.item())
    assert ((res[1]['WGC99'][1].cpu() - V_ref).abs().max() / V_ref.abs().max()).item() < 1e-9


@pytest.mark.parametrize('shape,terms,pipe', [((64, 128, 128), 'IHWP', 0), ((128, 128, 256), 'IHWP', 0), ((64, 128, 128), 'HW', 0),
                                              ((64, 64, 128), 'IWP', 0), ((64, 128, 128), 'IHWB', 0), ((64, 128, 128), 'IHWP', 1),
                                              ((64, 128, 128), 'IWP', 1)])
def test_fused_term_list_matches_term_by_term_and_oracle(shape, terms, pipe):
    """pad_eval_total with WGC99 as the kinetic term: IonElectron / LDA-x / PZ-c folded into the WGC99 mid pass, Hartree as a
    fourth field of its second transform batch (system.py:759-772 in ONE sweep) against the term-by-term calls and the oracle."""
    from oracle import ofdft_oracle as orc
    import profess_ad_b200.functionals as F
    from profess_ad_b200 import _density_opt as D, _native as nat
    lib = nat.load_library()
    dev = torch.device('cuda:0')
    box, den = orc.synth_rough(shape, seed=sum(shape) + len(terms), L=9.0)
    v_ext = -0.2 * torch.rand(*shape, dtype=torch.double, generator=torch.Generator().manual_seed(5))
    b, d, vx = box.to(dev), den.to(dev), v_ext.to(dev)
    table = {'I': (F.IonElectron, None), 'H': (F.Hartree, orc.Hartree), 'W': (F.WangGovindCarter99().forward, orc.WangGovindCarter99()),
             'P': (F.PerdewZunger, orc.PerdewZunger), 'B': (F.PerdewBurkeErnzerhof, orc.PerdewBurkeErnzerhof)}
    T = D.describe_terms([table[c][0] for c in terms])
    out = {}
    # 1: fused list, local terms in the one-field Hartree inverse pass (default); 2: fused list, local terms inside the mid
    # pass; 0: term by term
    for variant, (fuse, tail) in {1: (1, 1), 2: (1, 0), 0: (0, 1)}.items():
        old = lib.pad_set_option(b'fuse_terms', fuse)
        old_pipe = lib.pad_set_option(b'pipe', pipe)
        old_tail = lib.pad_set_option(b'local_tail', tail)
        try:
            for _ in range(3):          # direct, captured, replayed
                E, v = D.eval_total(b, d, vx if 'I' in terms else None, T)
            out[variant] = (E.item(), v.clone())
        finally:
            lib.pad_set_option(b'fuse_terms', old)
            lib.pad_set_option(b'pipe', old_pipe)
            lib.pad_set_option(b'local_tail', old_tail)
    for variant in (1, 2):
        assert abs(out[variant][0] - out[0][0]) <= 1e-12 * abs(out[0][0]), (variant, out[variant][0], out[0][0])
        assert ((out[variant][1] - out[0][1]).abs().max() / out[0][1].abs().max()).item() < 1e-11, variant
    E_ref, V_ref = 0.0, torch.zeros_like(den)
    for c in terms:
        if c == 'I':
            e, v = orc.energy_and_potential(box, den, lambda bb, nn: orc.IonElectron(bb, nn, v_ext))
        else:
            e, v = orc.energy_and_potential(box, den, table[c][1])
        E_ref, V_ref = E_ref + e.item(), V_ref + v
    assert abs(out[1][0] - E_ref) <= 1e-10 * abs(E_ref)
    assert ((out[1][1].cpu() - V_ref).abs().max() / V_ref.abs().max()).item() < 1e-9


def test_repeated_evaluation_replays_a_cuda_graph_with_identical_results():
    """pad_eval_wgc99 with unchanged arguments: direct on the first call, captured on the second, replayed as one graph launch
    afterwards -- same bits every time, the launch counter keeps counting the kernels of the graph, a changed density BUFFER or
    option 'graphs' = 0 goes back to direct launches, and a changed density in the SAME buffer is picked up by the replay."""
    from oracle import ofdft_oracle as orc
    from profess_ad_b200 import _native as nat
    lib = nat.load_library()
    dev = torch.device('cuda:0')
    shape = (64, 64, 128)
    box, den = orc.synth_rough(shape, seed=21, L=9.0)
    b, d = box.to(dev), den.to(dev).contiguous()
    plan = nat.get_plan(b, d)
    a98, b98 = (5 + 5 ** 0.5) / 6, (5 - 5 ** 0.5) / 6
    E = torch.zeros((), dtype=torch.double, device=dev)
    v = torch.zeros_like(d)

    def call(dd):
        l0 = lib.pad_launch_count()
        nat.check(lib.pad_eval_wgc99(plan.handle, nat.ptr(dd), a98, b98, 2.7, 1.0, nat.ptr(E), nat.ptr(v), 0, nat.stream_ptr(dev)))
        torch.cuda.synchronize()
        return E.item(), v.clone(), lib.pad_launch_count() - l0

    runs = [call(d) for _ in range(5)]
    for e, vv, n in runs[1:]:
        assert e == runs[0][0] and torch.equal(vv, runs[0][1]) and n == runs[0][2] > 0
    E_ref, V_ref = orc.energy_and_potential(box, den, orc.WangGovindCarter99())
    assert abs(runs[-1][0] - E_ref.item()) <= 1e-10 * abs(E_ref.item())
    # new contents in the same buffer: the replayed graph reads them (nothing about the data is baked in)
    d.mul_(1.01)
    e2, v2, _ = call(d)
    d2 = d.clone()
    e3, v3, _ = call(d2)              # different buffer: a direct evaluation
    assert e2 == e3 and torch.equal(v2, v3) and e2 != runs[0][0]
    old = lib.pad_set_option(b'graphs', 0)
    try:
        e4, v4, _ = call(d)
    finally:
        lib.pad_set_option(b'graphs', old)
    assert e4 == e2 and torch.equal(v4, v2)
