"""Spectral tools: pad_gradient / pad_laplacian through the C ABI and the API helpers
grad_i / grad_dot_grad / laplacian / reduced_* on CUDA against the CPU oracle (functional_tools.py:166-287), on even,
odd, mixed and skewed grids -- the cases where the reference's Nyquist convention makes i*k*F non-Hermitian."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(8, 6, 10), (7, 9, 5), (6, 8, 7), (12, 12, 12), (16, 9, 128), (64, 128, 128)]


def _rel(a, b):
    return ((a.cpu() - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize('shape', SHAPES)
def test_native_gradient_and_laplacian_match_oracle(shape):
    from oracle import ofdft_oracle as orc
    import profess_ad_b200.functional_tools as T
    dev = torch.device('cuda:0')
    box, den = orc.synth_rough(shape, seed=sum(shape))
    g = orc.Grid(box, shape)
    gx, gy, gz = T.spectral_gradient(box.to(dev), den.to(dev))
    for a, b in zip((gx, gy, gz), g.grad(den)):
        assert _rel(a, b) < 1e-12
    assert _rel(T.spectral_laplacian(box.to(dev), den.to(dev)), g.laplacian(den)) < 1e-12


@pytest.mark.parametrize('shape', SHAPES[:5])
def test_api_helpers_on_cuda_match_oracle(shape):
    """grad_i & co. take explicit k tensors (user functionals): on CUDA they must give the reference's CPU numbers,
    i.e. the Hermitian part on the self-conjugate planes, not whatever cuFFT's c2r does with non-Hermitian input."""
    import numpy as np
    from oracle import ofdft_oracle as orc
    import profess_ad_b200.functional_tools as T
    dev = torch.device('cuda:0')
    box, den = orc.synth_rough(shape, seed=3 + sum(shape))
    g = orc.Grid(box, shape)
    b, d = box.to(dev), den.to(dev)
    kx, ky, kz, k2 = T.wavevecs(b, shape)
    for k_dev, k_ref in zip((kx, ky, kz, k2), (*g.kvec, g.k2)):
        assert _rel(k_dev, k_ref) < 1e-15
    grads = g.grad(den)
    for k, ref in zip((kx, ky, kz), grads):
        assert _rel(T.grad_i(k, d), ref) < 1e-12
    gdg = grads[0] ** 2 + grads[1] ** 2 + grads[2] ** 2
    assert _rel(T.grad_dot_grad(kx, ky, kz, d), gdg) < 1e-12
    lap = g.laplacian(den)
    assert _rel(T.laplacian(k2, d), lap) < 1e-12
    c = 0.25 * (3 * np.pi * np.pi) ** (-2 / 3)
    assert _rel(T.reduced_gradient_squared(kx, ky, kz, d), c * gdg / den.pow(8 / 3)) < 1e-12
    assert _rel(T.reduced_gradient(kx, ky, kz, d), 0.5 * (3 * np.pi * np.pi) ** (-1 / 3) * gdg.clamp(min=0).sqrt() / den.pow(4 / 3)) < 1e-12
    assert _rel(T.reduced_laplacian(k2, d), c * lap / den.pow(5 / 3)) < 1e-12


def test_wgc99_potential_at_128_cubed_matches_oracle():
    """One oracle comparison of dE/dn at a production-size grid through the pipelined kernels (128^3, rough density on a
    skewed cell; ~3 s of CPU).  bench.py does the same at 256^3 with the oracle values of its cpu_baseline leg."""
    from oracle import ofdft_oracle as orc
    import profess_ad_b200.functionals as F
    dev = torch.device('cuda:0')
    box, den = orc.synth_rough((128, 128, 128), seed=5, L=16.0)
    E_ref, V_ref = orc.energy_and_potential(box, den, orc.WangGovindCarter99())
    E, V = F.energy_and_potential(box.to(dev), den.to(dev), F.WangGovindCarter99().forward)
    n_atoms = 32.0
    assert abs(E.item() - E_ref.item()) / n_atoms < 1e-8            # north_star: <= 1e-8 Ha/atom
    assert _rel(V, V_ref) < 1e-9                                    # north_star: <= 1e-9 relative max-abs
