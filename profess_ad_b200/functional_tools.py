"""Grid / FFT tools with the reference's names and signatures (src/professad/functional_tools.py).

Hot-path users (the functionals in ``functionals.py``) never materialise k-vectors: the CUDA
kernels derive them from the 9 reciprocal-lattice doubles.  The helpers here exist for API
compatibility with user code and tests written against the reference:

* ``get_functional_derivative``  (functional_tools.py:9-31)   -- autograd through our custom Functions
* ``wavevecs``                   (functional_tools.py:135-162) -- explicit k-vector tensors
* ``grad_i`` / ``laplacian`` ... (functional_tools.py:166-287) -- for a native (box, f) pair use
  ``spectral_gradient`` / ``spectral_laplacian`` below, which call the C ABI
* ``interpolate`` / ``interpolate_kernel`` / ``field_dependent_convolution`` (functional_tools.py:292-423)
"""
import numpy as np
import torch

from . import _native


def get_functional_derivative(box_vecs, den, functional, requires_grad=False):
    """delta F / delta n via autograd (functional_tools.py:9-31).  For the native functionals the
    "autograd" step is a custom backward that hands back the analytic potential the forward kernel
    already produced."""
    if requires_grad:
        raise NotImplementedError('second derivatives (requires_grad=True) are outside the B200 hot path')
    den.requires_grad = True
    try:
        grad = torch.autograd.grad(functional(box_vecs, den), den)[0]
    finally:
        den.requires_grad = False
    return grad / (torch.abs(torch.linalg.det(box_vecs)) / den.numel())


def get_stress(box_vecs, den, functional, requires_grad=False):
    """Functional contribution to the stress, (1/vol) dF/d eps at fixed electron number (functional_tools.py:73-100).
    The reference differentiates through ``box_vecs``; for the native functionals this is one C-ABI call evaluating
    the analytic expressions (csrc/stress.cu).  User-defined Python functionals are outside this path."""
    if requires_grad:
        raise NotImplementedError('second derivatives (requires_grad=True) are outside the B200 hot path')
    from . import _density_opt
    name = getattr(functional, '__qualname__', '') or getattr(functional, '__name__', '')
    T = _density_opt.describe_terms([functional], den.device) if name not in ('IonElectron', 'IonIon') else None
    if T is None:
        raise NotImplementedError('get_stress: only native functionals of (box_vecs, den) have an analytic stress here')
    return _density_opt.stress_terms(box_vecs, den, T)


def get_pressure(box_vecs, den, functional, requires_grad=False):
    """-dF/dvol at fixed electron number (functional_tools.py:103-127) = -trace(stress) / 3."""
    return -torch.trace(get_stress(box_vecs, den, functional, requires_grad)) / 3


def wavevecs(box_vecs, shape):
    """k_x, k_y, k_z, k^2 on the half-spectrum grid (functional_tools.py:135-162), Nyquist index
    positive on axes 0 and 1."""
    b = 2 * np.pi * torch.linalg.inv(box_vecs.T)
    assert not torch.any(torch.isnan(b)), 'Lattice vector matrix is not invertible.'
    idx = []
    for ax in range(2):
        n = int(shape[ax])
        j = torch.fft.fftfreq(n, dtype=torch.double, device=box_vecs.device) * n
        j[n // 2] = j[n // 2].abs()
        idx.append(j)
    idx.append(torch.fft.rfftfreq(int(shape[2]), dtype=torch.double, device=box_vecs.device) * int(shape[2]))
    nA, nB, nC = torch.meshgrid(*idx, indexing='ij')
    kx, ky, kz = (nA * b[0, c] + nB * b[1, c] + nC * b[2, c] for c in range(3))
    return kx, ky, kz, kx.pow(2) + ky.pow(2) + kz.pow(2)


def hermitian_symmetrize(spec, shape):
    """Hermitian part of a half spectrum on its two self-conjugate planes (j2 = 0, and j2 = n2/2 for even n2):
    S(p) <- (S(p) + conj S(pbar)) / 2, pbar = (-j0, -j1, j2).

    The multipliers of ``wavevecs`` (Nyquist index made positive on axes 0, 1) are not Hermitian on even grids, so
    ``i k F`` and, on skewed cells, ``k^2 F`` are not the transform of a real field.  The reference's CPU ``irfftn``
    (c2c over axes 0, 1, then c2r over axis 2) silently keeps exactly this Hermitian part; cuFFT's c2r is undefined
    for such input.  Making it explicit gives the reference's numbers on any backend."""
    n2 = int(shape[2])
    planes = [0] + ([n2 // 2] if n2 % 2 == 0 and n2 > 1 else [])
    out = spec.clone()
    for j2 in planes:
        P = spec[:, :, j2]
        partner = torch.roll(torch.flip(P, (0, 1)), (1, 1), (0, 1)).conj()
        out[:, :, j2] = 0.5 * (P + partner)
    return out


def _irfftn(spec, shape):
    """irfftn with the reference's CPU semantics on every device (see hermitian_symmetrize)."""
    return torch.fft.irfftn(hermitian_symmetrize(spec, shape), tuple(int(n) for n in shape))


def hermitian_symmetrize_nodes(spec, shape):
    """hermitian_symmetrize for a stack of half spectra (n0, n1, n2/2 + 1, n_nodes)"""
    n2 = int(shape[2])
    planes = [0] + ([n2 // 2] if n2 % 2 == 0 and n2 > 1 else [])
    out = spec.clone()
    for j2 in planes:
        P = spec[:, :, j2]
        partner = torch.roll(torch.flip(P, (0, 1)), (1, 1), (0, 1)).conj()
        out[:, :, j2] = 0.5 * (P + partner)
    return out


def grad_i(ki, f):
    """functional_tools.py:166-183 (explicit k_i tensor, library FFT: compatibility helper for user functionals;
    the native path is ``spectral_gradient`` / pad_gradient)."""
    return _irfftn(1j * ki * torch.fft.rfftn(f), f.shape)


def grad_dot_grad(kx, ky, kz, f):
    """functional_tools.py:186-206"""
    F = torch.fft.rfftn(f)
    g = [_irfftn(1j * k * F, f.shape) for k in (kx, ky, kz)]
    return g[0] * g[0] + g[1] * g[1] + g[2] * g[2]


def laplacian(k2, f):
    """functional_tools.py:209-227"""
    return _irfftn(-k2 * torch.fft.rfftn(f), f.shape)


def reduced_gradient(kx, ky, kz, den):
    """functional_tools.py:230-249"""
    gdg = grad_dot_grad(kx, ky, kz, den)
    return 0.5 * (3 * np.pi * np.pi) ** (-1 / 3) * torch.sqrt(torch.clamp(gdg, min=0.0)) / den.pow(4 / 3)


def reduced_gradient_squared(kx, ky, kz, den):
    """functional_tools.py:252-268"""
    return 0.25 * (3 * np.pi * np.pi) ** (-2 / 3) * grad_dot_grad(kx, ky, kz, den) / den.pow(8 / 3)


def reduced_laplacian(k2, den):
    """functional_tools.py:271-287"""
    return 0.25 * (3 * np.pi * np.pi) ** (-2 / 3) * laplacian(k2, den) / den.pow(5 / 3)


def spectral_gradient(box_vecs, f):
    """(df/dx, df/dy, df/dz) through the C ABI (pad_gradient): one r2c, one fused i*k multiply, three c2r."""
    _native.require_cuda(f, 'f')
    f = f.contiguous()
    plan = _native.get_plan(box_vecs, f)
    out = [torch.empty_like(f) for _ in range(3)]
    _native.check(plan.lib.pad_gradient(plan.handle, _native.ptr(f), _native.ptr(out[0]), _native.ptr(out[1]),
                                        _native.ptr(out[2]), _native.stream_ptr(f.device)))
    return tuple(out)


def spectral_laplacian(box_vecs, f):
    """Laplacian through the C ABI (pad_laplacian)."""
    _native.require_cuda(f, 'f')
    f = f.contiguous()
    plan = _native.get_plan(box_vecs, f)
    out = torch.empty_like(f)
    _native.check(plan.lib.pad_laplacian(plan.handle, _native.ptr(f), _native.ptr(out), _native.stream_ptr(f.device)))
    return out


def _hermite(t):
    t2 = t * t
    t3 = t2 * t
    return 1 - 3 * t2 + 2 * t3, t - 2 * t2 + t3, 3 * t2 - 2 * t3, t3 - t2


def interpolate(x, y, xs):
    """1-D cubic Hermite interpolation of y(x) at xs (functional_tools.py:292-334): interior slopes
    are the mean of the adjacent secants, end slopes one-sided."""
    sec = (y[1:] - y[:-1]) / (x[1:] - x[:-1])
    m = torch.cat([sec[:1], 0.5 * (sec[1:] + sec[:-1]), sec[-1:]])
    idx = torch.searchsorted(x[1:], xs)
    dx = x[idx + 1] - x[idx]
    h00, h10, h01, h11 = _hermite((xs - x[idx]) / dx)
    return h00 * y[idx] + h10 * m[idx] * dx + h01 * y[idx + 1] + h11 * m[idx + 1] * dx


def interpolate_kernel(xi_sparse, f, xis):
    """Per-voxel cubic Hermite along the node axis (functional_tools.py:337-378)."""
    dn = xi_sparse[1:] - xi_sparse[:-1]
    sec = (f[..., 1:] - f[..., :-1]) / dn
    m = torch.cat([sec[..., :1], 0.5 * (sec[..., 1:] + sec[..., :-1]), sec[..., -1:]], dim=-1)
    idx = torch.searchsorted(xi_sparse[1:], xis)
    dx = xi_sparse[idx + 1] - xi_sparse[idx]
    h00, h10, h01, h11 = _hermite((xis - xi_sparse[idx]) / dx)

    def pick(a, i):
        return torch.gather(a, 3, i.unsqueeze(3))[..., 0]
    return h00 * pick(f, idx) + h10 * pick(m, idx) * dx + h01 * pick(f, idx + 1) + h11 * pick(m, idx + 1) * dx


def xi_nodes(xi_min, xi_max, kappa, mode, device=None):
    """Node list of functional_tools.py:406-417 for the spline over xi."""
    if mode == 'arithmetic':
        lower = (np.floor(xi_min / kappa) - 3) * kappa
        upper = (np.ceil(xi_max / kappa) + 3) * kappa
        nodes = torch.arange(lower, upper, kappa, dtype=torch.double, device=device)
        nodes[nodes == 0] = xi_min
        return nodes
    if mode == 'geometric':
        assert kappa > 1, 'κ > 1 for geometric progression based spline for field_dependent_convolution'
        lower = kappa ** (-(np.ceil(-np.log(xi_min) / np.log(kappa)) + 3))
        count = np.ceil(np.log((xi_max + 1) / lower) / np.log(kappa)) + 3
        return lower * kappa ** torch.arange(count, dtype=torch.double, device=device)
    raise ValueError('Parameter \'mode\' can only be \'arithmetic\' or \'geometric\'')


def field_dependent_convolution(k, f_tilde, g, xis, kappa, mode='arithmetic'):
    """K(r) = int f(|r - r'|, xi(r)) g(r') dr' by a spline over xi (functional_tools.py:381-423).
    Generic (user-supplied ``f_tilde`` callable) version; the Huang-Carter functionals use the fused
    native path instead."""
    nodes = xi_nodes(xis.min().item(), xis.max().item(), kappa, mode, device=xis.device)
    g_ft = torch.fft.rfftn(g).unsqueeze(3)
    conv = torch.fft.irfftn(hermitian_symmetrize_nodes(f_tilde(k, nodes) * g_ft, g.shape), s=g.shape, dim=(0, 1, 2))
    return interpolate_kernel(nodes, conv, xis)
