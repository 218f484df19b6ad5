"""CPU oracle for the OFDFT hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU (torch fp64 + autograd, MKL FFT) restatement of the algorithms of the
PROFESS-AD reference for the path SURVEY.md section 8 names.  It is the *checker*: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``profess_ad_b200/`` imports it, and the product
path never falls back to it.

Why torch-on-CPU and not numpy: the reference *is* torch fp64 on CPU with potentials taken by
``torch.autograd``.  Restating it with the same primitives keeps the oracle's potentials
independent of the hand-derived analytic potentials the CUDA kernels implement, and makes the
multi-threaded CPU baseline representative of what a reference user runs.

Pinning: ``tests/golden/*.npz`` hold energies/potentials produced by importing the unmodified
reference in the build container (``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py``
checks this file against them and against the known answers of the reference's own tests
(``tests/test_match_profess4.py:23,35``).  Huang-Carter kernels depend on ``xitorch.solve_ivp``
(not installed, un-pinned in ``pyproject.toml:22``): for HC / revHC the omega(eta) table is
*injected* into both sides, and that row is "parity unpinned" at the ODE boundary.

All ``file:line`` citations are relative to ``/root/reference/``.
"""
import math

import numpy as np
import torch

DT = torch.double
PI = math.pi
C_TF = 0.3 * (3.0 * PI * PI) ** (2.0 / 3.0)
EV_PER_HA = 4.3597447222071e-18 / 1.602176634e-19   # system.py:30-31
A_PER_BOHR = 5.29177210903e-11 * 1e10               # system.py:27-28


# ----------------------------------------------------------------------------------------------
#  grid / reciprocal space
# ----------------------------------------------------------------------------------------------
class Grid:
    """Reciprocal-space bookkeeping for one (box, shape).  Follows functional_tools.py:135-162:
    b = 2 pi inv(box^T); integer frequencies with the Nyquist index made positive on axes 0 and 1;
    rfft frequencies on axis 2."""

    def __init__(self, box, shape):
        self.box = box
        self.shape = tuple(int(s) for s in shape)
        self.vol = torch.abs(torch.linalg.det(box))
        self.npts = int(np.prod(self.shape))
        self.dV = self.vol / self.npts
        recip = 2.0 * PI * torch.linalg.inv(box.T)
        freqs = []
        for ax in range(2):
            n = self.shape[ax]
            f = torch.fft.fftfreq(n, dtype=DT) * n
            f[n // 2] = f[n // 2].abs()
            freqs.append(f)
        n2 = self.shape[2]
        freqs.append(torch.fft.rfftfreq(n2, dtype=DT) * n2)
        mA, mB, mC = torch.meshgrid(*freqs, indexing='ij')
        self.kvec = [mA * recip[0, c] + mB * recip[1, c] + mC * recip[2, c] for c in range(3)]
        self.k2 = self.kvec[0] ** 2 + self.kvec[1] ** 2 + self.kvec[2] ** 2
        kabs = torch.zeros_like(self.k2)
        nz = self.k2 != 0
        kabs[nz] = torch.sqrt(self.k2[nz])
        self.kabs = kabs

    def fwd(self, f):
        return torch.fft.rfftn(f)

    def inv(self, F):
        return torch.fft.irfftn(F, self.shape)

    def integral(self, f):
        return torch.mean(f) * self.vol

    def grad(self, f):
        """functional_tools.py:166-183, with rfftn(f) shared between components (same numbers)."""
        F = self.fwd(f)
        return [self.inv(1j * k * F) for k in self.kvec]

    def laplacian(self, f):
        """functional_tools.py:209-227"""
        return self.inv(-self.k2 * self.fwd(f))


def wavevecs(box, shape):
    g = Grid(box, shape)
    return g.kvec[0], g.kvec[1], g.kvec[2], g.k2


# ----------------------------------------------------------------------------------------------
#  local and Coulomb terms
# ----------------------------------------------------------------------------------------------
def IonElectron(box, den, v_ext):
    """functionals.py:31-46"""
    return torch.mean(den * v_ext) * torch.abs(torch.linalg.det(box))


def Hartree(box, den):
    """functionals.py:49-72 : phi = irfftn(4 pi / k^2 * rfftn(n)), k=0 dropped; E = 1/2 int n phi."""
    g = Grid(box, den.shape)
    coul = torch.zeros_like(g.k2)
    nz = g.k2 != 0
    coul[nz] = 4.0 * PI / g.k2[nz]
    phi = g.inv(g.fwd(den) * coul)
    return 0.5 * g.integral(den * phi)


def ThomasFermi(box, den):
    """functionals.py:207-224"""
    return torch.mean(C_TF * den.pow(5.0 / 3.0)) * torch.abs(torch.linalg.det(box))


def Weizsaecker(box, den):
    """functionals.py:227-246 : ked = 1/4 lap(n) - 1/2 sqrt(n) lap(sqrt(n)); sqrt masked at n == 0."""
    g = Grid(box, den.shape)
    root = torch.zeros_like(den)
    pos = den != 0
    root[pos] = torch.sqrt(den[pos])
    ked = 0.25 * g.laplacian(den) - 0.5 * root * g.laplacian(root)
    return g.integral(ked)


# ----------------------------------------------------------------------------------------------
#  Lindhard response and Wang-Teter style functionals
# ----------------------------------------------------------------------------------------------
def lindhard_inverse_G(eta):
    """functionals.py:617-628 : 1/2 + (1-eta^2)/(4 eta) ln|(1+eta)/(1-eta)|, =1 at 0, =1/2 at 1."""
    out = torch.empty_like(eta)
    is0, is1 = eta == 0.0, eta == 1.0
    reg = ~(is0 | is1)
    e = eta[reg]
    out[reg] = 0.5 + ((1.0 - e * e) / (4.0 * e)) * torch.log(torch.abs((1.0 + e) / (1.0 - e)))
    out[is0] = 1.0
    out[is1] = 0.5
    return out


def nonlocal_wt_term(box, den, alpha, beta):
    """functionals.py:644-652 with G_inv_lindhard (:631-639).  n0 = N_elec / vol is a detached float."""
    g = Grid(box, den.shape)
    n_elec = (torch.mean(den) * g.vol).item()
    n0 = n_elec / g.vol
    kF = (3.0 * PI * PI * n0) ** (1.0 / 3.0)
    eta = g.kabs / (2.0 * kF)
    ginv = lindhard_inverse_G(eta)
    kern = 5.0 / (9.0 * alpha * beta * n0 ** (alpha + beta - 5.0 / 3.0)) * (1.0 / ginv - 3.0 * eta ** 2 - 1.0)
    conv = g.inv(kern * g.fwd(den.pow(beta) - n0 ** beta))
    return C_TF * g.integral((den.pow(alpha) - n0 ** alpha) * conv)


def _wt_family(alpha, beta):
    def functional(box, den):
        return Weizsaecker(box, den) + ThomasFermi(box, den) + nonlocal_wt_term(box, den, alpha, beta)
    return functional


WangTeter = _wt_family(5.0 / 6.0, 5.0 / 6.0)                                  # functionals.py:655-670
Perrot = _wt_family(1.0, 1.0)                                                  # :673-689
SmargiassiMadden = _wt_family(0.5, 0.5)                                        # :692-707
WangGovindCarter98 = _wt_family((5 + math.sqrt(5)) / 6, (5 - math.sqrt(5)) / 6)  # :710-725


def WangTeterStyle(alpha, beta, f, fprime0):
    """functionals.py:728-782 : vW + TF * f(T_NL / f'(0) / TF)."""
    def functional(box, den):
        tf = ThomasFermi(box, den)
        tnl = nonlocal_wt_term(box, den, alpha, beta) / fprime0
        return Weizsaecker(box, den) + tf * f(tnl / tf)
    return functional


# ----------------------------------------------------------------------------------------------
#  Wang-Govind-Carter 99
# ----------------------------------------------------------------------------------------------
def wgc99_series_coefficients(num_terms):
    """functionals.py:817-843 : recursions for the Taylor coefficients of the Lindhard function
    inside (B_i, powers eta^{2i}) and outside (A_i, powers eta^{-2i}) the Fermi sphere."""
    a = np.zeros(num_terms + 1)
    a[0] = 3.0
    for idx in range(1, num_terms + 1):
        i = idx - 1
        acc = 0.0
        for j in range(-1, i):
            acc += -3.0 * a[j + 1] / (4.0 * (i - j + 1) ** 2 - 1.0)
        a[idx] = acc
    A = a[1:].copy()
    A[0] -= 1.0
    b = np.zeros(num_terms)
    b[0] = 1.0
    for i in range(1, num_terms):
        acc = 0.0
        for j in range(i):
            acc += b[j] / (4.0 * (i - j) ** 2 - 1.0)
        b[i] = acc
    B = b.copy()
    B[0] = 0.0
    B[1] = b[1] - 3.0
    return A, B


def wgc99_kernel(eta, alpha, beta, gamma, num_terms=100):
    """functionals.py:845-939 : w(eta), w'(eta), w''(eta) = homogeneous + particular series solution
    of the WGC99 kernel ODE (Phys. Rev. B 78, 045105)."""
    A, B = wgc99_series_coefficients(num_terms)
    i = np.arange(num_terms, dtype=np.float64)
    u = 3.0 * (alpha + beta) - gamma / 2.0
    v = u * u - 36.0 * alpha * beta
    denA = (u + 2 * i) ** 2 - v
    denB = (u - 2 * i) ** 2 - v
    Sd = np.sum(A / denA - B / denB)
    Ss = -2.0 * np.sum(i * (A / denA + B / denB))
    sgn = np.sign(u)
    if v > 0:
        rv = math.sqrt(v)
        c1 = sgn * ((rv - u) * Sd + Ss)
        c2 = sgn * ((rv + u) * Sd - Ss) / (2.0 * rv)
    elif v == 0:
        c1 = sgn * Sd
        c2 = sgn * (Ss - u * Sd)
    else:
        c1 = sgn * Sd
        c2 = sgn * (Ss - u * Sd) / math.sqrt(-v)

    e = eta.detach().numpy().astype(np.float64)
    inside = e <= 1.0
    nzm = e != 0.0
    if u >= 0:
        C1 = np.where(inside, c1, 0.0)
        C2 = np.where(inside, c2, 0.0)
    else:
        C1 = np.where(inside, 0.0, c1)
        C2 = np.where(inside, 0.0, c2)
    H = [np.zeros_like(e) for _ in range(3)]
    en, C1n, C2n = e[nzm], C1[nzm], C2[nzm]
    if v > 0:
        x, y = u + math.sqrt(v), u - math.sqrt(v)
        H[0][nzm] = C1n * en ** x + C2n * en ** y
        H[1][nzm] = C1n * x * en ** (x - 1) + C2n * y * en ** (y - 1)
        H[2][nzm] = C1n * x * (x - 1) * en ** (x - 2) + C2n * y * (y - 1) * en ** (y - 2)
    elif v == 0:
        le = np.log(en)
        H[0][nzm] = en ** u * (C2n * le + C1n)
        H[1][nzm] = C2n * en ** (u - 1) * (1 + u * le) + C1n * u * en ** (u - 1)
        H[2][nzm] = C2n * ((u - 1) * en ** (u - 2) * (1 + u * le) + en ** (u - 2)) + C1n * u * (u - 1) * en ** (u - 2)
    else:
        sv = math.sqrt(-v)
        le = np.log(en)
        tc, ts = np.cos(sv * le), np.sin(sv * le)
        p1, p2 = u * tc - sv * ts, u * ts + sv * tc
        H[0][nzm] = en ** u * (C1n * tc + C2n * ts)
        H[1][nzm] = en ** (u - 1) * (C1n * p1 + C2n * p2)
        H[2][nzm] = en ** (u - 2) * ((u - 1) * (C1n * p1 + C2n * p2) + sv * (C2n * p1 - C1n * p2))

    P = [np.zeros_like(e) for _ in range(3)]
    m_in = inside & nzm
    ein = e[m_in][:, None]
    cB = B / denB
    P[0][m_in] = np.sum(cB * ein ** (2 * i), axis=-1)
    P[1][m_in] = np.sum(cB * (2 * i) * ein ** (2 * i - 1), axis=-1)
    P[2][m_in] = np.sum(cB * (2 * i) * (2 * i - 1) * ein ** (2 * i - 2), axis=-1)
    m_out = e > 1.0
    eout = e[m_out][:, None]
    cA = A / denA
    P[0][m_out] = np.sum(cA / eout ** (2 * i), axis=-1)
    P[1][m_out] = np.sum(cA * (-2 * i) / eout ** (2 * i + 1), axis=-1)
    P[2][m_out] = np.sum(cA * (2 * i) * (2 * i + 1) / eout ** (2 * i + 2), axis=-1)
    return [torch.from_numpy(H[d] + P[d]) for d in range(3)]


class WangGovindCarter99:
    """functionals.py:787-985.  ``forward(box, den)`` = vW + TF + Taylor-expanded density-dependent
    non-local term; kernel cached while eta (i.e. box, shape and round(N_elec)) is unchanged."""

    def __init__(self, alpha=(5 + math.sqrt(5)) / 6, beta=(5 - math.sqrt(5)) / 6, gamma=2.7, kappa=1.0):
        self.alpha, self.beta, self.gamma, self.kappa = alpha, beta, gamma, kappa
        self._eta = None
        self._w = None

    def forward(self, box, den):
        g = Grid(box, den.shape)
        n_elec = round((torch.mean(den) * g.vol).item())          # :952, rounded to an integer
        n_ref = self.kappa * n_elec / g.vol
        kF = (3.0 * PI * PI * n_ref) ** (1.0 / 3.0)
        eta = g.kabs / (2.0 * kF)
        if self._eta is None or not torch.equal(self._eta, eta):
            self._eta = eta
            self._w = wgc99_kernel(eta, self.alpha, self.beta, self.gamma)
        scale = 20.0 * n_ref ** (5.0 / 3.0 - self.alpha - self.beta)
        w0, w1, w2 = (scale * w for w in self._w)
        K1 = -eta * w1 / (6.0 * n_ref)
        K2 = (eta ** 2 * w2 + (7.0 - self.gamma) * eta * w1) / (36.0 * n_ref ** 2)
        K3 = (eta ** 2 * w2 + (1.0 + self.gamma) * eta * w1) / (36.0 * n_ref ** 2)
        theta = den - n_ref
        a = den.pow(self.beta)
        conv = g.inv(w0 * g.fwd(a)) + theta * g.inv(K1 * g.fwd(a)) + g.inv(K1 * g.fwd(a * theta)) \
            + theta ** 2 / 2 * g.inv(K2 * g.fwd(a)) + g.inv(K2 * g.fwd(a * theta ** 2 / 2)) \
            + theta * g.inv(K3 * g.fwd(a * theta))
        t_nl = C_TF * g.integral(den.pow(self.alpha) * conv)
        return Weizsaecker(box, den) + ThomasFermi(box, den) + t_nl

    __call__ = forward


# ----------------------------------------------------------------------------------------------
#  cubic Hermite tools and the field-dependent convolution (Huang-Carter family)
# ----------------------------------------------------------------------------------------------
def _hermite_basis(t):
    t2, t3 = t * t, t * t * t
    return 1 - 3 * t2 + 2 * t3, t - 2 * t2 + t3, 3 * t2 - 2 * t3, t3 - t2


def interpolate(x, y, xs):
    """functional_tools.py:292-334 : 1-D cubic Hermite, slopes = mean of adjacent secants
    (one-sided at the ends), interval by searchsorted(x[1:], xs)."""
    sec = (y[1:] - y[:-1]) / (x[1:] - x[:-1])
    m = torch.cat([sec[:1], 0.5 * (sec[1:] + sec[:-1]), sec[-1:]])
    idx = torch.searchsorted(x[1:], xs)
    dx = x[idx + 1] - x[idx]
    h00, h10, h01, h11 = _hermite_basis((xs - x[idx]) / dx)
    return h00 * y[idx] + h10 * m[idx] * dx + h01 * y[idx + 1] + h11 * m[idx + 1] * dx


def interpolate_kernel(nodes, f, xis):
    """functional_tools.py:337-378 : per-voxel Hermite along the last (node) axis of f."""
    dn = nodes[1:] - nodes[:-1]
    sec = (f[..., 1:] - f[..., :-1]) / dn
    m = torch.cat([sec[..., :1], 0.5 * (sec[..., 1:] + sec[..., :-1]), sec[..., -1:]], dim=-1)
    idx = torch.searchsorted(nodes[1:], xis)
    dx = nodes[idx + 1] - nodes[idx]
    h00, h10, h01, h11 = _hermite_basis((xis - nodes[idx]) / dx)

    def pick(arr, ii):
        return torch.gather(arr, 3, ii.unsqueeze(3))[..., 0]
    return h00 * pick(f, idx) + h10 * pick(m, idx) * dx + h01 * pick(f, idx + 1) + h11 * pick(m, idx + 1) * dx


def xi_nodes(xi_min, xi_max, kappa, mode):
    """functional_tools.py:406-417 : node list covering [min, max] plus 3 nodes of padding."""
    if mode == 'arithmetic':
        lower = (np.floor(xi_min / kappa) - 3) * kappa
        upper = (np.ceil(xi_max / kappa) + 3) * kappa
        nodes = torch.arange(lower, upper, kappa, dtype=DT)
        nodes[nodes == 0] = xi_min
        return nodes
    if mode == 'geometric':
        assert kappa > 1
        lower = kappa ** (-(np.ceil(-np.log(xi_min) / np.log(kappa)) + 3))
        count = np.ceil(np.log((xi_max + 1) / lower) / np.log(kappa)) + 3
        return lower * kappa ** torch.arange(count, dtype=DT)
    raise ValueError("Parameter 'mode' can only be 'arithmetic' or 'geometric'")


def field_dependent_convolution(k, f_tilde, g, xis, kappa, mode='arithmetic'):
    """functional_tools.py:381-423"""
    nodes = xi_nodes(xis.min().item(), xis.max().item(), kappa, mode)
    g_ft = torch.fft.rfftn(g).unsqueeze(3)
    conv = torch.fft.irfftn(f_tilde(k, nodes) * g_ft, s=g.shape, dim=(0, 1, 2))
    return interpolate_kernel(nodes, conv, xis)


def hc_kernel_table(beta, eta_max=50.0, n_eta=10000, rtol=1e-12, atol=1e-14):
    """functionals.py:1204-1230 : omega(eta) on linspace(0, eta_max, n_eta), from
    w' = -[(5/3)(1/Ginv - 3 eta^2 - 1) - (5 - 3 beta) beta w] / (beta eta), integrated from eta_max
    downwards with w(eta_max) = -(8/3)/((5 - 3 beta) beta); omega(0) := 0.
    The reference uses xitorch.solve_ivp defaults (not installed / un-pinned): here scipy's DOP853 at
    tight tolerance is the stand-in.  PARITY UNPINNED at this boundary."""
    from scipy.integrate import solve_ivp

    def ginv(e):
        if e == 0:
            return 1.0
        if e == 1:
            return 0.5
        return 0.5 + (1 - e * e) / (4 * e) * math.log(abs((1 + e) / (1 - e)))

    def rhs(e, w):
        aux = (5.0 / 3.0) * (1.0 / ginv(e) - 3 * e * e - 1) - (5 - 3 * beta) * beta * w[0]
        return [-aux / beta / e]
    etas = np.linspace(0.0, eta_max, n_eta)
    w_inf = -(8.0 / 3.0) / ((5 - 3 * beta) * beta)
    sol = solve_ivp(rhs, (etas[-1], etas[1]), [w_inf], t_eval=etas[1:][::-1], method='DOP853', rtol=rtol, atol=atol)
    w = np.concatenate([[0.0], sol.y[0][::-1]])
    return torch.from_numpy(np.stack([etas, w]))


class _HuangCarterBase:
    mode = 'geometric'

    def _xi(self, g, den):
        raise NotImplementedError

    def forward(self, box, den):
        g = Grid(box, den.shape)
        xis = self._xi(g, den)
        eta_1d, w_1d = self.kernel

        def w_tilde(q, nodes):
            eta = q.unsqueeze(3) / nodes
            return interpolate(eta_1d, w_1d, torch.minimum(eta, eta_1d[-1]))
        K = field_dependent_convolution(g.kabs, w_tilde, den.pow(self.beta), xis, self.kappa, self.mode)
        c_hc = C_TF * 8.0 * (3.0 * PI * PI)
        t_nl = c_hc * g.integral(den.pow(8.0 / 3.0 - self.beta) * K / xis.pow(3))
        return Weizsaecker(box, den) + ThomasFermi(box, den) + t_nl

    __call__ = forward


class HuangCarter(_HuangCarterBase):
    """functionals.py:1176-1269 : xi = 2 kF(n) (1 + lambda |grad n|^2 / (n^{8/3} + 1e-30))."""

    def __init__(self, lamb, beta, kappa, kernel=None):
        self.lamb, self.beta, self.kappa = lamb, beta, kappa
        self.kernel = hc_kernel_table(beta) if kernel is None else kernel

    def _xi(self, g, den):
        gx, gy, gz = g.grad(den)
        s2 = (gx * gx + gy * gy + gz * gz) / (den.pow(8.0 / 3.0) + 1e-30)
        return 2.0 * (3.0 * PI * PI * den).pow(1.0 / 3.0) * (1.0 + self.lamb * s2)


class RevisedHuangCarter(_HuangCarterBase):
    """functionals.py:1272-1365 : xi = 2 kF(n) (1 + a s^2 / (1 + b s^2)), s the reduced gradient."""

    def __init__(self, a, b, beta, kappa, kernel=None):
        self.a, self.b, self.beta, self.kappa = a, b, beta, kappa
        self.kernel = hc_kernel_table(beta) if kernel is None else kernel

    def _xi(self, g, den):
        gx, gy, gz = g.grad(den)
        s2 = 0.25 * (3.0 * PI * PI) ** (-2.0 / 3.0) * (gx * gx + gy * gy + gz * gz) / den.pow(8.0 / 3.0)
        return 2.0 * (3.0 * PI * PI * den).pow(1.0 / 3.0) * (1.0 + self.a * s2 / (1.0 + self.b * s2))


# ----------------------------------------------------------------------------------------------
#  exchange-correlation
# ----------------------------------------------------------------------------------------------
def lda_exchange(box, den):
    """functionals.py:1510-1512"""
    return -0.75 * (3.0 / PI) ** (1.0 / 3.0) * torch.mean(den.pow(4.0 / 3.0)) * torch.abs(torch.linalg.det(box))


def perdew_zunger_correlation(box, den):
    """functionals.py:1515-1521"""
    rs = (3.0 / 4.0 / PI / den).pow(1.0 / 3.0)
    high = 0.0311 * torch.log(rs) - 0.048 + 0.002 * rs * torch.log(rs) - 0.0116 * rs
    low = -0.1423 / (1.0 + 1.0529 * torch.sqrt(rs) + 0.3334 * rs)
    return torch.mean(torch.where(rs < 1, high, low) * den) * torch.abs(torch.linalg.det(box))


def PerdewZunger(box, den):
    """functionals.py:1540-1554"""
    return lda_exchange(box, den) + perdew_zunger_correlation(box, den)


def _pw92_eps(rs):
    A1, a1 = 0.0310907, 0.2137
    b1, b2, b3, b4 = 7.5957, 3.5876, 1.6382, 0.49294
    return -2 * A1 * (1 + a1 * rs) * torch.log(1 + 1 / (2 * A1 * (b1 * rs.pow(0.5) + b2 * rs + b3 * rs.pow(1.5) + b4 * rs.pow(2))))


def pbe_exchange(box, den):
    """functionals.py:1597-1603"""
    g = Grid(box, den.shape)
    gx, gy, gz = g.grad(den)
    s2 = 0.25 * (3.0 * PI * PI) ** (-2.0 / 3.0) * (gx * gx + gy * gy + gz * gz) / den.pow(8.0 / 3.0)
    kappa, mu = 0.804, 0.066725 * PI * PI / 3.0
    Fx = 1 + kappa - kappa / (1 + mu / kappa * s2)
    ex = -0.75 * (3.0 / PI) ** (1.0 / 3.0) * den.pow(4.0 / 3.0)
    return g.integral(Fx * ex)


def pbe_correlation(box, den):
    """functionals.py:1606-1618"""
    g = Grid(box, den.shape)
    rs = (3.0 / 4.0 / PI / den).pow(1.0 / 3.0)
    eps = _pw92_eps(rs)
    beta, gamma = 0.066725, (1 - math.log(2.0)) / PI / PI
    A = beta / gamma / (torch.exp(-eps / gamma) - 1 + 1e-30)
    gx, gy, gz = g.grad(den)
    t2 = (1.0 / 16.0) * (PI / 3.0) ** (1.0 / 3.0) * (gx * gx + gy * gy + gz * gz) / (den.pow(7.0 / 3.0) + 1e-30)
    At2 = A * t2
    H = gamma * torch.log(1 + beta / gamma * t2 * ((1 + At2) / (1 + At2 + At2 ** 2)))
    return g.integral((eps + H) * den)


def PerdewBurkeErnzerhof(box, den):
    """functionals.py:1621-1635"""
    return pbe_exchange(box, den) + pbe_correlation(box, den)


# ----------------------------------------------------------------------------------------------
#  potentials by autograd, chi-projection, optimizers, density optimisation loop
# ----------------------------------------------------------------------------------------------
def energy_and_potential(box, den, functional):
    """functional_tools.py:9-31 : dE/dn by autograd, divided by dV."""
    d = den.detach().clone().requires_grad_(True)
    E = functional(box, d)
    (grad,) = torch.autograd.grad(E, d)
    dV = torch.abs(torch.linalg.det(box)) / d.numel()
    return E.detach().reshape(()), grad / dV


def total_energy(box, den, terms, v_ext=None):
    """system.py:759-772 with for_den_opt=True (IonIon skipped).  ``terms`` is a list of callables;
    a term whose ``__name__`` is 'IonElectron' receives v_ext."""
    E = torch.zeros((), dtype=DT)
    for f in terms:
        name = getattr(f, '__name__', '')
        if name == 'IonElectron':
            E = E + f(box, den, v_ext)
        elif name == 'IonIon':
            continue
        else:
            E = E + f(box, den).reshape(())
    return E


def chi_gradient(box, chi, n_elec, terms, v_ext=None):
    """system.py:830-838 : n = N chi^2 / int chi^2;  returns E, dE/dchi_ijk (autograd partials), n."""
    c = chi.detach().clone().requires_grad_(True)
    vol = torch.abs(torch.linalg.det(box))
    n_tilde = torch.mean(c * c) * vol
    den = (n_elec / n_tilde) * c * c
    E = total_energy(box, den, terms, v_ext)
    (gr,) = torch.autograd.grad(E, c)
    return E.detach(), gr, den.detach()


def chi_gradient_from_potential(box, chi, n_elec, dEdn):
    """system.py:842-854 : the projection written out.  dEdn = delta E / delta n at n(chi)."""
    vol = torch.abs(torch.linalg.det(box))
    n_tilde = torch.mean(chi * chi) * vol
    den = (n_elec / n_tilde) * chi * chi
    mu = torch.mean(dEdn * den) * vol / n_elec
    return (n_elec / n_tilde) * 2 * chi * (dEdn - mu) * (vol / chi.numel())


class LbfgsState:
    """Fixed-step L-BFGS as driven by optimize_density (lbfgsnew.py:512-769 with
    line_search_fn=False, batch_mode=False).  ``step(closure)`` where closure(x) -> (loss, grad)."""

    def __init__(self, x, lr=0.1, max_iter=6, history=8, tol_grad=1e-5, tol_change=1e-9):
        self.x = x
        self.lr, self.max_iter, self.m = lr, max_iter, history
        self.max_eval = max_iter * 5 // 4
        self.tol_grad, self.tol_change = tol_grad, tol_change
        self.total_iter = 0
        self.d = self.t = self.prev_g = self.prev_loss = None
        self.Y, self.S = [], []
        self.H = 1.0
        self.grad = None
        self.closures = 0

    def step(self, closure):
        loss, g = closure(self.x)
        self.closures += 1
        self.grad = g
        loss = float(loss)
        evals = 1
        g1 = float(g.abs().sum())
        if g1 <= self.tol_grad:
            return loss
        it = 0
        gn = float(g.norm())
        d, t = self.d, self.t
        while it < self.max_iter and not math.isnan(gn):
            it += 1
            self.total_iter += 1
            if self.total_iter == 1:
                d = -g
                self.Y, self.S, self.H = [], [], 1.0
            else:
                y = g - self.prev_g
                s = d * t
                ys = float(y.dot(s))
                sn = float(s.norm())
                if ys > 1e-10 * sn * sn:
                    if len(self.Y) == self.m:
                        self.Y.pop(0)
                        self.S.pop(0)
                    self.Y.append(y)
                    self.S.append(s)
                    self.H = ys / float(y.dot(y))
                k = len(self.Y)
                rho = [1.0 / float(self.Y[i].dot(self.S[i])) for i in range(k)]
                al = [0.0] * k
                q = -g
                for i in range(k - 1, -1, -1):
                    al[i] = float(self.S[i].dot(q)) * rho[i]
                    q = q - al[i] * self.Y[i]
                d = q * self.H
                for i in range(k):
                    be = float(self.Y[i].dot(d)) * rho[i]
                    d = d + (al[i] - be) * self.S[i]
            self.prev_g = g.clone()
            self.prev_loss = loss
            t = min(1.0, 1.0 / g1) * self.lr if self.total_iter == 1 else self.lr
            gtd = float(g.dot(d))
            self.x.add_(d, alpha=t)
            if it != self.max_iter:
                loss, g = closure(self.x)
                self.closures += 1
                self.grad = g
                loss = float(loss)
                g1 = float(g.abs().sum())
                if math.isnan(g1):
                    break
                evals += 1
            if it == self.max_iter or evals >= self.max_eval or g1 <= self.tol_grad:
                break
            if gtd > -self.tol_change or float((d * t).abs().sum()) <= self.tol_change:
                break
            if abs(loss - self.prev_loss) < self.tol_change:
                break
        self.d, self.t = d, t
        return loss


class TpgdState:
    """two_point_gradient_descent.py:25-65 : Barzilai-Borwein step <dx,dx>/<dx,dg>."""

    def __init__(self, x, lr=0.1):
        self.x, self.lr, self.it = x, lr, 0
        self.x_prev = self.g_prev = None
        self.grad = None
        self.closures = 0

    def step(self, closure):
        loss, g = closure(self.x)
        self.closures += 1
        self.grad = g
        num = den = 0.0
        if self.it != 0:
            dx, dg = self.x - self.x_prev, g - self.g_prev
            num, den = float((dx * dx).sum()), float((dx * dg).sum())
        self.x_prev, self.g_prev = self.x.clone(), g.clone()
        if self.it == 0 or den == 0:
            a = self.lr
        else:
            a = num / den
            if a <= 0:
                a = self.lr
        self.x.add_(g, alpha=-a)
        self.it += 1
        return float(loss)


def optimize_density(box, den0, n_elec, terms, v_ext=None, ntol=1e-7, n_conv_cond_count=3, n_method='LBFGS',
                     n_step_size=0.1, n_maxiter=1000, potentials=None):
    """system.py:774-908 with conv_target='dE'.  Returns dict(den, energy [Ha], iterations, closures, trace)."""
    shape = den0.shape
    chi = torch.sqrt(den0).reshape(-1).clone()
    state = {}

    def closure(x):
        c = x.reshape(shape)
        if potentials is None:
            E, gr, den = chi_gradient(box, c, n_elec, terms, v_ext)
        else:
            vol = torch.abs(torch.linalg.det(box))
            den = (n_elec / (torch.mean(c * c) * vol)) * c * c
            E = total_energy(box, den, terms, v_ext)
            gr = chi_gradient_from_potential(box, c, n_elec, potentials(box, den))
        state['den'] = den
        return E, gr.reshape(-1)

    if n_method == 'LBFGS':
        opt = LbfgsState(chi, lr=n_step_size, history=8, max_iter=6)
    elif n_method == 'TPGD':
        opt = TpgdState(chi, lr=n_step_size)
    else:
        raise ValueError("Only 'LBFGS' or 'TPGD' recognized for 'n_method' argument")
    E_prev = float(total_energy(box, den0, terms, v_ext)) * EV_PER_HA
    conv, trace, iters = 0, [], 0
    for it in range(1, round(n_maxiter) + 1):
        iters = it
        opt.step(closure)
        E = float(total_energy(box, state['den'], terms, v_ext)) * EV_PER_HA
        dE, E_prev = E - E_prev, E
        trace.append(E)
        if it > 5:
            conv = conv + 1 if abs(dE) < ntol else 0
        if conv == n_conv_cond_count:
            break
    den = state['den']
    return dict(den=den, energy=float(total_energy(box, den, terms, v_ext)), iterations=iters,
                closures=opt.closures, trace=trace)


# ----------------------------------------------------------------------------------------------
#  ionic rows: local pseudopotential, structure factor, forces, stress
# ----------------------------------------------------------------------------------------------
_BOHR = 0.529177208607388                   # ion_utils.py:8-10 (the reference's own, older, constants)
_HARTREE_TO_EV = 27.2113834279111
_POT_CONV = 1.0 / (_BOHR ** 3 * _HARTREE_TO_EV)


def read_recpot(path):
    """ion_utils.py:62-75 : (k grid [1/bohr], v(k) [Ha bohr^3], ion charge) of a CASTEP .recpot file."""
    vals = []
    with open(path, 'r') as fh:
        for line in fh:
            if 'END COMMENT' in line:
                break
        fh.readline()
        k_max = float(fh.readline()) * _BOHR
        for line in fh:
            cols = line.split()
            if len(cols) == 3:
                vals += cols
    pot = np.asarray(vals, dtype=np.float64) * _POT_CONV
    ks, dk = np.linspace(0, k_max, pot.size, retstep=True)
    z = round((pot[1] - pot[0]) * dk * dk / (-4 * PI))
    return ks, pot, z


def interpolate_recpot(path, kabs):
    """ion_utils.py:49-81 : Coulomb tail added back, cubic Hermite on min(k, k_max), tail removed for k != 0."""
    ks, pot, z = read_recpot(path)
    pot = pot.copy()
    pot[1:] += 4 * PI * z / (ks[1:] * ks[1:])
    x, y = torch.as_tensor(ks, dtype=DT), torch.as_tensor(pot, dtype=DT)
    val = interpolate(x, y, torch.minimum(kabs, x[-1]))
    nz = kabs != 0
    out = val.clone()
    out[nz] = val[nz] - 4 * PI * z / kabs[nz].pow(2)
    return out


def ionic_potential(box, shape, species):
    """system.py:183-205 + ion_utils.py:88-137 (exact structure factor).  ``species`` = [(recpot path,
    (n, 3) CARTESIAN coordinates), ...].  Differentiable w.r.t. the coordinates and the box."""
    g = Grid(box, shape)
    k = g.kabs                                                   # system.py:185-186 (masked sqrt: k = 0 stays 0)
    v = torch.zeros(g.shape, dtype=DT)
    for path, cart in species:
        phase = (g.kvec[0].unsqueeze(-1) * cart[:, 0] + g.kvec[1].unsqueeze(-1) * cart[:, 1]
                 + g.kvec[2].unsqueeze(-1) * cart[:, 2])
        S = torch.complex(torch.cos(phase), -torch.sin(phase)).sum(-1)
        v = v + torch.fft.irfftn(S * interpolate_recpot(path, k), g.shape, norm='forward') / g.vol
    return v


def ion_electron_forces(box, den, species):
    """IonElectron part of system.py:913-925 : -d/dR of mean(den * v_ext[R]) * vol, by autograd."""
    carts = [c.detach().clone().requires_grad_(True) for _, c in species]
    v = ionic_potential(box, den.shape, [(p, c) for (p, _), c in zip(species, carts)])
    U = IonElectron(box, den, v)
    grads = torch.autograd.grad(U, carts)
    return -torch.cat(grads)


def stress(box, den, functional):
    """functional_tools.py:73-100 : (1/vol) dF/dh h^T by autograd, with the density rescaled so that the
    electron number is conserved under the strain."""
    b = box.detach().clone().requires_grad_(True)
    vol = torch.abs(torch.linalg.det(b))
    d = den * vol.detach() / vol
    (g,) = torch.autograd.grad(functional(b, d).reshape(()), b)
    return (g.T @ b.detach()) / vol.detach()


def ion_electron_stress(box, den, species_frac):
    """IonElectron part of system.py:927-935 : ions at fixed FRACTIONAL coordinates follow the strain."""
    b = box.detach().clone().requires_grad_(True)
    vol = torch.abs(torch.linalg.det(b))
    d = den * vol.detach() / vol
    v = ionic_potential(b, den.shape, [(p, f @ b) for p, f in species_frac])
    (g,) = torch.autograd.grad(IonElectron(b, d, v), b)
    st = (g.T @ b.detach()) / vol.detach()
    return 0.5 * (st + st.T)


# ----------------------------------------------------------------------------------------------
#  deterministic synthetic inputs
# ----------------------------------------------------------------------------------------------
def synth_smooth(n, side):
    """Smooth Al-like density on a cubic cell of side*a (a = 4.05 A), 4*side^3 atoms, 3 e/atom."""
    a = 4.05 / 0.529177210903
    L = side * a
    box = L * torch.eye(3, dtype=DT)
    x = torch.arange(n, dtype=DT) / n
    X, Y, Z = torch.meshgrid(x, x, x, indexing='ij')
    k = 2 * PI * side
    n0 = 12 * side ** 3 / L ** 3
    den = n0 * (1 + 0.3 * torch.cos(k * X) * torch.cos(k * Y) * torch.cos(k * Z)
                + 0.05 * torch.cos(2 * k * X) * torch.cos(2 * k * Y))
    return box, den


def synth_rough(shape, seed=0, L=7.6):
    """Rough density on a skewed cell: exercises Nyquist handling on even grids and the masks."""
    gen = torch.Generator().manual_seed(seed)
    box = L * torch.eye(3, dtype=DT) + 0.2 * torch.rand(3, 3, dtype=DT, generator=gen)
    den = 0.03 * (1 + 0.5 * torch.rand(*shape, dtype=DT, generator=gen))
    return box, den
