"""CPU model of the *analytic* formulation the CUDA kernels implement  --  TEST HELPER ONLY.

The product computes potentials analytically (no autograd) with

  * the minimum number of FFTs (vW 2, Hartree 2, WT 4, WGC98 6, WGC99 14, PBE 8),
  * reciprocal-space multipliers that are Hermitian-symmetrised on the self-conjugate planes
    (j2 == 0 and, for even n2, j2 == n2/2), which reproduces what the reference's
    ``torch.fft.irfftn`` does with the non-Hermitian "Nyquist made positive" multipliers of
    functional_tools.py:152-155 on even grids.

``tests/test_analytic_model.py`` checks every formula here against the autograd oracle on CPU, so a
GPU parity failure can be split into "formula wrong" vs "kernel wrong".  Nothing in the product
imports this file.
"""
import math

import numpy as np
import torch

PI = math.pi
C_TF = 0.3 * (3.0 * PI * PI) ** (2.0 / 3.0)


class KGrid:
    def __init__(self, box, shape):
        self.shape = tuple(shape)
        n0, n1, n2 = self.shape
        self.vol = float(torch.abs(torch.linalg.det(box)))
        self.N = n0 * n1 * n2
        self.dV = self.vol / self.N
        b = (2 * PI * torch.linalg.inv(box.T)).numpy()

        def freq(n):
            f = np.fft.fftfreq(n) * n
            f[n // 2] = abs(f[n // 2])
            return f
        f0, f1, f2 = freq(n0), freq(n1), np.fft.rfftfreq(n2) * n2
        # partner frequencies: -f except at a Nyquist index (which is its own partner with the same +n/2)
        p0, p1 = -f0.copy(), -f1.copy()
        if n0 % 2 == 0:
            p0[n0 // 2] = f0[n0 // 2]
        if n1 % 2 == 0:
            p1[n1 // 2] = f1[n1 // 2]
        A, B, C = np.meshgrid(f0, f1, f2, indexing='ij')
        PA, PB, _ = np.meshgrid(p0, p1, f2, indexing='ij')
        self.k = [torch.from_numpy(A * b[0, c] + B * b[1, c] + C * b[2, c]) for c in range(3)]
        self.kp = [torch.from_numpy(PA * b[0, c] + PB * b[1, c] + C * b[2, c]) for c in range(3)]
        sc = np.zeros(n2 // 2 + 1, dtype=bool)
        sc[0] = True
        if n2 % 2 == 0:
            sc[-1] = True
        self.selfconj = torch.from_numpy(np.broadcast_to(sc, A.shape).copy())

    def sym(self, fn):
        """M_eff = M(k(p)) off the self-conjugate planes, (M(k(p)) + conj M(k(pbar)))/2 on them."""
        M = fn(*self.k).to(torch.complex128)
        Mp = fn(*self.kp).to(torch.complex128)
        return torch.where(self.selfconj, 0.5 * (M + Mp.conj()), M)

    def fwd(self, f):
        return torch.fft.rfftn(f)

    def inv(self, F):
        return torch.fft.irfftn(F, self.shape)

    def integ(self, f):
        return float(f.sum()) * self.dV


def _kabs(kx, ky, kz):
    return torch.sqrt(kx * kx + ky * ky + kz * kz)


def hartree(box, den):
    g = KGrid(box, den.shape)

    def coul(kx, ky, kz):
        k2 = kx * kx + ky * ky + kz * kz
        return torch.where(k2 != 0, 4 * PI / torch.where(k2 != 0, k2, torch.ones_like(k2)), torch.zeros_like(k2))
    phi = g.inv(g.sym(coul) * g.fwd(den))
    return 0.5 * g.integ(den * phi), phi


def thomas_fermi(box, den):
    g = KGrid(box, den.shape)
    c = torch.pow(den, 1.0 / 3.0)
    return g.integ(C_TF * den * c * c), (5.0 / 3.0) * C_TF * c * c


def weizsaecker(box, den):
    g = KGrid(box, den.shape)
    chi = torch.sqrt(den)
    lap = g.inv(g.sym(lambda kx, ky, kz: -(kx * kx + ky * ky + kz * kz)) * g.fwd(chi))
    v = torch.where(den != 0, -0.5 * lap / torch.where(den != 0, chi, torch.ones_like(chi)), torch.zeros_like(chi))
    return g.integ(-0.5 * chi * lap), v


def lindhard_kernel(eta):
    ginv = torch.ones_like(eta)
    reg = (eta != 0) & (eta != 1)
    e = eta[reg]
    ginv[reg] = 0.5 + (1 - e * e) / (4 * e) * torch.log(torch.abs((1 + e) / (1 - e)))
    ginv[eta == 1] = 0.5
    return 1.0 / ginv - 3 * eta * eta - 1


def wt_nonlocal(box, den, alpha, beta):
    """T_NL and its potential only (2 FFTs if alpha == beta else 4)."""
    g = KGrid(box, den.shape)
    n0 = float(den.mean())                       # = N_elec / vol, detached
    kF = (3 * PI * PI * n0) ** (1.0 / 3.0)
    pref = 5.0 / (9 * alpha * beta * n0 ** (alpha + beta - 5.0 / 3.0))
    K = g.sym(lambda kx, ky, kz: pref * lindhard_kernel(_kabs(kx, ky, kz) / (2 * kF)))
    pa, pb = den.pow(alpha), den.pow(beta)
    conv_b = g.inv(K * g.fwd(pb - n0 ** beta))
    E = C_TF * g.integ((pa - n0 ** alpha) * conv_b)
    conv_a = conv_b if alpha == beta else g.inv(K * g.fwd(pa - n0 ** alpha))
    v = C_TF * (alpha * pa / den * conv_b + beta * pb / den * conv_a)
    return E, v


def wt_family(box, den, alpha, beta):
    parts = [thomas_fermi(box, den), weizsaecker(box, den), wt_nonlocal(box, den, alpha, beta)]
    return sum(p[0] for p in parts), sum(p[1] for p in parts)


def wgc99(box, den, w_of_eta, alpha, beta, gamma, kappa):
    """14-FFT form (12 for the non-local part + 2 for vW).  ``w_of_eta(eta) -> (w, w', w'')``."""
    g = KGrid(box, den.shape)
    n_ref = kappa * round(float(den.mean()) * g.vol) / g.vol
    kF = (3 * PI * PI * n_ref) ** (1.0 / 3.0)
    scale = 20.0 * n_ref ** (5.0 / 3.0 - alpha - beta)

    def kernels(which):
        def fn(kx, ky, kz):
            eta = _kabs(kx, ky, kz) / (2 * kF)
            w0, w1, w2 = (scale * w for w in w_of_eta(eta))
            if which == 0:
                return w0
            if which == 1:
                return -eta * w1 / (6 * n_ref)
            if which == 2:
                return (eta ** 2 * w2 + (7 - gamma) * eta * w1) / (36 * n_ref ** 2)
            return (eta ** 2 * w2 + (1 + gamma) * eta * w1) / (36 * n_ref ** 2)
        return g.sym(fn)
    W0, K1, K2, K3 = (kernels(i) for i in range(4))
    th = den - n_ref
    a, P = den.pow(beta), den.pow(alpha)
    A, B, C = g.fwd(a), g.fwd(a * th), g.fwd(0.5 * a * th * th)
    u1, u2, u3 = g.inv(W0 * A + K1 * B + K2 * C), g.inv(K1 * A + K3 * B), g.inv(K2 * A)
    conv = u1 + th * u2 + 0.5 * th * th * u3
    E_nl = C_TF * g.integ(P * conv)
    Pf, Pt, Pt2 = g.fwd(P), g.fwd(P * th), g.fwd(0.5 * P * th * th)
    g1, g2, g3 = g.inv(W0 * Pf + K1 * Pt + K2 * Pt2), g.inv(K1 * Pf + K3 * Pt), g.inv(K2 * Pf)
    dadn, dPdn = beta * a / den, alpha * P / den
    v_nl = C_TF * (dPdn * conv + P * (u2 + th * u3)
                   + dadn * g1 + (dadn * th + a) * g2 + (0.5 * dadn * th * th + a * th) * g3)
    e_tf, v_tf = thomas_fermi(box, den)
    e_vw, v_vw = weizsaecker(box, den)
    return E_nl + e_tf + e_vw, v_nl + v_tf + v_vw


def perdew_zunger(box, den):
    g = KGrid(box, den.shape)
    cx = -0.75 * (3 / PI) ** (1.0 / 3.0)
    c = den.pow(1.0 / 3.0)
    ex, vx = cx * den * c, (4.0 / 3.0) * cx * c
    rs = (3.0 / (4 * PI * den)).pow(1.0 / 3.0)
    A, B, C, D = 0.0311, -0.048, 0.002, -0.0116
    ga, b1, b2 = -0.1423, 1.0529, 0.3334
    lr, sr = torch.log(rs), torch.sqrt(rs)
    den_lo = 1 + b1 * sr + b2 * rs
    eps = torch.where(rs < 1, A * lr + B + C * rs * lr + D * rs, ga / den_lo)
    vc = torch.where(rs < 1, lr * (A + 2.0 / 3.0 * C * rs) + (B - A / 3) + rs / 3 * (2 * D - C),
                     ga * (1 + 7.0 / 6.0 * b1 * sr + 4.0 / 3.0 * b2 * rs) / den_lo ** 2)
    return g.integ(ex + eps * den), vx + vc


def pbe_local(den, sig):
    """f, df/dn, df/dsigma of the PBE exchange-correlation energy density (sigma = |grad n|^2)."""
    # exchange
    cx = -0.75 * (3 / PI) ** (1.0 / 3.0)
    cs = 0.25 * (3 * PI * PI) ** (-2.0 / 3.0)
    kap, mu = 0.804, 0.066725 * PI * PI / 3
    c13 = den.pow(1.0 / 3.0)
    ex_unif = cx * den * c13
    r83 = den * den * c13 * c13
    s2 = cs * sig / r83
    q = 1 + mu / kap * s2
    Fx, dF = 1 + kap - kap / q, mu / (q * q)
    f = Fx * ex_unif
    f_rho = Fx * (4.0 / 3.0) * cx * c13 + ex_unif * dF * (-8.0 / 3.0) * s2 / den
    f_sig = ex_unif * dF * cs / r83
    # correlation
    A1, a1 = 0.0310907, 0.2137
    b1, b2, b3, b4 = 7.5957, 3.5876, 1.6382, 0.49294
    rs = (3.0 / (4 * PI * den)).pow(1.0 / 3.0)
    sr = torch.sqrt(rs)
    Q = 2 * A1 * (b1 * sr + b2 * rs + b3 * rs * sr + b4 * rs * rs)
    lg = torch.log(1 + 1 / Q)
    eps = -2 * A1 * (1 + a1 * rs) * lg
    dQ = A1 * (b1 / sr + 2 * b2 + 3 * b3 * sr + 4 * b4 * rs)
    deps = (-2 * A1 * a1 * lg + 2 * A1 * (1 + a1 * rs) * dQ / (Q * (Q + 1))) * (-rs / (3 * den))
    be, ga = 0.066725, (1 - math.log(2.0)) / PI / PI
    ee = torch.exp(-eps / ga)
    Aa = be / ga / (ee - 1 + 1e-30)
    dAa = Aa * Aa / be * ee * deps
    ct = (1.0 / 16.0) * (PI / 3) ** (1.0 / 3.0)
    r73 = den * den * c13 + 1e-30
    t2 = ct * sig / r73
    dt2_rho = -ct * sig * (7.0 / 3.0) * den * c13 / (r73 * r73)
    dt2_sig = ct / r73
    X = Aa * t2
    num, dnm = 1 + X, 1 + X + X * X
    Rr = num / dnm
    dR = -X * (2 + X) / (dnm * dnm)
    inner = 1 + be / ga * t2 * Rr
    H = ga * torch.log(inner)
    dH_rho = be / inner * (Rr * dt2_rho + t2 * dR * (Aa * dt2_rho + t2 * dAa))
    dH_sig = be / inner * (Rr + t2 * dR * Aa) * dt2_sig
    f = f + den * (eps + H)
    f_rho = f_rho + eps + H + den * (deps + dH_rho)
    f_sig = f_sig + den * dH_sig
    return f, f_rho, f_sig


def pbe(box, den):
    """8-FFT PBE: rfft(n); 3 c2r for grad n; local GGA kernel; 3 r2c of w_i = 2 f_sigma g_i; 1 c2r of i k . w."""
    g = KGrid(box, den.shape)
    ik = [g.sym(lambda kx, ky, kz, c=c: 1j * (kx, ky, kz)[c]) for c in range(3)]
    R = g.fwd(den)
    gr = [g.inv(m * R) for m in ik]
    sig = gr[0] ** 2 + gr[1] ** 2 + gr[2] ** 2
    f, f_rho, f_sig = pbe_local(den, sig)
    div = sum(m * g.fwd(2 * f_sig * gi) for m, gi in zip(ik, gr))
    return g.integ(f), f_rho - g.inv(div)


def chi_projection(box, chi, n_elec, v):
    """system.py:842-854"""
    vol = float(torch.abs(torch.linalg.det(box)))
    dV = vol / chi.numel()
    scale = n_elec / (float((chi * chi).sum()) * dV)
    den = scale * chi * chi
    mu = float((v * den).sum()) * dV / n_elec
    return scale * 2 * chi * (v - mu) * dV


# ----------------------------------------------------------------------------------------------
#  analytic stresses  sigma_ij = (1/vol) dE/d eps_ij  at fixed electron number (the density scales as 1/vol)
#  -- what functional_tools.py:73-100 obtains by autograd through box_vecs.  For even multipliers the
#  Hermitian symmetrisation drops out of the energy (the self-conjugate planes contain p and pbar), so the
#  reciprocal-space sums run over the plain "Nyquist made positive" k with half-spectrum weights w.
# ----------------------------------------------------------------------------------------------
def _weights(g):
    return torch.where(g.selfconj, torch.ones((), dtype=torch.double), 2.0 * torch.ones((), dtype=torch.double))


def _tensor_sum(g, scalar):
    """sum_k scalar(k) k_i k_j  as a symmetric 3 x 3 tensor."""
    out = torch.zeros(3, 3, dtype=torch.double)
    for i in range(3):
        for j in range(i, 3):
            out[i, j] = out[j, i] = float((scalar * g.k[i] * g.k[j]).sum())
    return out


def stress_local(box, den, E, v):
    """TF, LDA exchange, PZ correlation: delta_ij (E - int v n) / vol."""
    g = KGrid(box, den.shape)
    return (E - g.integ(v * den)) / g.vol * torch.eye(3, dtype=torch.double)


def stress_hartree(box, den):
    g = KGrid(box, den.shape)
    c = g.fwd(den) / g.N
    k2 = g.k[0] ** 2 + g.k[1] ** 2 + g.k[2] ** 2
    safe = torch.where(k2 != 0, k2, torch.ones_like(k2))
    aux = torch.where(k2 != 0, _weights(g) * 4 * PI * (c.real ** 2 + c.imag ** 2) / (safe * safe), torch.zeros_like(k2))
    E = hartree(box, den)[0]
    return _tensor_sum(g, aux) - E / g.vol * torch.eye(3, dtype=torch.double)


def stress_weizsaecker(box, den):
    g = KGrid(box, den.shape)
    c = g.fwd(torch.sqrt(den)) / g.N
    return -_tensor_sum(g, _weights(g) * (c.real ** 2 + c.imag ** 2))


def stress_wt_nonlocal(box, den, alpha, beta):
    """tests/tools_for_tests.py:258-307 (non_local_KEF_stress without its TF and vW parts)."""
    g = KGrid(box, den.shape)
    n0 = float(den.mean())
    kF = (3 * PI * PI * n0) ** (1.0 / 3.0)
    E = wt_nonlocal(box, den, alpha, beta)[0]
    pref = 0.5 * PI * PI / alpha / beta / n0 ** (alpha + beta - 2) / kF
    a, b = g.fwd(den.pow(alpha)) / g.N, g.fwd(den.pow(beta)) / g.N
    ab = (a * b.conj()).real * _weights(g)                          # filter * aux1 = (w / 2) * 2 Re(a conj b)
    k2 = g.k[0] ** 2 + g.k[1] ** 2 + g.k[2] ** 2
    nz = k2 != 0
    safe = torch.where(nz, k2, torch.ones_like(k2))
    eta = torch.sqrt(safe) / (2 * kF)
    lg = torch.log(torch.abs((1 + eta) / (1 - eta)))
    lind = 0.5 + (1 - eta * eta) / (4 * eta) * lg
    aux3 = eta / lind ** 2 * (0.5 / eta - 0.25 * (1 + 1 / (eta * eta)) * lg) + 6 * eta * eta
    s = torch.where(nz, ab * aux3 / safe, torch.zeros_like(k2))
    iso = float(torch.where(nz, ab * aux3, torch.zeros_like(k2)).sum()) / 3.0
    eye = torch.eye(3, dtype=torch.double)
    return pref * (_tensor_sum(g, s) - iso * eye) - 2.0 * E / 3.0 / g.vol * eye


def stress_pbe(box, den):
    """tests/tools_for_tests.py:367-472: delta_ij mean(f - n f_n - 2 sigma f_sigma) - 2 mean(f_sigma g_i g_j)."""
    g = KGrid(box, den.shape)
    ik = [g.sym(lambda kx, ky, kz, c=c: 1j * (kx, ky, kz)[c]) for c in range(3)]
    R = g.fwd(den)
    gr = [g.inv(m * R) for m in ik]
    f, f_rho, f_sig = pbe_local(den, gr[0] ** 2 + gr[1] ** 2 + gr[2] ** 2)
    out = float((f - den * f_rho - 2 * (gr[0] ** 2 + gr[1] ** 2 + gr[2] ** 2) * f_sig).mean()) * torch.eye(3, dtype=torch.double)
    for i in range(3):
        for j in range(3):
            out[i, j] -= 2 * float((f_sig * gr[i] * gr[j]).mean())
    return out


def recpot_value_and_slope(ks, smooth, z, k):
    """v(k) = H(min(k, k_max)) - 4 pi z / k^2 and dv/dk for the Hermite interpolant H of the tail-free table."""
    x, y = torch.as_tensor(ks), torch.as_tensor(smooth)
    sec = (y[1:] - y[:-1]) / (x[1:] - x[:-1])
    m = torch.cat([sec[:1], 0.5 * (sec[1:] + sec[:-1]), sec[-1:]])
    kc = torch.minimum(k, x[-1])
    idx = torch.searchsorted(x[1:], kc)
    dx = x[idx + 1] - x[idx]
    t = (kc - x[idx]) / dx
    t2, t3 = t * t, t * t * t
    val = (1 - 3 * t2 + 2 * t3) * y[idx] + (t - 2 * t2 + t3) * m[idx] * dx + (3 * t2 - 2 * t3) * y[idx + 1] + (t3 - t2) * m[idx + 1] * dx
    slope = ((-6 * t + 6 * t2) * y[idx] + (1 - 4 * t + 3 * t2) * m[idx] * dx + (6 * t - 6 * t2) * y[idx + 1] + (3 * t2 - 2 * t) * m[idx + 1] * dx) / dx
    slope = torch.where(k > x[-1], torch.zeros_like(slope), slope)
    nz = k != 0
    ksafe = torch.where(nz, k, torch.ones_like(k))
    val = torch.where(nz, val - 4 * PI * z / ksafe ** 2, val)
    slope = torch.where(nz, slope + 8 * PI * z / ksafe ** 3, torch.zeros_like(slope))
    return val, slope


def stress_ion_electron(box, den, species):
    """species = [(ks, smooth table, z, (n, 3) fractional coordinates)].  At fixed fractional coordinates S(k) does not
    depend on the cell:  sigma_ij = -delta_ij E / vol - (1/vol) sum_k w v'(k) k_i k_j / k Re(S conj c)."""
    g = KGrid(box, den.shape)
    n0, n1, n2 = g.shape
    c = g.fwd(den) / g.N
    kabs = torch.sqrt(g.k[0] ** 2 + g.k[1] ** 2 + g.k[2] ** 2)
    f0 = np.fft.fftfreq(n0) * n0
    f0[n0 // 2] = abs(f0[n0 // 2])
    f1 = np.fft.fftfreq(n1) * n1
    f1[n1 // 2] = abs(f1[n1 // 2])
    f2 = np.fft.rfftfreq(n2) * n2
    A, B, C = (torch.from_numpy(a) for a in np.meshgrid(f0, f1, f2, indexing='ij'))
    E, scal = 0.0, torch.zeros_like(kabs)
    w = _weights(g)
    for ks, smooth, z, frac in species:
        ph = -2 * PI * (A.unsqueeze(-1) * frac[:, 0] + B.unsqueeze(-1) * frac[:, 1] + C.unsqueeze(-1) * frac[:, 2])
        S = torch.complex(torch.cos(ph), torch.sin(ph)).sum(-1)
        val, slope = recpot_value_and_slope(ks, smooth, z, kabs)
        re = w * (S * c.conj()).real
        E += float((val * re).sum())
        scal = scal + torch.where(kabs != 0, slope * re / torch.where(kabs != 0, kabs, torch.ones_like(kabs)), torch.zeros_like(kabs))
    return -E / g.vol * torch.eye(3, dtype=torch.double) - _tensor_sum(g, scal) / g.vol


def stress_huang_carter_nonlocal(box, den, f):
    """Analytic stress of the non-local Huang-Carter term for an oracle functional object ``f`` (HuangCarter or
    RevisedHuangCarter of oracle/ofdft_oracle.py) -- the formula a CUDA implementation has to evaluate; the xi-node
    list is a constant, as in the reference (functional_tools.py:408-416 reads min / max on the host):

      sigma_ab = (1/vol) [ delta_ab (E_NL - int v_NL n)                       volume element + density scaling (n ~ 1/vol)
                           - sum_r (dE/dg_a)(r) g_b(r)                        strain of the spectral gradient, g -> (1 - eps) g
                           - sum_j sum_k w (omega'(|k| / xi_j) / xi_j) (k_a k_b / |k|) Re[conj(W_j^) g^] / N ]
    with g = grad n, dE/dg_a = 2 E_xi xi_sigma g_a dV, W_j = dE/dconv_j (the node weights of the adjoint convolution)
    and g^ = rfftn(n^beta).  The partial derivatives are taken by autograd here; the kernels hold them explicitly
    (csrc/hc.cu: Ex, xi_s, W_j)."""
    from oracle import ofdft_oracle as orc
    g = KGrid(box, den.shape)
    og = orc.Grid(box, den.shape)
    eta_1d, w_1d = f.kernel
    n = den.clone().requires_grad_(True)
    gr = [x.detach().clone().requires_grad_(True) for x in og.grad(den)]
    sig_field = gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]
    if hasattr(f, 'lamb'):
        xis = 2 * (3 * PI * PI * n).pow(1 / 3) * (1 + f.lamb * sig_field / (n.pow(8 / 3) + 1e-30))
    else:
        s2 = 0.25 * (3 * PI * PI) ** (-2 / 3) * sig_field / n.pow(8 / 3)
        xis = 2 * (3 * PI * PI * n).pow(1 / 3) * (1 + f.a * s2 / (1 + f.b * s2))
    nodes = orc.xi_nodes(xis.min().item(), xis.max().item(), f.kappa, 'geometric')
    eta = og.kabs.unsqueeze(3) / nodes
    G = torch.fft.rfftn(den.pow(f.beta)).unsqueeze(3)
    omega = orc.interpolate(eta_1d, w_1d, torch.minimum(eta, eta_1d[-1]))
    conv = torch.fft.irfftn(omega * G, s=den.shape, dim=(0, 1, 2)).detach().clone().requires_grad_(True)
    K = orc.interpolate_kernel(nodes, conv, xis)
    E_nl = C_TF * 8 * (3 * PI * PI) * torch.mean(n.pow(8 / 3 - f.beta) * K / xis.pow(3)) * g.vol
    dE_dg = torch.autograd.grad(E_nl, gr, retain_graph=True)
    (W,) = torch.autograd.grad(E_nl, conv)
    _, v_nl = orc.energy_and_potential(box, den, lambda b, d: f.forward(b, d) - orc.ThomasFermi(b, d) - orc.Weizsaecker(b, d))
    out = (float(E_nl) - g.integ(v_nl * den)) / g.vol * torch.eye(3, dtype=torch.double)
    for a in range(3):
        for b in range(3):
            out[a, b] -= float((dE_dg[a] * gr[b].detach()).sum()) / g.vol
    # d omega / d eta of the Hermite table (zero beyond its end: the argument is clamped)
    sec = (w_1d[1:] - w_1d[:-1]) / (eta_1d[1:] - eta_1d[:-1])
    m = torch.cat([sec[:1], 0.5 * (sec[1:] + sec[:-1]), sec[-1:]])
    ec = torch.minimum(eta, eta_1d[-1])
    idx = torch.searchsorted(eta_1d[1:], ec)
    dx = eta_1d[idx + 1] - eta_1d[idx]
    t = (ec - eta_1d[idx]) / dx
    slope = ((-6 * t + 6 * t * t) * w_1d[idx] + (1 - 4 * t + 3 * t * t) * m[idx] * dx + (6 * t - 6 * t * t) * w_1d[idx + 1]
             + (3 * t * t - 2 * t) * m[idx + 1] * dx) / dx
    slope = torch.where(eta > eta_1d[-1], torch.zeros_like(slope), slope)
    re = (torch.fft.rfftn(W, dim=(0, 1, 2)).conj() * G).real / g.N * _weights(g).unsqueeze(3)
    kk = og.kabs.unsqueeze(3)
    scal = torch.where(kk != 0, slope / nodes * re / torch.where(kk != 0, kk, torch.ones_like(kk)), torch.zeros_like(re)).sum(3)
    return out - _tensor_sum(g, scal) / g.vol
