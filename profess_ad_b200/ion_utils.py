"""Ion-related terms (src/professad/ion_utils.py): recpot reader, structure factor (exact and
particle-mesh Ewald), lattice sum -> v_ext, ion-ion electrostatic energy.

These run once per geometry, not per optimiser iteration, and are
written with device-resident torch ops (no Python loop over ions, no O(N_k N_ion) temporaries); they
produce the constant ``v_ext`` the CUDA hot path consumes.
"""
import ctypes

import numpy as np
import torch

from .functional_tools import wavevecs, interpolate

bohr = 0.529177208607388
hartree_to_ev = 27.2113834279111
pot_conv_factor = 1 / (bohr * bohr * bohr * hartree_to_ev)

_recpot_cache = {}


def _read_recpot(path):
    """Parse a CASTEP-style .recpot file once: comment block up to 'END COMMENT', a '3 5' line,
    k_max [1/Angstrom], then three values per line [eV Angstrom^3] (ion_utils.py:20-81)."""
    hit = _recpot_cache.get(path)
    if hit is not None:
        return hit
    values = []
    with open(path, 'r') as fh:
        for line in fh:
            if 'END COMMENT' in line:
                break
        fh.readline()
        k_max = float(fh.readline()) * bohr
        for line in fh:
            cols = line.split()
            if len(cols) == 3:
                values += cols
    pot = np.asarray(values, dtype=np.float64) * pot_conv_factor
    ks, dk = np.linspace(0, k_max, pot.size, retstep=True)
    z = round((pot[1] - pot[0]) * dk * dk / (-4 * np.pi))
    _recpot_cache[path] = (ks, pot, z)
    return _recpot_cache[path]


def get_ion_charge(path):
    """Ion charge from the small-k Coulomb tail of the tabulated potential (ion_utils.py:20-46)."""
    return _read_recpot(path)[2]


def interpolate_recpot(path, ks_interp):
    """Reciprocal-space ionic potential on the |k| grid (ion_utils.py:49-81): the Coulomb tail is
    added back before the cubic Hermite interpolation and removed afterwards."""
    ks, pot, z = _read_recpot(path)
    smooth = pot.copy()
    smooth[1:] += 4 * np.pi * z / (ks[1:] * ks[1:])
    ks_t = torch.as_tensor(ks, dtype=torch.double, device=ks_interp.device)
    sm_t = torch.as_tensor(smooth, dtype=torch.double, device=ks_interp.device)
    val = interpolate(ks_t, sm_t, torch.minimum(ks_interp, ks_t[-1]))
    nz = ks_interp != 0
    safe = torch.where(nz, ks_interp, torch.ones_like(ks_interp))
    return torch.where(nz, val - 4 * np.pi * z / safe.pow(2), val)


class PadSpecies(ctypes.Structure):
    """pad_species of include/professad_b200.h"""
    _fields_ = [('table_dev', ctypes.c_void_p), ('n_table', ctypes.c_int), ('k_max', ctypes.c_double),
                ('z', ctypes.c_double), ('frac_dev', ctypes.c_void_p), ('n_ions', ctypes.c_int)]


_table_cache = {}


def _species_table(path, device):
    """[k | v(k) + 4 pi z / k^2] on the device: the smooth table the reference interpolates (ion_utils.py:62-66)."""
    key = (path, str(device))
    hit = _table_cache.get(key)
    if hit is None:
        ks, pot, z = _read_recpot(path)
        smooth = pot.copy()
        smooth[1:] += 4 * np.pi * z / (ks[1:] * ks[1:])
        tab = torch.as_tensor(np.concatenate([ks, smooth]), dtype=torch.double, device=device).contiguous()
        hit = (tab, int(ks.size), float(ks[-1]), float(z))
        _table_cache[key] = hit
    return hit


def _pad_species_array(species, device):
    """ctypes array of pad_species for [(recpot path, (n, 3) fractional coordinates), ...]; also returns the
    tensors that must stay alive for the duration of the call."""
    arr = (PadSpecies * len(species))()
    keep = []
    for i, (path, frac) in enumerate(species):
        tab, n_tab, k_max, z = _species_table(path, device)
        fr = frac.detach().to(device=device, dtype=torch.double).contiguous()
        keep += [tab, fr]
        arr[i] = PadSpecies(tab.data_ptr(), n_tab, k_max, z, fr.data_ptr(), int(fr.shape[0]))
    return arr, keep


def ionic_potential(box_vecs, shape, species, pme_order=None):
    """v_ext on the grid from the ion positions, one C-ABI call: structure factor x interpolated local pseudopotential
    -> c2r.  ``species`` = [(recpot path, (n, 3) fractional coordinates), ...].  ``pme_order`` None: exact structure factor
    (pad_ionic_potential); even n: particle-mesh Ewald of that spline order (pad_ionic_potential_pme).
    Inside ``parallel.slab(...)`` ``shape`` is the LOCAL slab shape and the local slab of v_ext is returned.
    Replaces System.__potential_from_ions (system.py:183-205)."""
    from . import _native
    dev = box_vecs.device
    v_ext = torch.empty(tuple(int(s) for s in shape), dtype=torch.double, device=dev)
    _native.require_cuda(v_ext, 'box_vecs')
    plan = _native.get_plan(box_vecs, v_ext)
    arr, keep = _pad_species_array(species, dev)
    if pme_order is None:
        _native.check(plan.lib.pad_ionic_potential(plan.handle, arr, len(species), _native.ptr(v_ext), _native.stream_ptr(dev)))
    else:
        assert (pme_order % 2 == 0) & (pme_order >= 2), 'Requires even order n ≥ 2'
        _native.check(plan.lib.pad_ionic_potential_pme(plan.handle, arr, len(species), int(pme_order), _native.ptr(v_ext),
                                                       _native.stream_ptr(dev)))
    del keep
    return v_ext


def ion_electron_forces(box_vecs, den, species, pme_order=None):
    """-d IonElectron / d R_I at fixed density, (N_ion, 3) Cartesian in Ha/bohr (pad_ion_forces); the IonElectron
    part of System.__compute_forces (system.py:913-925).  ``pme_order`` (even, <= 32): the forces of the
    particle-mesh structure factor -- what the reference's autograd gives when the System was built with pme_order
    (pad_ion_forces_pme).  Inside ``parallel.slab(...)`` the partial sums of the ranks are added here."""
    from . import _native, parallel
    _native.require_cuda(den)
    den = den.detach().contiguous()
    plan = _native.get_plan(box_vecs, den)
    arr, keep = _pad_species_array(species, den.device)
    n_ions = sum(int(f.shape[0]) for _, f in species)
    forces = torch.zeros((n_ions, 3), dtype=torch.double, device=den.device)
    ctx = parallel.current()
    if pme_order is not None:
        _native.check(plan.lib.pad_ion_forces_pme(plan.handle, arr, len(species), int(pme_order), _native.ptr(den), _native.ptr(forces),
                                                  _native.stream_ptr(den.device)))
    else:
        _native.check(plan.lib.pad_ion_forces(plan.handle, arr, len(species), _native.ptr(den), _native.ptr(forces),
                                              _native.stream_ptr(den.device)))
    del keep
    if ctx is not None:
        ctx.comm.all_reduce(forces.view(-1))
    return forces


def ion_electron_stress(box_vecs, den, species, pme_order=None):
    """IonElectron part of the stress, (3, 3) in Ha/bohr^3 (pad_ion_stress; system.py:927-935): ions at fixed
    fractional coordinates, electron number conserved.  ``pme_order``: with the particle-mesh structure factor
    (pad_ion_stress_pme)."""
    from . import _native, parallel
    _native.require_cuda(den)
    den = den.detach().contiguous()
    plan = _native.get_plan(box_vecs, den)
    arr, keep = _pad_species_array(species, den.device)
    out = torch.empty(9, dtype=torch.double, device=den.device)
    if pme_order is not None:
        _native.check(plan.lib.pad_ion_stress_pme(plan.handle, arr, len(species), int(pme_order), _native.ptr(den), _native.ptr(out), 0,
                                                  _native.stream_ptr(den.device)))
    else:
        _native.check(plan.lib.pad_ion_stress(plan.handle, arr, len(species), _native.ptr(den), _native.ptr(out), 0,
                                              _native.stream_ptr(den.device)))
    del keep
    return out.reshape(3, 3)


def hermitian_symmetrize(G, n2):
    """Make a half-spectrum Hermitian-consistent on its self-conjugate planes (j2 = 0 and, for even
    n2, j2 = n2/2): G <- (G(p) + conj G(pbar)) / 2.  This is what the reference's CPU irfftn
    (c2c over axes 0,1 then c2r over axis 2) effectively does with the non-Hermitian
    "Nyquist made positive" spectra; doing it explicitly makes any c2r implementation agree."""
    n0, n1, nzh = G.shape
    i0 = (-torch.arange(n0, device=G.device)) % n0
    i1 = (-torch.arange(n1, device=G.device)) % n1
    out = G.clone()
    for j2 in [0] + ([nzh - 1] if n2 % 2 == 0 else []):
        P = G[:, :, j2]
        out[:, :, j2] = 0.5 * (P + P[i0][:, i1].conj())
    return out


def _irfftn_reference_semantics(G, shape, norm='backward'):
    shape = tuple(int(s) for s in shape)
    return torch.fft.irfftn(hermitian_symmetrize(G, shape[2]), shape, norm=norm)


def structure_factor(box_vecs, shape, cart_ion_coords, chunk=64):
    """Exact S(q) = sum_i exp(-i q.r_i) (ion_utils.py:121-137), accumulated over chunks of ions so the
    temporary is N_k x chunk instead of N_k x N_ion."""
    kx, ky, kz, _ = wavevecs(box_vecs, shape)
    S = torch.zeros(kx.shape, dtype=torch.complex128, device=box_vecs.device)
    for start in range(0, cart_ion_coords.shape[0], chunk):
        r = cart_ion_coords[start:start + chunk]
        phase = kx.unsqueeze(-1) * r[:, 0] + ky.unsqueeze(-1) * r[:, 1] + kz.unsqueeze(-1) * r[:, 2]
        S = S + torch.complex(torch.cos(phase), -torch.sin(phase)).sum(-1)
    return S


def cardinal_b_spline_values(x, order):
    """[M_n(x + i) for i = 0..n-1] for x in [0, 1) (ion_utils.py:140-204), by the Cox-de Boor
    recursion M_n[i] = ((x+i) M_{n-1}[i] + (n-x-i) M_{n-1}[i-1]) / (n-1), written functionally
    (no in-place updates, so it stays differentiable)."""
    assert torch.all(x >= 0.0) and torch.all(x < 1.0), 'Requires 0 ≤ x < 1'
    assert order >= 2, 'Requires order n ≥ 2'
    zero = torch.zeros_like(x)
    M = [x, 1 - x] + [zero] * (order - 2)
    for n in range(3, order + 1):
        new = []
        for i in range(order):
            if i >= n:
                new.append(zero)
                continue
            left = (x + i) * M[i]
            right = (n - x - i) * M[i - 1] if i >= 1 else zero
            new.append((left + right) / (n - 1))
        M = new
    return torch.stack(M)


def exponential_spline_b(m, N, order):
    """Euler exponential spline coefficient b(m) of the smooth PME (ion_utils.py:207-215)."""
    M = cardinal_b_spline_values(torch.zeros_like(m), order)
    i = torch.arange(0, order, dtype=torch.double, device=m.device).unsqueeze(1)
    denom = torch.sum(M * torch.exp(1j * 2 * np.pi * m * (i - 1) / N), axis=0)
    return torch.exp(1j * 2 * np.pi * m * (order - 1) / N) / denom


def structure_factor_spline(box_vecs, shape, cart_ion_coords, order):
    """Particle-mesh Ewald structure factor (ion_utils.py:218-286; Essmann et al. 1995).  CUDA tensors: native spreading
    kernel + r2c + exponential-spline factors (pad_pme_structure_factor); CPU tensors (host-side checks): the torch
    restatement below -- B-spline charge spreading as ONE scatter-add over (ion, order^3 stencil), FFT, b factors."""
    if cart_ion_coords.is_cuda and not cart_ion_coords.requires_grad and not box_vecs.requires_grad and order <= 32:
        from . import _native
        dev = cart_ion_coords.device
        frac = torch.matmul(cart_ion_coords, torch.linalg.inv(box_vecs)).double().contiguous()
        probe = torch.empty(tuple(int(n) for n in shape), dtype=torch.double, device=dev)
        plan = _native.get_plan(box_vecs, probe)
        S = torch.empty((int(shape[0]), int(shape[1]), int(shape[2]) // 2 + 1), dtype=torch.complex128, device=dev)
        _native.check(plan.lib.pad_pme_structure_factor(plan.handle, _native.ptr(frac), int(frac.shape[0]), int(order), _native.ptr(S),
                                                        _native.stream_ptr(dev)))
        return S
    return _structure_factor_spline_torch(box_vecs, shape, cart_ion_coords, order)


def _structure_factor_spline_torch(box_vecs, shape, cart_ion_coords, order):
    N0, N1, N2 = (int(s) for s in shape)
    frac = torch.matmul(cart_ion_coords, torch.linalg.inv(box_vecs))
    frac = frac - torch.floor(frac)
    frac = frac - torch.floor(frac)
    assert torch.all(frac >= 0) and torch.all(frac < 1), 'Fractional ionic coordinates don\'t all lie in [0,1)'
    dev = box_vecs.device
    dims = torch.tensor([N0, N1, N2], dtype=torch.double, device=dev)
    u = frac * dims
    fl = torch.floor(u)
    w = [cardinal_b_spline_values(u[:, a] - fl[:, a], order) for a in range(3)]          # (order, n_ion) each
    o = torch.arange(order, dtype=torch.int64, device=dev).unsqueeze(1)
    idx = [torch.remainder(o - fl[:, a].to(torch.int64), n) for a, n in enumerate((N0, N1, N2))]
    lin = (idx[0][:, None, None, :] * N1 + idx[1][None, :, None, :]) * N2 + idx[2][None, None, :, :]
    wt = w[0][:, None, None, :] * w[1][None, :, None, :] * w[2][None, None, :, :]
    Q = torch.zeros(N0 * N1 * N2, dtype=torch.double, device=dev)
    Q.index_add_(0, lin.reshape(-1), wt.reshape(-1))
    Q_ft = torch.fft.rfftn(Q.reshape(N0, N1, N2))
    b = [exponential_spline_b(torch.arange(0, Q_ft.shape[a], dtype=torch.double, device=dev), n, order)
         for a, n in enumerate((N0, N1, N2))]
    return torch.conj(b[0][:, None, None] * b[1][None, :, None] * b[2][None, None, :] * Q_ft)


def lattice_sum(box_vecs, shape, cart_ion_coords, f_tilde, order=None):
    """F(r) = irfftn(S(q) f~(q)) / vol (ion_utils.py:88-118)."""
    if order is None:
        S = structure_factor(box_vecs, shape, cart_ion_coords)
    else:
        assert (order % 2 == 0) & (order >= 2), 'Requires even order n ≥ 2'
        S = structure_factor_spline(box_vecs, shape, cart_ion_coords, order)
    return _irfftn_reference_semantics(S * f_tilde, shape, norm='forward') / torch.abs(torch.linalg.det(box_vecs))


def _pair_list(box_vecs, coords, Rc, chunk=None):
    """All (i, j, lattice shift) with |r_j + shift - r_i| < Rc, i != j or shift != 0.  Replaces
    torch_nl.compute_neighborlist (ion_utils.py:313-316): images are enumerated from the interplanar
    spacings and filtered on the device in chunks of shifts."""
    dev = coords.device
    heights = 1.0 / torch.sqrt(torch.sum(torch.linalg.inv(box_vecs.detach()).T.pow(2), 1))
    reps = [int(np.ceil(float(Rc) / h.item())) + 1 for h in heights]
    rng = [torch.arange(-r, r + 1, dtype=torch.double, device=dev) for r in reps]
    shifts = torch.stack(torch.meshgrid(*rng, indexing='ij'), -1).reshape(-1, 3)
    # prune images whose cell cannot reach the cutoff sphere
    centre = shifts @ box_vecs.detach()
    diag = torch.linalg.norm(box_vecs.detach().sum(0)) + torch.linalg.norm(box_vecs.detach(), dim=1).max()
    shifts = shifts[centre.norm(dim=1) < float(Rc) + 2 * diag]
    n = coords.shape[0]
    if chunk is None:
        chunk = max(1, (1 << 22) // (n * n))
    ii, jj = torch.meshgrid(torch.arange(n, device=dev), torch.arange(n, device=dev), indexing='ij')
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    base = coords.detach()[jj] - coords.detach()[ii]
    out_i, out_j, out_s = [], [], []
    for start in range(0, shifts.shape[0], chunk):
        sh = shifts[start:start + chunk]
        d = (base.unsqueeze(1) + (sh @ box_vecs.detach()).unsqueeze(0)).norm(dim=2)
        ok = d < float(Rc)
        ok &= ~((sh.abs().sum(1) == 0).unsqueeze(0) & (ii == jj).unsqueeze(1))
        pr, si = torch.nonzero(ok, as_tuple=True)
        out_i.append(ii[pr]); out_j.append(jj[pr]); out_s.append(sh[si])
    return torch.cat(out_i), torch.cat(out_j), torch.cat(out_s)


class _IonIonSum(torch.autograd.Function):
    """ion_interaction_sum on the device (csrc/ionion.cu): E, and in backward dE/dcoords and dE/dbox_vecs at fixed coords."""

    @staticmethod
    def forward(ctx, box_vecs, coords, charges, Rc, Rd):
        from . import _native
        lib = _native.load_library()
        dev = coords.device
        box_h = (ctypes.c_double * 9)(*[float(x) for x in box_vecs.detach().double().cpu().reshape(-1)])
        c = coords.detach().double().contiguous()
        z = charges.detach().to(device=dev, dtype=torch.double).contiguous()
        n = int(c.shape[0])
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        E = torch.empty((), dtype=torch.double, device=dev)
        dcart = torch.empty_like(c) if need else None
        dbox = torch.empty(9, dtype=torch.double, device=dev) if need else None
        nwork = int(lib.pad_ion_ion_work_doubles(box_h, n, float(Rc)))
        if nwork == 0:
            raise ValueError('Lattice vector matrix is not invertible.')
        work = torch.empty(nwork, dtype=torch.double, device=dev)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        _native.check(lib.pad_ion_ion(box_h, _native.ptr(c), _native.ptr(z), n, float(z.sum().item()), float(Rc), float(Rd),
                                      _native.ptr(E), _native.ptr(dcart), _native.ptr(dbox), _native.ptr(work), idx,
                                      _native.stream_ptr(dev)))
        if need:
            ctx.save_for_backward(dcart, dbox)
        ctx.box_device = box_vecs.device
        return E

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        dcart, dbox = ctx.saved_tensors
        g_box = (grad_out * dbox.reshape(3, 3)).to(ctx.box_device) if ctx.needs_input_grad[0] else None
        g_cart = grad_out * dcart if ctx.needs_input_grad[1] else None
        return g_box, g_cart, None, None, None


def ion_interaction_sum(box_vecs, coords, charges, Rc, Rd):
    """Real-space damped pairwise electrostatic sum in a neutralising background
    (ion_utils.py:293-333; Phys. Rev. Materials 2, 013806).  CUDA tensors: one native sweep over the (ion, image) candidates
    with closed-form derivatives (pad_ion_ion); CPU tensors (host-side checks only): the pair-list restatement below."""
    if coords.is_cuda:
        Rc_f = float(Rc.item()) if torch.is_tensor(Rc) else float(Rc)
        Rd_f = float(Rd.item()) if torch.is_tensor(Rd) else float(Rd)
        return _IonIonSum.apply(box_vecs, coords, charges, Rc_f, Rd_f)
    mi, mj, shifts = _pair_list(box_vecs, coords, Rc)
    rho = torch.sum(charges) / torch.abs(torch.linalg.det(box_vecs))
    Zi, Zj = charges[mi], charges[mj]
    Qi = torch.scatter_add(charges, 0, mi, Zj)
    aux = (0.75 / np.pi) * Qi / rho
    Ra = aux.sign() * aux.abs().pow(1 / 3)
    r_ij = (coords[mj] + shifts @ box_vecs - coords[mi]).norm(p=2, dim=1)
    E_local = torch.sum(0.5 * Zi * Zj * torch.erfc(r_ij / Rd) / r_ij)
    E_corr = torch.sum(-np.pi * charges * rho * Ra.square()
                       + np.pi * charges * rho * (Ra.square() - 0.5 * Rd * Rd) * torch.erf(Ra / Rd)
                       + np.sqrt(np.pi) * charges * rho * Ra * Rd * torch.exp(-Ra.square() / (Rd * Rd))
                       - charges.square() / np.sqrt(np.pi) / Rd)
    return E_local + E_corr
