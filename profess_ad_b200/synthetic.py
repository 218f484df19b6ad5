"""Deterministic synthetic inputs for benchmarks.

No files, no reference code: a smooth Al-like density on a cubic supercell.  bench.py uses this for
the B200 arm; the CPU baseline arm and the tests generate their inputs independently.
"""
import math

import torch


def smooth_supercell(n, side, device='cpu', x_range=None):
    """(box_vecs, den): cubic cell of side*a (a = 4.05 A in bohr), 4*side^3 Al atoms, 3 e-/atom, n^3 grid.

    den = n0 (1 + 0.3 cos(kX) cos(kY) cos(kZ) + 0.05 cos(2kX) cos(2kY)), k = 2 pi side, x = i / n.
    ``x_range=(lo, hi)`` returns only the planes lo <= i < hi of axis 0 (a rank's slab, see parallel.py).
    """
    dt = torch.double
    a = 4.05 / 0.529177210903
    L = side * a
    box = L * torch.eye(3, dtype=dt, device=device)
    x = torch.arange(n, dtype=dt, device=device) / n
    k = 2 * math.pi * side
    c1, c2 = torch.cos(k * x), torch.cos(2 * k * x)
    n0 = 12 * side ** 3 / L ** 3
    lo, hi = x_range if x_range is not None else (0, n)
    c1x, c2x = c1[lo:hi], c2[lo:hi]
    den = n0 * (1 + 0.3 * c1x[:, None, None] * c1[None, :, None] * c1[None, None, :]
                + 0.05 * (c2x[:, None, None] * c2[None, :, None]).expand(hi - lo, n, n))
    return box, den.contiguous()


def fcc_supercell(side, a_angstrom=4.05):
    """(box_vecs in bohr, fractional coordinates): side^3 conventional fcc cells, 4 side^3 atoms."""
    dt = torch.double
    a = a_angstrom / 0.529177210903
    basis = torch.tensor([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]], dtype=dt)
    r = torch.arange(side, dtype=dt)
    cells = torch.stack(torch.meshgrid(r, r, r, indexing='ij'), dim=-1).reshape(-1, 1, 3)
    frac = ((cells + basis[None]) / side).reshape(-1, 3)
    return side * a * torch.eye(3, dtype=dt), frac
