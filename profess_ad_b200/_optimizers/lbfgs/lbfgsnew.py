"""Fixed-step limited-memory BFGS with the interface and stopping rules of the reference's
``LBFGSNew`` (src/professad/_optimizers/lbfgs/lbfgsnew.py:512-769) for the configuration
``System.optimize_density`` uses: ``line_search_fn=False, batch_mode=False``.

Written from the algorithm, not from the reference file: the (y, s) history lives in two
preallocated (m, N) ring buffers, the two-loop recursion works on rows of them, and the
curvature products are computed once per update.  ``line_search_fn=True`` (geometry optimisation,
system.py:937-1198) selects the strong-Wolfe line search of lbfgsnew.py:208-509 (Fletcher's bracketing +
zoom with finite-difference slopes), restated in ``_WolfeSearch`` below; ``batch_mode=True`` (stochastic
training) is outside this path and raises.

Semantics kept exactly:
  * first-ever step length t = min(1, 1/|g|_1) * lr, afterwards t = lr;
  * history update only if  y.s > 1e-10 |s|^2, oldest pair dropped at ``history_size``;
  * H_diag = y.s / y.y of the most recent accepted pair;
  * the closure is re-evaluated after every inner iteration except the last (``max_iter``);
  * break on: max_iter, max_eval, |g|_1 <= tolerance_grad, g.d > -tolerance_change,
    |t d|_1 <= tolerance_change, |loss - prev_loss| < tolerance_change.

This host-driven class serves user-supplied (Python) energy terms.  When every term of a System is
native, ``System.optimize_density`` runs the same algorithm device-resident through the C ABI
(``pad_lbfgs_*``) instead, with no host synchronisation inside an outer iteration.
"""
import math

import torch
from torch.optim import Optimizer


class _WolfeSearch:
    """Strong-Wolfe line search along ``d`` from the current point (lbfgsnew.py:208-509, after Fletcher, Practical
    Methods of Optimization): bracketing phase with at most three trial steps starting at 10 lr, then a zoom phase
    of at most four cubic-interpolation refinements; directional derivatives by central differences of the closure
    with half-width ``h``.  The variable is restored to its starting value before ``run`` returns."""
    SIGMA, RHO, T1, T2, T3 = 0.1, 0.01, 9.0, 0.1, 0.5

    def __init__(self, x, d, closure, lr):
        self.x, self.d, self.closure, self.lr = x, d, closure, lr
        self.x0 = x.data.clone()
        self.evals = 0

    def phi(self, alpha):
        self.x.data.copy_(self.x0)
        self.x.data.add_(self.d.view_as(self.x.data), alpha=alpha)
        self.evals += 1
        v = self.closure()
        return float(v.detach()) if isinstance(v, torch.Tensor) else float(v)

    def slope(self, alpha, h):
        return (self.phi(alpha + h) - self.phi(alpha - h)) / (2.0 * h)

    def restore(self):
        self.x.data.copy_(self.x0)

    def interpolate(self, a, b, h):
        """Minimiser of the cubic through (a, phi, phi') and (b, phi, phi'), clipped to the better end point
        (lbfgsnew.py:334-424); a > b is allowed."""
        f0, f0d = self.phi(a), self.slope(a, h)
        f1, f1d = self.phi(b), self.slope(b, h)
        aa = 3.0 * (f0 - f1) / (b - a) + f1d - f0d
        disc = aa * aa - f0d * f1d
        if disc <= 0.0:
            return a if f0 < f1 else b
        cc = math.sqrt(disc)
        if f1d - f0d + 2.0 * cc == 0.0:
            return 0.5 * (a + b)
        z0 = b - (f1d + cc - aa) * (b - a) / (f1d - f0d + 2.0 * cc)
        if z0 > max(a, b) or z0 < min(a, b):
            fz0 = f0 + f1
        else:
            fz0 = self.phi(a + z0 * (b - a))          # sic: the reference samples a + z0 (b - a), lbfgsnew.py:398
        if f0 < f1 and f0 < fz0:
            return a
        if f1 < fz0:
            return b
        return z0

    def zoom(self, a, b, phi0, g0, h):
        """lbfgsnew.py:429-509"""
        aj, bj, alphaj = a, b, a
        for _ in range(4):
            alphaj = self.interpolate(aj + self.T2 * (bj - aj), bj - self.T3 * (bj - aj), h)
            phi_j = self.phi(alphaj)
            phi_aj = self.phi(aj)
            if phi_j > phi0 + self.RHO * alphaj * g0 or phi_j >= phi_aj:
                bj = alphaj
                continue
            gj = self.slope(alphaj, h)
            if (aj - alphaj) * gj <= h:               # round-off termination (Fletcher p. 38)
                return alphaj
            if abs(gj) <= -self.SIGMA * g0:
                return alphaj
            if gj * (bj - aj) >= 0.0:
                bj = aj
            aj = alphaj
        return alphaj

    def run(self, h):
        """lbfgsnew.py:208-331; returns the step length."""
        try:
            alphak = self.lr
            phi0 = self.phi(0.0)
            tol = min(phi0 * 0.01, 1e-6)
            g0 = self.slope(0.0, h)
            if abs(g0) < 1e-12:
                return 1.0
            mu = (tol - phi0) / (self.RHO * g0)
            if math.isnan(mu):
                return 1.0
            alphai, alphai1, phi_prev = 10.0 * self.lr, 0.0, phi0
            ci = 1
            while ci < 4:
                phi_i = self.phi(alphai)
                if phi_i < tol:
                    return alphai
                if phi_i > phi0 + alphai * g0 or (ci > 1 and phi_i >= phi_prev):
                    return self.zoom(alphai1, alphai, phi0, g0, h)
                gi = self.slope(alphai, h)
                if abs(gi) <= -self.SIGMA * g0:
                    return alphai
                if gi >= 0.0:
                    return self.zoom(alphai, alphai1, phi0, g0, h)
                if mu <= 2.0 * alphai - alphai1:
                    alphai1, alphai = alphai, mu
                else:
                    # (the previous trial step is deliberately NOT advanced on this branch: lbfgsnew.py:312-318)
                    lo, hi = 2.0 * alphai - alphai1, min(mu, alphai + self.T1 * (alphai - alphai1))
                    alphai = self.interpolate(lo, hi, h)
                phi_prev = phi_i
                ci += 1
            return alphak
        finally:
            self.restore()


class LBFGSNew(Optimizer):

    def __init__(self, params, lr=1, max_iter=10, max_eval=None, tolerance_grad=1e-5, tolerance_change=1e-9,
                 history_size=7, line_search_fn=False, batch_mode=False, max_step=None):
        # max_step (not in the reference; None = the reference's behaviour exactly): restart guard.  When a quasi-Newton
        # move would change some component of the variable by more than max_step, the history is dropped and the move is
        # the first-iteration steepest-descent step instead.  The geometry drivers call step() once per outer iteration
        # with a NEW objective (re-optimised density), so the pair (s, y) formed across two calls mixes a tiny s with a y
        # that is dominated by the change of objective; when that pair happens to be nearly orthogonal (y.s -> 0+) it
        # passes the curvature test and scales the direction by 1 / cos^2 -- the reference then throws the ions across
        # many cells and never recovers.  The guard only acts on such moves; all other iterates are unchanged.
        self._max_step = None if max_step is None else float(max_step)
        self.restarts = 0
        if batch_mode:
            raise NotImplementedError('LBFGSNew: batch mode (stochastic training) is outside the density / geometry '
                                      'optimisation path')
        self._line_search = bool(line_search_fn)
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        defaults = dict(lr=lr, max_iter=max_iter, max_eval=max_eval, tolerance_grad=tolerance_grad,
                        tolerance_change=tolerance_change, history_size=history_size)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("LBFGS doesn't support per-parameter options (parameter groups)")
        self._params = self.param_groups[0]['params']
        if len(self._params) != 1:
            raise ValueError('LBFGSNew (B200 build) optimises a single flat variable (chi)')
        self._x = self._params[0]
        self._n_iter_total = 0
        self._hist_y = self._hist_s = None
        self._order = []            # ring-buffer rows, oldest first
        self._H = 1.0
        self._d = None
        self._t = None
        self._prev_g = None
        self._prev_loss = None
        self.func_evals = 0

    def _grad(self):
        g = self._x.grad
        if g is None:
            return torch.zeros_like(self._x.data).reshape(-1)
        return g.data.reshape(-1)

    def _push_pair(self, y, s):
        m = self.param_groups[0]['history_size']
        if self._hist_y is None:
            self._hist_y = torch.empty((m, y.numel()), dtype=y.dtype, device=y.device)
            self._hist_s = torch.empty_like(self._hist_y)
        if len(self._order) == m:
            row = self._order.pop(0)
        else:
            row = len(self._order)
        self._hist_y[row].copy_(y)
        self._hist_s[row].copy_(s)
        self._order.append(row)

    def _direction(self, g):
        """two-loop recursion: d = -H g"""
        rows = self._order
        Y, S = self._hist_y, self._hist_s
        rho = [1.0 / float(torch.dot(Y[r], S[r])) for r in rows]
        q = g.neg()
        alpha = [0.0] * len(rows)
        for k in range(len(rows) - 1, -1, -1):
            alpha[k] = float(torch.dot(S[rows[k]], q)) * rho[k]
            q.add_(Y[rows[k]], alpha=-alpha[k])
        d = q.mul_(self._H)
        for k in range(len(rows)):
            beta = float(torch.dot(Y[rows[k]], d)) * rho[k]
            d.add_(S[rows[k]], alpha=alpha[k] - beta)
        return d

    def step(self, closure):
        grp = self.param_groups[0]
        lr, max_iter, max_eval = grp['lr'], grp['max_iter'], grp['max_eval']
        tol_g, tol_c = grp['tolerance_grad'], grp['tolerance_change']

        orig_loss = closure()
        loss = float(orig_loss.detach()) if isinstance(orig_loss, torch.Tensor) else float(orig_loss)
        evals = 1
        self.func_evals += 1
        g = self._grad()
        g_l1 = float(g.abs().sum())
        if g_l1 <= tol_g:
            return orig_loss
        g_l2 = float(g.norm())
        it = 0
        d, t = self._d, self._t
        while it < max_iter and not math.isnan(g_l2):
            it += 1
            self._n_iter_total += 1
            if self._n_iter_total == 1:
                d = g.neg()
                self._order = []
                self._H = 1.0
            else:
                y = g - self._prev_g
                s = d * t
                ys = float(torch.dot(y, s))
                s_l2 = float(s.norm())
                if ys > 1e-10 * s_l2 * s_l2:
                    self._push_pair(y, s)
                    self._H = ys / float(torch.dot(y, y))
                if math.isnan(self._H):
                    print('Warning H_diag nan')
                d = self._direction(g)
            if self._prev_g is None:
                self._prev_g = g.clone()
            else:
                self._prev_g.copy_(g)
            self._prev_loss = loss
            t = min(1.0, 1.0 / g_l1) * lr if self._n_iter_total == 1 else lr
            gtd = float(torch.dot(g, d))
            if math.isnan(gtd):
                print('Warning grad norm infinite')
            if self._line_search:
                # lbfgsnew.py:691-707: phi(alpha) = E(x + alpha d) sampled through the closure, step 1e-6 for the slopes
                search = _WolfeSearch(self._x, d.clone(), closure, lr)
                t = search.run(1e-6)
                self.func_evals += search.evals
                if math.isnan(t):
                    print('Warning: stepsize nan')
                    t = lr
            if self._max_step is not None and float(d.abs().max()) * abs(t) > self._max_step:
                self.restarts += 1
                self._order = []
                self._H = 1.0
                d = g.neg()
                t = min(1.0, 1.0 / g_l1) * lr
                gtd = float(torch.dot(g, d))
            self._x.data.add_(d.view_as(self._x.data), alpha=t)
            if it != max_iter:
                loss = closure()
                loss = float(loss.detach()) if isinstance(loss, torch.Tensor) else float(loss)
                g = self._grad()
                g_l1 = float(g.abs().sum())
                if math.isnan(g_l1):
                    print('Warning: gradient nan')
                    break
                evals += 1
                self.func_evals += 1
            if it == max_iter or evals >= max_eval or g_l1 <= tol_g:
                break
            if gtd > -tol_c or float(d.abs().sum()) * abs(t) <= tol_c:
                break
            if abs(loss - self._prev_loss) < tol_c:
                break
        self._d, self._t = d, t
        return orig_loss
