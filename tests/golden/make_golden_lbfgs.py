"""Trajectories of the UNMODIFIED reference LBFGSNew with its strong-Wolfe line search
(src/professad/_optimizers/lbfgs/lbfgsnew.py, line_search_fn=True -- what System.optimize_geometry uses) on two
small analytic test functions.  Run in the build container only:  python tests/golden/make_golden_lbfgs.py"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def rosen(x):
    return (100 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2).sum()


def quart(x):
    A = torch.linspace(0.5, 3, x.numel(), dtype=torch.double)
    return (A * x * x).sum() + 0.1 * (x ** 4).sum() + torch.sin(x).sum() * 0.3 + 2.0


CASES = {'rosen': (rosen, [-1.2, 1.0, 0.5, -0.3, 0.8]), 'quart': (quart, np.linspace(-1, 2, 12).tolist())}


def trajectory(cls, fn, x0, steps=12, **kw):
    x = torch.tensor(x0, dtype=torch.double, requires_grad=True)
    opt = cls([x], lr=0.1, history_size=8, max_iter=6, **kw)

    def closure():
        if torch.is_grad_enabled():
            opt.zero_grad()
        E = fn(x)
        if E.requires_grad:
            E.backward()
        return E
    out = []
    for _ in range(steps):
        opt.step(closure)
        out.append(x.detach().clone().numpy())
    return np.stack(out)


if __name__ == '__main__':
    spec = importlib.util.spec_from_file_location('ref_lbfgs', '/root/reference/src/professad/_optimizers/lbfgs/lbfgsnew.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    for name, (fn, x0) in CASES.items():
        out[name + '_linesearch'] = trajectory(mod.LBFGSNew, fn, x0, line_search_fn=True)
        out[name + '_fixed'] = trajectory(mod.LBFGSNew, fn, x0)
    np.savez_compressed(os.path.join(HERE, 'lbfgs_trajectories.npz'), **out)
    print({k: v.shape for k, v in out.items()})
