"""Two-point (Barzilai-Borwein) gradient descent with the interface of the reference's ``TPGD``
(src/professad/_optimizers/tpgd/two_point_gradient_descent.py:25-65):
alpha = <dx, dx> / <dx, dg>; the fixed ``lr`` is used on the first step, when <dx, dg> = 0, or
when alpha <= 0 (a maximum)."""
import torch
from torch.optim import Optimizer


class TPGD(Optimizer):

    def __init__(self, params, lr=1e-1):
        if lr <= 0.0:
            raise ValueError('Invalid initial learning rate: {} - should be > 0'.format(lr))
        super().__init__(params, dict(lr=lr))
        assert len(self.param_groups) == 1, "TPGD doesn't support per-parameter options (parameter groups)"
        self.iter = 0
        self._params = self.param_groups[0]['params']

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        num = den = 0.0
        for p in self._params:
            if p.grad is None:
                continue
            st = self.state[p]
            if self.iter != 0:
                dx = p.data - st['x_prev']
                dg = p.grad.data - st['g_prev']
                # one device->host read for both inner products
                pair = torch.stack([torch.sum(dx * dx), torch.sum(dx * dg)]).tolist()
                num += pair[0]
                den += pair[1]
                st['x_prev'].copy_(p.data)
                st['g_prev'].copy_(p.grad.data)
            else:
                st['x_prev'] = p.data.clone()
                st['g_prev'] = p.grad.data.clone()
        alpha = self.param_groups[0]['lr']
        if self.iter != 0 and den != 0:
            bb = num / den
            if bb > 0:
                alpha = bb
        for p in self._params:
            if p.grad is not None:
                p.data.add_(p.grad.data, alpha=-alpha)
        self.iter += 1
        return loss
