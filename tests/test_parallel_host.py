"""Host side of the slab decomposition on CPU, world_size 2 over gloo: the transpose bookkeeping the C library
uses (block r of the send buffer = the y range of rank r; block r of the receive buffer = the x range of rank r)
reproduces a full rfftn / irfftn, and parallel.TorchDistComm / slab_bounds behave as documented.
The local transforms here are torch.fft on CPU tensors: this is a test model of csrc/plan.cu, not a product path."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _slab_rfftn(local, comm, n1):
    """Model of fft_forward_slab: (n0_loc, n1, n2) real -> (n0, n1_loc, nzh) complex."""
    w = comm.world
    n0_loc, _, _ = local.shape
    n1_loc = n1 // w
    a = torch.fft.rfft2(local, dim=(1, 2))                                   # (n0_loc, n1, nzh)
    nzh = a.shape[2]
    send = a.reshape(n0_loc, w, n1_loc, nzh).permute(1, 0, 2, 3).contiguous().reshape(-1)
    recv = torch.empty_like(send)
    comm.all_to_all(recv, send)
    b = recv.reshape(w * n0_loc, n1_loc, nzh)                                # block r = x range of rank r
    return torch.fft.fft(b, dim=0)


def _slab_irfftn(spec, comm, n1, n2):
    """Model of fft_inverse_slab (unnormalised like cuFFT): (n0, n1_loc, nzh) -> (n0_loc, n1, n2)."""
    w = comm.world
    n0, n1_loc, nzh = spec.shape
    n0_loc = n0 // w
    send = (torch.fft.ifft(spec, dim=0) * n0).contiguous().reshape(-1)       # already blocked by x range
    recv = torch.empty_like(send)
    comm.all_to_all(recv, send)
    a = recv.reshape(w, n0_loc, n1_loc, nzh).permute(1, 0, 2, 3).reshape(n0_loc, n1, nzh)
    return torch.fft.irfft2(a, s=(n1, n2), dim=(1, 2)) * (n1 * n2)


def _worker(rank, world, port, shape):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from profess_ad_b200 import parallel
        comm = parallel.TorchDistComm()
        assert (comm.rank, comm.world) == (rank, world)
        gen = torch.Generator().manual_seed(5)
        full = torch.rand(*shape, dtype=torch.double, generator=gen)
        lo, hi = parallel.slab_bounds(shape[0], rank, world)
        assert hi - lo == shape[0] // world
        local = parallel.local_slab(full, comm=comm)
        assert torch.equal(local, full[lo:hi])
        spec = _slab_rfftn(local, comm, shape[1])
        ref = torch.fft.rfftn(full)
        n1_loc = shape[1] // world
        assert (spec - ref[:, rank * n1_loc:(rank + 1) * n1_loc]).abs().max().item() < 1e-12
        back = _slab_irfftn(spec, comm, shape[1], shape[2]) / full.numel()
        assert (back - local).abs().max().item() < 1e-13
        s = torch.tensor([float(rank + 1), 2.0], dtype=torch.double)
        comm.all_reduce(s)
        assert s.tolist() == [float(sum(range(1, world + 1))), 2.0 * world]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('shape', [(8, 6, 10), (6, 4, 7)])
def test_slab_transpose_model_world2_gloo(shape):
    port = 29620 + shape[0]
    mp.spawn(_worker, args=(2, port, shape), nprocs=2, join=True)


def test_slab_bounds_and_errors():
    from profess_ad_b200 import parallel
    assert parallel.slab_bounds(12, 2, 4) == (6, 9)
    with pytest.raises(ValueError):
        parallel.slab_bounds(10, 0, 4)
    with pytest.raises(RuntimeError):
        parallel.local_slab(torch.zeros(4, 2, 2))


class _FakeSystem:
    """Duck-typed stand-in for System (the real one needs a GPU): E(V) is a Birch-Murnaghan curve."""
    GPa_per_atomic, eV_per_Ha, A_per_b = 29421.02648438959, 27.211386245988, 0.529177210903

    def __init__(self):
        self.box = 16.9 ** (1 / 3) * torch.eye(3, dtype=torch.double)
        self.optimised = 0

    def volume(self, units):
        return abs(torch.linalg.det(self.box).item())

    def lattice_vectors(self, units):
        return self.box

    def set_lattice(self, box, units='a'):
        self.box = box.clone()

    def optimize_density(self, **kw):
        self.optimised += 1

    def ion_count(self):
        return 1

    def energy(self, units):
        from profess_ad_b200.elastic_tools import birch_murnaghan
        import numpy as np
        return float(birch_murnaghan(np.array([self.volume('a3')]), 0.49, 4.3, -57.2, 16.76)[0])


def _eos_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from profess_ad_b200 import parallel
        made = []

        def make():
            made.append(_FakeSystem())
            return made[-1]
        params, err = parallel.eos_fit(make, f=0.05, N=9, eos='bm')
        # 9 volumes over 2 ranks: 5 + 4 density optimisations, the same fit on every rank
        assert made[0].optimised == (5 if rank == 0 else 4)
        s = made[0]
        k0 = 0.49 * s.GPa_per_atomic / (s.eV_per_Ha / s.A_per_b ** 3)
        assert abs(params[0] - k0) < 1e-6 * k0 and abs(params[2] + 57.2) < 1e-8 and abs(params[3] - 16.76) < 1e-7
        out.put((rank, [float(p) for p in params]))
    finally:
        dist.destroy_process_group()


def test_distributed_eos_fit_world2_gloo():
    """parallel.eos_fit: the volumes of the scan are dealt round-robin to the ranks (independent systems, no
    collective in the loop), gathered once, and every rank returns the same fit."""
    out = mp.get_context('spawn').SimpleQueue()
    mp.spawn(_eos_worker, args=(2, 29655, out), nprocs=2, join=True)
    got = dict(out.get() for _ in range(2))
    assert got[0] == got[1]
