"""Slab-decomposed evaluation (profess_ad_b200/parallel.py, csrc/plan.cu "slab plans") against the CPU oracle.

The GPU test box has one device, so world > 1 runs as threads of one process sharing the GPU with
parallel.ThreadComm standing in for NCCL; the C library, the kernels and the index arithmetic are exactly
those of a multi-GPU run (scripts/slab_bench.py exercises the NCCL binding under torchrun)."""
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu


def _functionals():
    import profess_ad_b200.functionals as F
    from oracle import ofdft_oracle as orc
    return [
        ('WGC99', lambda: F.WangGovindCarter99().forward, lambda: orc.WangGovindCarter99()),
        ('WT', lambda: F.WangTeter, lambda: orc.WangTeter),
        ('WGC98', lambda: F.WangGovindCarter98, lambda: orc.WangGovindCarter98),
        ('Hartree', lambda: F.Hartree, lambda: orc.Hartree),
        ('vW', lambda: F.Weizsaecker, lambda: orc.Weizsaecker),
        ('TF', lambda: F.ThomasFermi, lambda: orc.ThomasFermi),
        ('PZ', lambda: F.PerdewZunger, lambda: orc.PerdewZunger),
        ('PBE', lambda: F.PerdewBurkeErnzerhof, lambda: orc.PerdewBurkeErnzerhof),
    ]


def _run_rank(comm, global_shape, box, den_global, make_f, out, idx, errors):
    from profess_ad_b200 import parallel
    try:
        dev = torch.device('cuda:0')
        with torch.cuda.stream(torch.cuda.Stream(dev)):
            with parallel.slab(global_shape, comm=comm):
                d = parallel.local_slab(den_global.to(dev)).requires_grad_(True)
                E = make_f()(box.to(dev), d)
                (g,) = torch.autograd.grad(E, d)
                torch.cuda.current_stream(dev).synchronize()
                out[idx] = (E.item(), g.cpu())
    except BaseException as e:      # noqa: BLE001
        errors.append(e)
        try:
            comm.shared.barrier.abort()
        except Exception:
            pass


def _evaluate_slabs(world, global_shape, box, den, make_f):
    from profess_ad_b200 import parallel
    out, errors = [None] * world, []
    if world == 1:
        _run_rank(parallel.SingleComm(), global_shape, box, den, make_f, out, 0, errors)
    else:
        shared = parallel.ThreadComm.Shared(world)
        threads = [threading.Thread(target=_run_rank, args=(parallel.ThreadComm(shared, r), global_shape, box, den,
                                                            make_f, out, r, errors)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=300)
    if errors:
        raise errors[0]
    energies = [o[0] for o in out]
    return energies, torch.cat([o[1] for o in out], dim=0)


@pytest.mark.parametrize('world,shape', [(1, (9, 10, 12)), (1, (8, 6, 7)), (2, (8, 6, 10)), (2, (12, 10, 9)), (4, (8, 12, 6))])
def test_slab_matches_oracle(world, shape):
    from oracle import ofdft_oracle as orc
    box, den = orc.synth_rough(shape, seed=17 + world, L=8.5)
    dV = abs(torch.linalg.det(box).item()) / den.numel()
    for name, make_f, make_o in _functionals():
        E_ref, V_ref = orc.energy_and_potential(box, den, make_o())
        energies, g = _evaluate_slabs(world, shape, box, den, make_f)
        for E in energies:          # every rank returns the global energy
            assert abs(E - E_ref.item()) <= 1e-10 * max(1.0, abs(E_ref.item())), (name, world, E, E_ref.item())
        assert max(energies) - min(energies) <= 1e-13 * max(1.0, abs(E_ref.item())), (name, energies)
        err = ((g / dV - V_ref).abs().max() / V_ref.abs().max()).item()
        assert err < 1e-9, (name, world, err)


def test_slab_rejects_wrong_slab_shape_and_unsupported_terms():
    import profess_ad_b200.functionals as F
    from oracle import ofdft_oracle as orc
    from profess_ad_b200 import parallel
    dev = torch.device('cuda:0')
    box, den = orc.synth_rough((8, 6, 10), seed=3)
    with parallel.slab((8, 6, 10), comm=parallel.SingleComm()):
        with pytest.raises(ValueError):
            F.ThomasFermi(box.to(dev), den[:4].contiguous().to(dev))
    with pytest.raises(ValueError):
        with parallel.slab((9, 6, 10), comm=parallel.ThreadComm(parallel.ThreadComm.Shared(2), 0)):
            F.ThomasFermi(box.to(dev), den.to(dev)[:4].contiguous())
