"""One large grid over several GPUs: slab-decomposed evaluation of the native functionals.

Real space is split along axis 0 (each rank holds ``n0 / world`` planes of the density), the half
spectrum along axis 1.  Every 3-D transform inside the C library becomes a local batched 2-D (y, z)
cuFFT, ONE all-to-all and a local strided 1-D (x) cuFFT (``csrc/plan.cu``, "slab plans"); energies and
the scalars the functionals need mid-evaluation (sum of n for n0 / n_ref) are all-reduced.  The library
itself does not link a communication library: it calls back into this module, which uses
``torch.distributed`` (NCCL over NVLink / NVSwitch) on the caller's current stream.

Usage (one process per GPU, ``torch.distributed`` initialised, the same script on every rank)::

    from profess_ad_b200 import parallel
    from profess_ad_b200.functionals import WangGovindCarter99, Hartree
    with parallel.slab(global_shape=(512, 512, 512)):
        den_local = parallel.local_slab(den_global)            # (n0 / world, n1, n2), or build it locally
        den_local.requires_grad_(True)
        E = WangGovindCarter99().forward(box_vecs, den_local)  # the GLOBAL energy, identical on every rank
        (g,) = torch.autograd.grad(E, den_local)               # this rank's slab of dE/dn * dV

Inside the context the ordinary functional callables (HuangCarter included) take LOCAL slabs, and
``parallel.optimize_density`` runs the device-resident L-BFGS / TPGD loop with every rank holding its slab of
chi, the gradient and the history; the inner products of an iteration travel in one batched all-reduce.

There is no analogue in the reference (single process).
"""
import contextlib
import ctypes
import os
import threading

import torch

from . import _native

_COMM_FN = _native.COMM_FN
COMM_SCRATCH = 64
_state = threading.local()


class TorchDistComm:
    """Communication through ``torch.distributed`` (backend nccl for CUDA tensors)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_to_all(self, recv, send):
        self.dist.all_to_all_single(recv, send, group=self.group)

    def all_reduce(self, t):
        self.dist.all_reduce(t, group=self.group)

    def all_reduce_max(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)

    def symmetric_buffer(self, numel, device):
        """``numel`` complex128 of symmetric memory (torch.distributed._symmetric_memory: the same allocation on every rank,
        mapped into every process of the node over NVLink).  Returns (tensor, [address of rank r's copy in THIS process]),
        or None when the box cannot do it (no peer access; PAD_SLAB_PEER=0) -- the caller then uses the staged exchange."""
        if os.environ.get('PAD_SLAB_PEER', '1') == '0' or self.world > 8:
            return None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty(numel, dtype=torch.complex128, device=device)
            hdl = symm_mem.rendezvous(t, self.group if self.group is not None else self.dist.group.WORLD)
            t.zero_()
            self._symm = getattr(self, '_symm', []) + [(t, hdl)]
            self._barrier_hdl = hdl
            hdl.barrier()
            return t, [int(p) for p in hdl.buffer_ptrs]
        except Exception as e:      # noqa: BLE001 -- an optimisation, never a failure
            self.symm_error = repr(e)
            return None

    def barrier(self, channel=0):
        """Stream-ordered barrier over the ranks (signal pads of the symmetric allocation; a 1-element all-reduce otherwise)."""
        hdl = getattr(self, '_barrier_hdl', None)
        if hdl is not None:
            hdl.barrier(channel=channel)
        else:
            if not hasattr(self, '_one'):
                self._one = torch.zeros(1, device='cuda')
            self.dist.all_reduce(self._one, group=self.group)


class SingleComm:
    """world = 1: the exchange is a copy (exercises the slab code path on one GPU)."""
    rank, world = 0, 1

    def all_to_all(self, recv, send):
        recv.copy_(send)

    def all_reduce(self, t):
        pass

    def all_reduce_max(self, t):
        pass

    def symmetric_buffer(self, numel, device):
        t = torch.zeros(numel, dtype=torch.complex128, device=device)
        return t, [t.data_ptr()]

    def barrier(self, channel=0):
        pass


class ThreadComm:
    """`world` ranks as threads of ONE process sharing one GPU -- a test double for NCCL.

    Every rank runs its evaluation in its own thread and its own CUDA stream; the collectives meet at a
    ``threading.Barrier`` after synchronising the streams.  Used by tests/test_gpu_parallel.py to check
    world = 2, 4 slab results on the single-GPU test box."""

    class Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.send = [None] * world
            self.vals = [None] * world

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world = shared, rank, shared.world

    def all_to_all(self, recv, send):
        sh = self.shared
        torch.cuda.current_stream(send.device).synchronize()
        sh.send[self.rank] = send
        sh.barrier.wait()
        n = send.numel() // self.world
        for r in range(self.world):
            recv[r * n:(r + 1) * n].copy_(sh.send[r][self.rank * n:(self.rank + 1) * n])
        torch.cuda.current_stream(send.device).synchronize()
        sh.barrier.wait()

    def all_reduce(self, t, op=torch.add):
        sh = self.shared
        torch.cuda.current_stream(t.device).synchronize()
        sh.vals[self.rank] = t.clone()
        sh.barrier.wait()
        total = sh.vals[0].clone()
        for r in range(1, self.world):      # fixed order: every rank gets bit-identical results
            total = op(total, sh.vals[r])
        torch.cuda.current_stream(t.device).synchronize()
        sh.barrier.wait()
        t.copy_(total)

    def all_reduce_max(self, t):
        self.all_reduce(t, op=torch.maximum)

    def symmetric_buffer(self, numel, device):
        """The ranks are threads of one process on one GPU: every rank's buffer is directly addressable -- the peer-pointer
        kernels run exactly as over NVLink."""
        if os.environ.get('PAD_SLAB_PEER', '1') == '0':
            return None
        sh = self.shared
        t = torch.zeros(numel, dtype=torch.complex128, device=device)
        torch.cuda.current_stream(device).synchronize()
        sh.vals[self.rank] = t
        sh.barrier.wait()
        ptrs = [sh.vals[r].data_ptr() for r in range(self.world)]
        sh.barrier.wait()
        return t, ptrs

    def barrier(self, channel=0):
        torch.cuda.current_stream().synchronize()
        self.shared.barrier.wait()


class SlabPlan(_native.Plan):
    """``pad_plan`` for this rank's slab of a global grid."""

    def __init__(self, box_host, global_shape, device_index, comm, overlap=True, fast=True):
        self.lib = _native.load_library()
        self.comm = comm
        self.global_shape = tuple(int(s) for s in global_shape)
        n0, n1, n2 = self.global_shape
        if n0 % comm.world or n1 % comm.world:
            raise ValueError(f'slab decomposition needs n0 = {n0} and n1 = {n1} to be multiples of the world size {comm.world}')
        self.shape = (n0 // comm.world, n1, n2)                  # the local real-space slab
        self.device_index = device_index
        dev = torch.device('cuda', device_index)
        nk_loc = n0 * (n1 // comm.world) * (n2 // 2 + 1)
        # cuFFT slab path over peer memory: both receive buffers in ONE symmetric allocation, so that the pack kernels of
        # the other ranks can store straight into them (pad_plan_set_slab_peer_recv); plain tensors + all-to-all otherwise
        # (measured on 8 x B200, revHC 512^3 / PBE 1024^3: 2 GPUs 68.4 -> 59.8 ms with the peer form, but 8 GPUs 20.8 -> 22.4 /
        #  32.7 -> 48.7 ms -- two barriers + a copy kernel per exchange, each a host callback, against ONE pipelined NCCL
        #  all-to-all; so by default the peer form is used for world <= 2 only.  PAD_SLAB_PEER_RECV=1 / 0 overrides.)
        self.recv_sym = None
        want = os.environ.get('PAD_SLAB_PEER_RECV', '')
        use_peer_recv = (want == '1') or (want != '0' and comm.world <= 2) or isinstance(comm, ThreadComm)
        if use_peer_recv and overlap and comm.world > 1 and hasattr(comm, 'symmetric_buffer'):
            self.recv_sym = comm.symmetric_buffer(2 * nk_loc, dev)
        self.send = torch.empty(nk_loc, dtype=torch.complex128, device=dev)
        self.recv = self.recv_sym[0][:nk_loc] if self.recv_sym is not None else torch.empty(nk_loc, dtype=torch.complex128, device=dev)
        self.scratch = torch.zeros(COMM_SCRATCH, dtype=torch.double, device=dev)
        self.error = None
        # second exchange pair: lets the library overlap the all-to-all of one field with the FFTs of the next
        self.overlap = overlap and comm.world > 1
        if self.overlap:
            self.send2 = torch.empty(nk_loc, dtype=torch.complex128, device=dev)
            self.recv2 = (self.recv_sym[0][nk_loc:] if self.recv_sym is not None
                          else torch.empty(nk_loc, dtype=torch.complex128, device=dev))
        # fused FFT pipeline on the slabs (csrc/fftz.cu: own z / y / x passes, the y pass stores its rows blocked by
        # destination rank so the exchange needs no pack kernel): four spectrum fields + two exchange stagings in the
        # padded layout, registered with the library; the callback exchanges them by index
        # Preferred form (NVLink / NVSwitch): ONE symmetric allocation per rank holding the four fields in the local and in
        # the transposed layout, mapped into every process -- the y and x passes then store their results straight into the
        # owner ranks' buffers and the only collective left is a barrier.
        self.fast, self.peer = [], None
        if fast and all(n in (64, 128, 256, 512) for n in (n0, n1)) and n2 in (128, 256, 512):
            nzp = n2 // 2 + 8
            nfast = (n0 // comm.world) * n1 * nzp
            sym = comm.symmetric_buffer(8 * nfast, dev) if hasattr(comm, 'symmetric_buffer') else None
            if sym is not None:
                self.peer = sym
            else:
                self.fast = [torch.zeros(nfast, dtype=torch.complex128, device=dev) for _ in range(6)]
        streams = {}

        def callback(_user, op, count, stream):
            # the library names the stream the collective has to be ordered with (its own communication stream for
            # the pipelined exchanges, the caller's stream otherwise)
            try:
                key = int(stream or 0)
                cur = torch.cuda.current_stream(dev)
                if key == cur.cuda_stream:
                    ext = cur                   # (wrapping the caller's own stream as an ExternalStream breaks the ordering
                elif key == 0:                  #  of torch's NCCL collectives when the handle is the NULL stream)
                    ext = torch.cuda.default_stream(dev)
                else:
                    ext = streams.get(key)
                    if ext is None:
                        ext = streams[key] = torch.cuda.ExternalStream(key, device=dev)
                with torch.cuda.stream(ext):
                    if op == 0:
                        self.comm.all_to_all(self.recv, self.send)
                    elif op == 3:
                        self.comm.all_to_all(self.recv2, self.send2)
                    elif op == 1:
                        self.comm.all_reduce(self.scratch[:count])
                    elif op == 2:
                        self.comm.all_reduce_max(self.scratch[:count])
                    elif op == 4:
                        self.comm.barrier()
                    elif op == 5:
                        self.comm.barrier(1)
                    elif op >= 16:
                        dst, src = divmod(op - 16, 8)
                        self.comm.all_to_all(self.fast[dst], self.fast[src])
                    else:
                        raise ValueError(f'unknown communication op {op}')
                return 0
            except BaseException as e:      # noqa: BLE001 -- must not propagate through the C frame
                self.error = e
                return 1

        self._callback = _COMM_FN(callback)                      # keep alive as long as the plan
        self.handle = ctypes.c_void_p()
        self.box = None
        box_arr = (ctypes.c_double * 9)(*box_host)
        shp = (ctypes.c_int * 3)(*self.global_shape)
        _native.check(self.lib.pad_plan_create_slab(ctypes.byref(self.handle), box_arr, shp, device_index, comm.rank,
                                                    comm.world, _native.ptr(self.send), _native.ptr(self.recv),
                                                    _native.ptr(self.scratch), self._callback, None))
        if self.overlap:
            _native.check(self.lib.pad_plan_set_overlap_buffers(self.handle, _native.ptr(self.send2), _native.ptr(self.recv2)))
        if self.recv_sym is not None:
            ptrs = self.recv_sym[1]
            a1 = (ctypes.c_void_p * comm.world)(*ptrs)
            a2 = (ctypes.c_void_p * comm.world)(*[q + 16 * nk_loc for q in ptrs])
            _native.check(self.lib.pad_plan_set_slab_peer_recv(self.handle, a1, a2, comm.world))
        if self.fast:
            arr = (ctypes.c_void_p * 6)(*[t.data_ptr() for t in self.fast])
            _native.check(self.lib.pad_plan_set_slab_fast_buffers(self.handle, arr))
        if self.peer is not None:
            arr = (ctypes.c_void_p * comm.world)(*self.peer[1])
            _native.check(self.lib.pad_plan_set_slab_peer_buffers(self.handle, arr, comm.world))
        self.box = tuple(box_host)
        self._set_geometry()

    def _set_geometry(self):
        super()._set_geometry()
        n0, n1, n2 = self.global_shape
        self.npts = n0 * n1 * n2                                 # dV is the GLOBAL volume element
        self.dV = self.vol / self.npts


class _SlabContext:
    def __init__(self, global_shape, comm, overlap=True):
        self.global_shape = tuple(int(s) for s in global_shape)
        self.comm = comm
        self.overlap = overlap
        self.plans = {}

    @property
    def local_shape(self):
        n0, n1, n2 = self.global_shape
        return (n0 // self.comm.world, n1, n2)

    def plan_for(self, box_vecs, den):
        if tuple(den.shape) != self.local_shape:
            raise ValueError(f'inside parallel.slab(global_shape={self.global_shape}) rank {self.comm.rank} of '
                             f'{self.comm.world} expects a density slab of shape {self.local_shape}, got {tuple(den.shape)}')
        dev = den.device.index if den.device.index is not None else torch.cuda.current_device()
        host = _native.box_to_host(box_vecs)
        plan = self.plans.get(dev)
        if plan is None:
            plan = SlabPlan(host, self.global_shape, dev, self.comm, self.overlap)
            self.plans[dev] = plan
        else:
            plan.set_box(host)
        return plan

    def close(self):
        for p in self.plans.values():
            p.close()
        self.plans.clear()


def current():
    """The active slab context of this thread, or None."""
    return getattr(_state, 'ctx', None)


@contextlib.contextmanager
def slab(global_shape, comm=None, group=None, overlap=None):
    """Evaluate native functionals on this rank's slab of a ``global_shape`` grid (see the module docstring).
    ``overlap``: allocate a second exchange buffer pair so that batches of transforms are software-pipelined
    (all-to-all of one field on a communication stream while the neighbouring fields' FFTs run)."""
    if comm is None:
        import torch.distributed as dist
        comm = TorchDistComm(group) if dist.is_available() and dist.is_initialized() else SingleComm()
    if overlap is None:
        overlap = os.environ.get('PAD_SLAB_OVERLAP', '1') != '0'
    ctx = _SlabContext(global_shape, comm, overlap)
    prev = current()
    _state.ctx = ctx
    try:
        yield ctx
    finally:
        _state.ctx = prev
        ctx.close()


def slab_bounds(n0, rank, world):
    """[lo, hi) planes of axis 0 owned by ``rank``."""
    if n0 % world:
        raise ValueError(f'n0 = {n0} is not a multiple of the world size {world}')
    m = n0 // world
    return rank * m, (rank + 1) * m


def local_slab(field_global, comm=None):
    """This rank's contiguous slab (a copy) of a full (n0, n1, n2) field."""
    ctx = current()
    comm = comm or (ctx.comm if ctx else None)
    if comm is None:
        raise RuntimeError('local_slab needs an active parallel.slab(...) context or an explicit comm')
    lo, hi = slab_bounds(field_global.shape[0], comm.rank, comm.world)
    return field_global[lo:hi].contiguous()


def optimize_density(box_vecs, den_local, v_ext_local, terms, n_elec, ntol=1e-7, n_conv_cond_count=3, n_method='LBFGS',
                     n_step_size=0.1, n_maxiter=1000, conv_target='dE'):
    """``System.optimize_density`` (system.py:774-908) for one grid spread over the ranks: call inside
    ``with parallel.slab(global_shape):`` with this rank's slabs of the starting density and of the ionic
    potential.  ``den_local`` is overwritten with the optimised density; returns (result dict, trace) --
    identical on every rank.  ``terms`` is the reference-style list of native functionals."""
    from . import _density_opt
    if current() is None:
        raise RuntimeError('parallel.optimize_density must be called inside a parallel.slab(...) context')
    T = _density_opt.describe_terms(terms, den_local.device)
    if T is None:
        raise NotImplementedError('parallel.optimize_density needs every term to be a native functional')
    return _density_opt.run(box_vecs, den_local, v_ext_local, T, n_elec, ntol, n_conv_cond_count, n_method,
                            n_step_size, n_maxiter, conv_target)


def eos_fit(make_system, f=0.05, N=9, eos='bm', group=None, **den_opt_kwargs):
    """Energy-volume scan + equation-of-state fit (``System.eos_fit``, system.py:568-621) with the N volumes spread
    over the ranks, one independent System per GPU -- the "independent systems" mode of SURVEY.md section 8e (no
    collective in the loop, one gather of N (volume, energy) pairs at the end).

    ``make_system()`` builds the System at its reference volume on this rank's device.  The reference carries the
    density from one volume to the next as a warm start; here every rank starts its first volume from the
    reference-volume density (rescaled by ``set_lattice``) and warm-starts along its own sub-sequence.  Returns
    (params, err) in the reference's order -- K0 [GPa], K0', E0 [eV], V0 [A^3] -- on every rank."""
    import numpy as np
    import torch.distributed as dist
    from .elastic_tools import fit_eos
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    opts = {'ntol': 1e-10, 'n_conv_cond_count': 3, 'n_method': 'LBFGS', 'n_step_size': 0.1, 'n_maxiter': 1000,
            'conv_target': 'dE', 'n_verbose': False, 'from_uniform': False}
    opts.update(den_opt_kwargs)
    s = make_system()
    v0 = s.volume('a3')
    shape_vecs = s.lattice_vectors('a') / v0 ** (1 / 3)
    volumes = v0 * np.linspace(1 - f, 1 + f, N)
    mine = []
    for i in range(rank, N, world):
        s.set_lattice(volumes[i] ** (1 / 3) * shape_vecs, units='a')
        s.optimize_density(**opts)
        mine.append((i, s.volume('a3') / s.ion_count(), s.energy('eV') / s.ion_count()))
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
        mine = [row for part in gathered for row in part]
    mine.sort()
    vols, enes = [r[1] for r in mine], [r[2] for r in mine]
    params, err = fit_eos(vols, enes, eos, False)
    to_gpa = s.GPa_per_atomic / (s.eV_per_Ha / s.A_per_b ** 3)
    params[0] *= to_gpa
    err[0] *= to_gpa
    return params, err
