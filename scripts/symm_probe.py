"""Probe: torch symmetric memory (peer pointers over NVLink) on this box.  torchrun --nproc-per-node 2 scripts/symm_probe.py"""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n = 1 << 24
t = symm_mem.empty(n, dtype=torch.complex128, device=dev)
h = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, 'ptrs', [hex(p) for p in h.buffer_ptrs], 'multicast', h.has_multicast_support, flush=True)
t.fill_(rank + 1)
h.barrier()
peer = (rank + 1) % world
remote = h.get_buffer(peer, (n,), torch.complex128)
src = torch.full((n,), 10.0 * (rank + 1), dtype=torch.complex128, device=dev)
torch.cuda.synchronize()
for _ in range(3):
    remote[: n // 2].copy_(src[: n // 2])          # push over NVLink
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    remote[: n // 2].copy_(src[: n // 2])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(rank, 'push GB/s', (n // 2) * 16 / dt / 1e9, flush=True)
h.barrier()
torch.cuda.synchronize()
print(rank, 'local first', t[0].item(), 'local last', t[-1].item(), flush=True)
dist.destroy_process_group()
