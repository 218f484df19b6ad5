"""Host-resident batches: evaluate E[n] and dE/dn for a stream of densities that live in pinned host memory.

A single evaluation from host memory is PCIe-bound (a 256^3 density is 134 MB each way).  When many
independent densities have to be evaluated (scans, training-set generation, the benchmark's end-to-end leg)
the copies of neighbouring evaluations can overlap the kernels: three CUDA streams (host->device, compute,
device->host), `depth` device buffers, events between them.  The evaluation itself is the ordinary public
call `functional(box_vecs, den)` + `torch.autograd.grad`, issued on the compute stream.
"""
import torch


class HostPipeline:
    def __init__(self, functional, box_vecs, shape, device, depth=2):
        self.functional = functional
        self.box = box_vecs.to(device)
        self.device = torch.device(device)
        self.depth = depth
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self.den = [torch.empty(shape, dtype=torch.double, device=self.device) for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_cmp = [torch.cuda.Event() for _ in range(depth)]
        self.ev_free_in = [torch.cuda.Event() for _ in range(depth)]
        self.pending = [None] * depth

    def run(self, inputs, out_v, out_e):
        """inputs[i]: pinned host density; out_v[i]: pinned host potential buffer; out_e: pinned (len(inputs),)
        energies.  Returns after everything has been ENQUEUED; the caller's current stream waits for the results."""
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        used = [False] * self.depth
        for i, den_h in enumerate(inputs):
            k = i % self.depth
            with torch.cuda.stream(self.s_in):
                if used[k]:
                    self.s_in.wait_event(self.ev_free_in[k])          # evaluation i - depth has consumed den[k]
                self.den[k].copy_(den_h, non_blocking=True)
                self.ev_in[k].record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(self.ev_in[k])
                d = self.den[k].requires_grad_(True)
                E = self.functional(self.box, d)
                (g,) = torch.autograd.grad(E, d)
                self.den[k].requires_grad_(False)
                self.ev_free_in[k].record(self.s_cmp)
                self.ev_cmp[k].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_cmp[k])
                out_v[i].copy_(g, non_blocking=True)
                out_e[i:i + 1].copy_(E.detach().reshape(1), non_blocking=True)
                g.record_stream(self.s_out)
                E.record_stream(self.s_out)
            used[k] = True
        for s in (self.s_in, self.s_cmp, self.s_out):
            cur.wait_stream(s)
