// Software-pipelined (z, y) passes: ONE persistent kernel runs the z pass (contiguous-axis real <-> half-complex
// FFT with the real-space work fused in) and the y pass of every x-plane, with the plane handed from one pass to
// the other through the L2 instead of through HBM.
//
// Why not a thread-block cluster with the plane in distributed shared memory: a (y, z) half-spectrum plane of ONE
// field is n1 * (n2/2 + 1) * 16 B = 528 KB at 256^3; the inverse side needs the planes of all four fields of a
// point together (2.1 MB, more than a portable 8-CTA cluster holds), and DSMEM moves ~20 B/clk/SM -- about the HBM
// rate.  The L2 holds tens of planes.  So the passes stay separate code, but they are ITEMS of one kernel:
//
//   stage 0 item (plane p, sub s)   no dependencies, handed out by a ticket counter in plane order
//   stage k item (plane p, sub s)   becomes claimable when ALL stage k-1 items of plane p are complete
//
// A CTA that looks for work runs the highest stage for which it holds a runnable ticket ("consumer first"): a plane
// is finished soon after it can be, so the data in flight between two stages is the few planes being worked on
// (tens of MB, L2 resident) and the z items (fp64-issue bound) share every SM with y items (memory bound).
// A CTA never waits while it holds a runnable ticket, so the scheme cannot deadlock whatever the number of resident
// CTAs; a bounded spin (only when stage 0 is exhausted and the last planes are still in flight) raises
// PipeCtl::error instead of hanging.
//
// Memory ordering: producer threads write with plain stores, __syncthreads(), thread 0 fences (gpu scope) and
// bumps the plane's counter; the consumer's thread 0 reads the counters with ld.acquire.gpu, fences, __syncthreads(),
// and every load of pipelined data goes through the L2 (ld.global.cg / cp.async.cg), never the incoherent L1.
#pragma once
#include "common.cuh"
#include "fft_core.cuh"
#include "fft_strided.cuh"

#define PIPE_MAX_PLANES 1024

struct PipeCtl {                      // zero-initialised once; every kernel leaves it zeroed (last CTA out resets it)
    unsigned int next[4];             // ticket counters of the stages
    unsigned int exited, error, pad[2];
    unsigned int done[2][PIPE_MAX_PLANES];      // completed items of stage 0 / 1 per plane
};
struct PipeShape {
    int nstage, nplanes;
    int items[3];                     // items per plane of each stage
};
struct PipeItem {
    int stage, plane, sub;            // stage < 0: nothing left
};
// Scheduling.  Every stage hands out its items (plane-major) through ONE atomicAdd ticket counter -- pipelined at the
// L2, no compare-and-swap loop (a CAS loop serialises at one claim per L2 round trip: measured 0.29 us per item over
// the whole GPU).  A CTA always HOLDS one ticket per stage; it runs the ticket of the highest stage whose plane is
// complete in the stage below (stage-0 tickets always are), and replaces that ticket at once -- the atomicAdd that
// fetches the replacement returns while the item is being worked on.  Tickets past the end are void; a CTA leaves
// when all its tickets are void.  No CTA ever waits while it holds a runnable ticket and stage-0 tickets need
// nothing, so by induction over the stages every item gets run whatever the number of resident CTAs.
struct PipeSched {                    // thread-0 registers
    unsigned t0, t1, t2;
};

#ifndef PAD_HOST_EMU      // (tests/host_emu: the scheduler runs on host threads with these two replaced by std::atomic loads)
__device__ __forceinline__ unsigned pipe_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned pipe_ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif

template <int NSTAGE>
__device__ __forceinline__ void pipe_begin(PipeCtl* c, PipeSched& ps) {
    ps.t0 = ps.t1 = ps.t2 = 0xffffffffu;
    if (threadIdx.x == 0) {
        ps.t0 = atomicAdd(&c->next[0], 1u);
        if (NSTAGE > 1) ps.t1 = atomicAdd(&c->next[1], 1u);
        if (NSTAGE > 2) ps.t2 = atomicAdd(&c->next[2], 1u);
    }
}

// sm: 4 ints of shared memory
template <int NSTAGE>
__device__ __forceinline__ PipeItem pipe_next(PipeCtl* c, const PipeShape& sh, int* sm, PipeSched& ps) {
    if (threadIdx.x == 0) {
        int stage = -1, plane = 0, sub = 0;
        const unsigned tot0 = (unsigned)(sh.nplanes * sh.items[0]);
        const unsigned tot1 = NSTAGE > 1 ? (unsigned)(sh.nplanes * sh.items[1]) : 0u;
        const unsigned tot2 = NSTAGE > 2 ? (unsigned)(sh.nplanes * sh.items[2]) : 0u;
        unsigned spins = 0;
        for (;;) {
            const bool v1 = NSTAGE > 1 && ps.t1 < tot1, v2 = NSTAGE > 2 && ps.t2 < tot2;
            const unsigned p1 = v1 ? ps.t1 / (unsigned)sh.items[1] : 0u, p2 = v2 ? ps.t2 / (unsigned)sh.items[2] : 0u;
            // both availability words are requested before either is looked at: one L2 round trip
            unsigned d1 = 0u, d2 = 0u;
            if (v2) d2 = pipe_ld_acquire(&c->done[1][p2]);
            if (v1) d1 = pipe_ld_acquire(&c->done[0][p1]);
            if (v2 && d2 >= (unsigned)sh.items[1]) {
                stage = 2; plane = (int)p2; sub = (int)(ps.t2 - p2 * (unsigned)sh.items[2]);
                ps.t2 = atomicAdd(&c->next[2], 1u);
                break;
            }
            if (v1 && d1 >= (unsigned)sh.items[0]) {
                stage = 1; plane = (int)p1; sub = (int)(ps.t1 - p1 * (unsigned)sh.items[1]);
                ps.t1 = atomicAdd(&c->next[1], 1u);
                break;
            }
            if (ps.t0 < tot0) {
                stage = 0; plane = (int)(ps.t0 / (unsigned)sh.items[0]); sub = (int)(ps.t0 - (unsigned)plane * (unsigned)sh.items[0]);
                ps.t0 = atomicAdd(&c->next[0], 1u);
                break;
            }
            if (!v1 && !v2) break;                                        // every ticket held is void: done
            if (pipe_ld_relaxed(&c->error) != 0u) break;
            __nanosleep(100);
            if (++spins > (1u << 24)) { atomicExch(&c->error, 1u); break; }
        }
        __threadfence();
        sm[0] = stage; sm[1] = plane; sm[2] = sub;
    }
    __syncthreads();
    PipeItem it{sm[0], sm[1], sm[2]};
    return it;
}

// all threads; ends the item: its global writes are published, the plane's counter of the stage goes up by one
__device__ __forceinline__ void pipe_done(PipeCtl* c, const PipeShape& sh, const PipeItem& it) {
    __syncthreads();
    if (threadIdx.x == 0 && it.stage < sh.nstage - 1) {
        __threadfence();
        atomicAdd(&c->done[it.stage][it.plane], 1u);
    }
}

// the last CTA to leave resets the control block for the next launch
__device__ __forceinline__ void pipe_exit(PipeCtl* c, const PipeShape& sh, int* sm) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        sm[3] = (atomicAdd(&c->exited, 1u) + 1u == gridDim.x) ? 1 : 0;
    }
    __syncthreads();
    if (sm[3]) {
        for (int i = threadIdx.x; i < sh.nplanes; i += blockDim.x) { c->done[0][i] = 0u; c->done[1][i] = 0u; }
        if (threadIdx.x == 0) {
            c->next[0] = c->next[1] = c->next[2] = c->next[3] = 0u;
            __threadfence();
            c->exited = 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
//  y items: tiles of ONE x-plane.  A plane of nf fields has nf * chunks regular tiles (field f, 8 consecutive
//  z columns) and one tile that holds the Nyquist column (z = nzh - 1) of field c in its column c.
// ------------------------------------------------------------------------------------------------
struct ZYGeom {
    int n1;               // lines per plane
    int nzp, nzh;         // padded / live half-spectrum row length
    int chunks;           // nzh / 8 (nzh = 8 m + 1 for every supported n2)
    int lpi;              // lines per z item
    int tpi;              // tiles per y item
    long long plane;      // complex elements per spectrum plane: n1 * nzp
};

struct YTile {
    bool live;
    cd* base;
};
__device__ __forceinline__ YTile ytile_locate(const SPassFields& F, int nf, const ZYGeom& g, int plane, int w, int ntiles, int c) {
    YTile a;
    const int regular = nf * g.chunks;
    int f = 0, z = 0;
    bool live = false;
    if (w < regular) {
        f = w / g.chunks;
        z = (w - f * g.chunks) * 8 + c;
        live = true;
    } else if (w < ntiles) {
        f = c < nf ? c : 0;
        z = g.nzh - 1;
        live = c < nf;
    }
    a.live = live;
    a.base = F.f[f] + (long long)plane * g.plane + z;
    return a;
}

__device__ __forceinline__ cd ldcg_cd(const cd* p) {
    const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
    return cd{v.x, v.y};
}
__device__ __forceinline__ void stcs_cd(cd* p, cd v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }

// in-place FFT along y of the tiles [w_begin, w_end) of a plane; loads straight into registers (for kernels that
// keep 16 warps per SM resident).  STREAM: the results are not read again soon (evict-first stores).
template <int L, int DIR, bool STREAM>
__device__ __forceinline__ void ytile_item_direct(const SPassFields& F, int nf, const ZYGeom& g, int plane, int w_begin, int w_end,
                                                  int ntiles, cd* Sarea, const cd* tw) {
    using P = SPass<L>;
    const int tile_in_cta = threadIdx.x / P::TILE_THREADS;
    const int tid = threadIdx.x % P::TILE_THREADS;
    const int t = tid / P::ZC, c = tid % P::ZC;
    cd* S = Sarea + (size_t)tile_in_cta * P::TILE_CD;
    for (int w0 = w_begin; w0 < w_end; w0 += P::TPC) {
        const int w = w0 + tile_in_cta;
        const YTile a = ytile_locate(F, nf, g, plane, w < w_end ? w : ntiles, ntiles, c);
        cd v[P::EPT];
#pragma unroll
        for (int j = 0; j < P::EPT; ++j) v[j] = a.live ? ldcg_cd(a.base + (size_t)(t + P::TPL * j) * g.nzp) : cd{0.0, 0.0};
        tile_fft<L, DIR>(v, S, t, c, tw);
        if (a.live) {
#pragma unroll
            for (int s = 0; s < P::EPT; ++s) {
                cd* dst = a.base + (size_t)spass_out_index<L>(t, s) * g.nzp;
                if (STREAM) stcs_cd(dst, v[s]);
                else *dst = v[s];
            }
        }
    }
}

// same, the tiles land by cp.async in two alternating buffers (for kernels with 8 warps per SM: the copy of the
// tile after next flies while this one is transformed).  Sarea: 2 * TPC * TILE_CD complex.
template <int L, int DIR, bool STREAM>
__device__ __forceinline__ void ytile_item_async(const SPassFields& F, int nf, const ZYGeom& g, int plane, int w_begin, int w_end,
                                                 int ntiles, cd* Sarea, const cd* tw) {
    using P = SPass<L>;
    const int tile_in_cta = threadIdx.x / P::TILE_THREADS;
    const int tid = threadIdx.x % P::TILE_THREADS;
    const int t = tid / P::ZC, c = tid % P::ZC;
    constexpr int ROWSTEP = P::TPL * P::ZC;
    const int own = t * P::ZC + c;
    const int nt = (w_end - w_begin + P::TPC - 1) / P::TPC;
    auto buf = [&](int q) { return Sarea + (size_t)(q * P::TPC + tile_in_cta) * P::TILE_CD; };
    auto locate = [&](int it) {
        const int w = w_begin + it * P::TPC + tile_in_cta;
        return ytile_locate(F, nf, g, plane, (it < nt && w < w_end) ? w : ntiles, ntiles, c);
    };
    auto issue = [&](int it, int q) {
        const YTile a = locate(it);
        if (a.live) {
            cd* dst = buf(q) + own;
#pragma unroll
            for (int j = 0; j < P::EPT; ++j) cp_async16(dst + j * ROWSTEP, a.base + (size_t)(t + P::TPL * j) * g.nzp);
        }
        cp_async_commit();
    };
    issue(0, 0);
    issue(1, 1);
    for (int it = 0; it < nt; ++it) {
        const int q = it & 1;
        cp_async_wait<1>();
        const YTile a = locate(it);
        cd* B = buf(q);
        cd v[P::EPT];
#pragma unroll
        for (int j = 0; j < P::EPT; ++j) v[j] = a.live ? B[own + j * ROWSTEP] : cd{0.0, 0.0};
        tile_fft<L, DIR>(v, B, t, c, tw);
        issue(it + 2, q);
        if (a.live) {
#pragma unroll
            for (int s = 0; s < P::EPT; ++s) {
                cd* dst = a.base + (size_t)spass_out_index<L>(t, s) * g.nzp;
                if (STREAM) stcs_cd(dst, v[s]);
                else *dst = v[s];
            }
        }
    }
    cp_async_wait<0>();
}
