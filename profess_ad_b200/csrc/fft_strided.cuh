// Strided-axis (x or y) complex FFT passes over the padded half-spectrum layout spec[x][y][nzp],
// with the reciprocal-space multiply fused between the forward and the inverse x transform.
//
// Tile = L points along the transformed axis (L = 64 ... 512) x 8 consecutive z columns (128-byte row segments).
// TPL threads share a line; thread (t, c) owns the points n = t + TPL j of column c, so one warp-wide
// 16-byte load touches 4 rows x 128 contiguous bytes.  The L-point FFT is two register radix stages
// (fft_reg) around one shared-memory exchange; the result comes out in natural order across
// (thread, slot), so it is stored (y pass) or multiplied and fed straight back into the inverse
// transform (x pass) without another trip through shared or global memory.
//
//   y pass :  spec <- FFT_y(spec)                                      1 read + 1 write per field
//   x pass :  spec <- IFFT_x( Mix_k( FFT_x(spec_0..NF-1) ) )           1 read + 1 write per field + kernel arrays
//
// The x pass transforms NF coupled fields of one tile back to back; spectra that have to wait for
// their partners are parked in thread-private shared-memory slots (no synchronisation).
#pragma once
#include "common.cuh"
#include "fft_core.cuh"

// WIDE (L = 512 only, used by the fused x pass): 32 threads per line with 16 points each instead of 16 threads with 32 points --
// twice the warps per tile and half the registers per thread (a 32-point register FFT plus its temporaries does not fit
// next to the parked spectra: 1 CTA of 4 warps per SM, measured 1 TB/s).  The line is split 512 = 16 (j, registers) x 32 (t):
// a 16-point register FFT over j, the twiddle, one shared-memory exchange, then the 32-point FFT over t of every k_j is
// shared by the two threads (k_j, h): thread h computes the outputs k_t = 2 m + h as a 16-point register FFT of
// (y[t'] +- y[t' + 16]) W32^{t' h}.  The result is distributed over the threads exactly like the input:
// thread t, slot r holds X[t + 32 fft_nat<16>(r)].
#ifndef PAD_Y256_CTAS
#define PAD_Y256_CTAS 4
#endif
template <int L, bool WIDE = false>
struct SPass {
    static_assert(!WIDE || L == 512 || L == 256, "strided pass: the wide layout exists for L = 512 and 256");
    static constexpr int TPL = WIDE ? 32 : ((L >= 256) ? 16 : 8);      // threads per line
    static constexpr int EPT = L / TPL;                  // points per thread: 32 (L = 512), 16 (L = 256, 128, wide 512), 8 (L = 64)
    static constexpr int G = WIDE ? 1 : EPT / TPL;       // second-stage FFTs per thread: 2, 1, 2, 1
    static constexpr int ZC = 8;                         // z columns per tile
    static constexpr int TILE_THREADS = TPL * ZC;        // 128, 64, 64; wide: 256
    static constexpr int THREADS = WIDE ? 256 : 128;
    static constexpr int TPC = THREADS / TILE_THREADS;   // tiles per CTA
    static constexpr int TILE_CD = L * ZC;               // complex numbers per tile
    static constexpr int CTAS_PER_SM = WIDE ? (L == 512 ? 2 : 4) : ((L == 512) ? 2 : (L == 256 ? PAD_Y256_CTAS : 4));   // plain pass: 32 complex points per thread need > 128 registers
    static_assert(L == 512 || L == 256 || L == 128 || L == 64, "strided pass: L must be 64, 128, 256 or 512");
};

// natural index along the line held by register slot s after tile_fft
template <int L, bool WIDE = false>
__device__ __forceinline__ int spass_out_index(int t, int s) {
    using P = SPass<L, WIDE>;
    if constexpr (WIDE) return t + P::TPL * fft_nat<P::EPT>(s);
    else return (t + P::TPL * (s / P::TPL)) + P::EPT * fft_nat<P::TPL>(s % P::TPL);
}
// register slot (after tile_fft) that holds the input j of a following tile_fft: k = t + TPL j
template <int L, bool WIDE = false>
__device__ __forceinline__ constexpr int spass_slot_of_input(int j) {
    using P = SPass<L, WIDE>;
    if constexpr (WIDE) return fft_slot<P::EPT>(j);
    // k = t + TPL g + EPT k2 = t + TPL (g + G k2)  ->  j = g + G k2
    else return (j % P::G) * P::TPL + fft_slot<P::TPL>(j / P::G);
}

// In: v[j] = z[t + TPL j].  Out: v[s] = Z[spass_out_index(t, s)].  S: tile scratch of L * 8 cd.  tw[e] = exp(-2 pi i e / L).
template <int L, int DIR, bool WIDE = false>
__device__ __forceinline__ void tile_fft(cd* v, cd* S, int t, int c, const cd* __restrict__ tw) {
    using P = SPass<L, WIDE>;
    if constexpr (WIDE) {
        // L = EPT (registers, j) x 32 (threads, t): EPT = 16 (L = 512) or 8 (L = 256)
        constexpr int E = P::EPT;
        fft_reg<E, DIR>(v);
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const int kj = fft_nat<E>(r);
            cd a = v[r];
            if (kj != 0) a = cmul(a, tw_dir<DIR>(tw[t * kj]));
            S[(kj * 32 + t) * P::ZC + c] = a;
        }
        __syncthreads();
        if constexpr (E == 16) {
            // the 32-point FFT over t of every k_j is shared by two threads (k_j, h): outputs k_t = 2 m + h
            const int kj = t & 15, h = t >> 4;
#pragma unroll
            for (int t2 = 0; t2 < 16; ++t2) {
                const cd lo = S[(kj * 32 + t2) * P::ZC + c], hi = S[(kj * 32 + t2 + 16) * P::ZC + c];
                cd u = h ? lo - hi : lo + hi;
                if (t2 != 0) u = cmul(u, tw_dir<DIR>(tw[h * (L / 32) * t2]));      // W32^{t2 h}; h = 0: tw[0] = 1
                v[t2] = u;
            }
        } else {
            // ... by four threads (k_j, h): outputs k_t = 4 m + h as an 8-point FFT of (sum_q y[t' + 8 q] W4^{q h}) W32^{t' h}
            const int kj = t & 7, h = t >> 3;
            const bool odd = h & 1, neg = h & 2;
#pragma unroll
            for (int t2 = 0; t2 < 8; ++t2) {
                const cd y0 = S[(kj * 32 + t2) * P::ZC + c], y1 = S[(kj * 32 + t2 + 8) * P::ZC + c];
                const cd y2 = S[(kj * 32 + t2 + 16) * P::ZC + c], y3 = S[(kj * 32 + t2 + 24) * P::ZC + c];
                const cd A = odd ? y0 - y2 : y0 + y2;
                const cd Bq = y1 - y3, Bs = y1 + y3;
                const cd B = odd ? mul_i<DIR>(Bq) : Bs;
                cd u = neg ? A - B : A + B;
                if (t2 != 0) u = cmul(u, tw_dir<DIR>(tw[h * (L / 32) * t2]));
                v[t2] = u;
            }
        }
        fft_reg<E, DIR>(v);
        __syncthreads();
    } else {
        fft_reg<P::EPT, DIR>(v);
#pragma unroll
        for (int r = 0; r < P::EPT; ++r) {
            const int k1 = fft_nat<P::EPT>(r);
            cd a = v[r];
            if (k1 != 0) a = cmul(a, tw_dir<DIR>(tw[t * k1]));
            S[(k1 * P::TPL + t) * P::ZC + c] = a;
        }
        __syncthreads();
#pragma unroll
        for (int g = 0; g < P::G; ++g) {
            const int k1 = t + P::TPL * g;
            cd u[P::TPL];
#pragma unroll
            for (int t2 = 0; t2 < P::TPL; ++t2) u[t2] = S[(k1 * P::TPL + t2) * P::ZC + c];
            fft_reg<P::TPL, DIR>(u);
#pragma unroll
            for (int r = 0; r < P::TPL; ++r) v[g * P::TPL + r] = u[r];
        }
        __syncthreads();
    }
}

template <int L>
__device__ __forceinline__ void spass_load_twiddles(cd* tw) {
    for (int e = threadIdx.x; e < L; e += blockDim.x) {
        const double2 w = g_fft_tw[e * (FFT_TW_N / L)];
        tw[e] = cd{w.x, w.y};
    }
    __syncthreads();
}

struct SPassGeom {
    long long axis_stride;     // complex elements between consecutive points of a line
    long long outer_stride;    // complex elements between consecutive line groups (the non-transformed, non-z axis)
    int n_outer;               // extent of that axis
    int nzh;                   // live z columns (n2/2 + 1); columns >= nzh are padding and are not touched
};

// Tiles of one field.  A regular tile is one outer index x 8 consecutive z columns.  nzh = 8 m + 1 for every
// supported n2, so the last (Nyquist) column would cost a whole tile per outer index with 1 lane in 8 live;
// instead the Nyquist column of 8 consecutive outer indices is packed into one tile (16-byte row segments,
// but only 1 / nzh of the data).
__host__ __device__ inline bool spass_packed_tail(const SPassGeom& g) { return (g.nzh % 8) == 1; }
__host__ __device__ inline int spass_chunks(const SPassGeom& g) { return spass_packed_tail(g) ? g.nzh / 8 : (g.nzh + 7) / 8; }
__host__ __device__ inline long long spass_tiles(const SPassGeom& g) {
    return (long long)g.n_outer * spass_chunks(g) + (spass_packed_tail(g) ? (g.n_outer + 7) / 8 : 0);
}

struct TileAt {
    bool live;
    int o, z;
    long long off;             // element offset of (o, z) inside a field
};
// thread column c of tile w (w < spass_tiles, else dead)
__device__ __forceinline__ TileAt spass_locate(const SPassGeom& g, long long w, long long tiles, int c) {
    TileAt a;
    const int chunks = spass_chunks(g);
    const long long regular = (long long)g.n_outer * chunks;
    const bool in_range = w < tiles;
    const long long ww = in_range ? w : 0;
    if (ww < regular) {
        a.o = (int)(ww / chunks);
        a.z = (int)(ww - (long long)a.o * chunks) * 8 + c;
        a.live = in_range && a.z < g.nzh;
    } else {
        a.o = (int)(ww - regular) * 8 + c;
        a.z = g.nzh - 1;
        a.live = in_range && a.o < g.n_outer;
        if (!a.live) a.o = 0;
    }
    a.off = (long long)a.o * g.outer_stride + a.z;
    return a;
}

struct SPassFields {
    cd* f[4];
};

// ------------------------------------------------------------------------------------------------
//  plain pass: in-place FFT along the strided axis for nf fields
// ------------------------------------------------------------------------------------------------
template <int L, int DIR, bool WIDE = false>
__global__ void __launch_bounds__((SPass<L, WIDE>::THREADS), (SPass<L, WIDE>::CTAS_PER_SM)) spass_kernel(SPassFields fields, int nf, SPassGeom geo) {
    using P = SPass<L, WIDE>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* tw = reinterpret_cast<cd*>(smem_raw);
    spass_load_twiddles<L>(tw);
    const int tile_in_cta = threadIdx.x / P::TILE_THREADS;
    const int tid = threadIdx.x % P::TILE_THREADS;
    const int t = tid / P::ZC, c = tid % P::ZC;
    cd* S = tw + L + (size_t)tile_in_cta * P::TILE_CD;
    const long long per_field = spass_tiles(geo);
    const long long total = per_field * nf;
    for (long long w0 = (long long)blockIdx.x * P::TPC; w0 < total; w0 += (long long)gridDim.x * P::TPC) {
        const long long w = w0 + tile_in_cta;
        const int fi = w < total ? (int)(w / per_field) : 0;
        const TileAt a = spass_locate(geo, w < total ? w - (long long)fi * per_field : per_field, per_field, c);
        cd* base = fields.f[fi] + a.off;
        cd v[P::EPT];
#pragma unroll
        for (int j = 0; j < P::EPT; ++j) v[j] = a.live ? base[(long long)(t + P::TPL * j) * geo.axis_stride] : cd{0.0, 0.0};
        tile_fft<L, DIR, WIDE>(v, S, t, c, tw);
        if (a.live) {
#pragma unroll
            for (int s = 0; s < P::EPT; ++s) base[(long long)spass_out_index<L, WIDE>(t, s) * geo.axis_stride] = v[s];
        }
    }
}

// ------------------------------------------------------------------------------------------------
//  slab plans: the y pass next to the all-to-all.  The transposition (n0_loc, n1, nzp) <-> (n0, n1_loc, nzp) moves, for every
//  pair of ranks, the rows y of the partner's range of all local x-planes.  With the rows of a plane BLOCKED by owner rank,
//      blocked[r][x_loc][y_l][z],   y = r n1_loc + y_l,
//  the block sent to rank r is contiguous and what arrives from rank r is the x range of r in the transposed layout, so the
//  exchange needs no pack / unpack pass of its own: the forward y pass stores its result rows blocked (BLK = 1), the inverse
//  y pass loads its input rows blocked (BLK = 2); the other side of either is the natural local layout.  Out of place.
// ------------------------------------------------------------------------------------------------
struct SPassBlocked {
    int rows;                  // n1_loc: rows per rank block
    long long block_stride;    // complex elements between the blocks of consecutive ranks: n0_loc * n1_loc * nzp
    long long outer_stride;    // complex elements between consecutive x-planes inside a block: n1_loc * nzp
};

// BLK = 3: as BLK = 1, but the block of rank r is stored straight into rank r's transposed-layout buffer over NVLink (peer
// pointers, symmetric memory): peers.p[r] + my_rank * block_stride is where "the block that came from me" lives over there.
// The transfer then overlaps the transform tile by tile and there is neither a staging buffer nor a separate collective.
struct SPassPeers {
    cd* p[8];
    int my_rank;
};

template <int L, int DIR, int BLK>
__global__ void __launch_bounds__(128, SPass<L>::CTAS_PER_SM) spass_blocked_kernel(const cd* __restrict__ src, cd* __restrict__ dst,
                                                                                 SPassGeom geo, SPassBlocked bl, const __grid_constant__ SPassPeers peers) {
    using P = SPass<L>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* tw = reinterpret_cast<cd*>(smem_raw);
    spass_load_twiddles<L>(tw);
    const int tile_in_cta = threadIdx.x / P::TILE_THREADS;
    const int tid = threadIdx.x % P::TILE_THREADS;
    const int t = tid / P::ZC, c = tid % P::ZC;
    cd* S = tw + L + (size_t)tile_in_cta * P::TILE_CD;
    const long long total = spass_tiles(geo);
    for (long long w0 = (long long)blockIdx.x * P::TPC; w0 < total; w0 += (long long)gridDim.x * P::TPC) {
        const TileAt a = spass_locate(geo, w0 + tile_in_cta, total, c);
        const long long nat = a.off;                                              // o * outer_stride + z
        const long long blk = (long long)a.o * bl.outer_stride + a.z;
        auto blocked_row = [&](int y) { const int r = y / bl.rows; return (long long)r * bl.block_stride + (long long)(y - r * bl.rows) * geo.axis_stride; };
        cd v[P::EPT];
#pragma unroll
        for (int j = 0; j < P::EPT; ++j) {
            const int y = t + P::TPL * j;
            const long long off = BLK == 2 ? blk + blocked_row(y) : nat + (long long)y * geo.axis_stride;
            v[j] = a.live ? src[off] : cd{0.0, 0.0};
        }
        tile_fft<L, DIR>(v, S, t, c, tw);
        if (a.live) {
#pragma unroll
            for (int s = 0; s < P::EPT; ++s) {
                const int y = spass_out_index<L>(t, s);
                if constexpr (BLK == 3) {
                    const int r = y / bl.rows;
                    peers.p[r][(long long)peers.my_rank * bl.block_stride + blk + (long long)(y - r * bl.rows) * geo.axis_stride] = v[s];
                } else {
                    const long long off = BLK == 1 ? blk + blocked_row(y) : nat + (long long)y * geo.axis_stride;
                    dst[off] = v[s];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
//  asynchronous global -> shared copies (LDGSTS): each thread copies the elements it will read itself,
//  so a cp.async.wait_group is all the synchronisation the consumer needs
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* g) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(g)); }

// ------------------------------------------------------------------------------------------------
//  x pass with fused multiply:  NF coupled fields.  The transformed axis is axis 0, the outer axis is
//  axis 1.  Mix concept:
//      typename Mix::Coef;  Mix::kRing   multiplier slots kept in flight ahead of their use (8 for table reads, 1 for computed ones)
//      Line  line(kg, ky, z)                per-thread constants of the tile column (e.g. a KLine for |k|^2 along kx)
//      Coef  fetch(line, kx, pidx, live)    multiplier data of one k-point (pidx = padded index (kx n1 + ky) nzp + z)
//      void  apply(coef, q[NF])             edits the NF spectral values in place
//
//  Per tile every field has one shared-memory buffer B[f] of L x 8 complex.  It is, in turn, the landing
//  zone of the asynchronous loads (thread-owned rows t + TPL j), the exchange scratch of the forward
//  transform, the parking place of the spectrum until its partners are ready (same thread-owned rows),
//  and the exchange scratch of the inverse transform.  As soon as the inverse transform of field f has
//  left the buffer, the loads of field f of the CTA's NEXT tile are issued into it, so they fly during
//  the remaining inverse transforms, the stores and the next tile's first forward transforms.
// ------------------------------------------------------------------------------------------------
// CTAs per SM the fused x pass is compiled for: a single field leaves room for more resident tiles (latency hiding)
template <int L, int NF>
constexpr int xmix_ctas_per_sm() { return L >= 512 ? (NF == 1 ? 2 : 1) : (NF == 1 ? 3 : (L >= 128 ? 2 : 3)); }
template <int L>
inline constexpr bool kXmixWide = (L == 512);

// PUSH (slab plans with peer pointers): the result of the inverse x transform is not stored in place but straight into the
// LOCAL-layout buffers (n0_loc, n1, nzp) of the ranks that own the x-planes -- the transposition back rides on the stores of
// this kernel over NVLink, and the inverse y pass that follows is the ordinary local in-place pass.
struct XmixPush {
    cd* peer[4][8];            // [field][rank]: that rank's local-layout buffer of the field
    int n0_loc_log2;           // x-planes per rank = 1 << n0_loc_log2
    int n1;                    // rows of a local-layout plane
    int y0;                    // first global row of this rank: rank * n1_loc
};

// NIN / NOUT (gradient: one spectrum in, three out; divergence: three in, one out): fields f >= NIN are neither loaded nor
// transformed forward, fields f >= NOUT are neither transformed back nor stored; all NF take part in the multiply.
template <int L, int NF, class Mix, bool PUSH = false, int NIN = NF, int NOUT = NF>
__global__ void __launch_bounds__((SPass<L, kXmixWide<L>>::THREADS), (xmix_ctas_per_sm<L, NF>()))
    xmix_kernel(SPassFields fields, SPassGeom geo, KGeom kg, Mix mix, const __grid_constant__ XmixPush push) {
    constexpr bool W = kXmixWide<L>;
    using P = SPass<L, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* tw = reinterpret_cast<cd*>(smem_raw);
    spass_load_twiddles<L>(tw);
    const int tile_in_cta = threadIdx.x / P::TILE_THREADS;
    const int tid = threadIdx.x % P::TILE_THREADS;
    const int t = tid / P::ZC, c = tid % P::ZC;
    cd* B0 = tw + L + (size_t)tile_in_cta * (NF * P::TILE_CD);          // B[f] = B0 + f * TILE_CD
    cd* own = B0 + t * P::ZC + c;                                       // own[f * TILE_CD + (TPL j) * ZC]: row t + TPL j
    constexpr int ROWSTEP = P::TPL * P::ZC;
    const long long total = spass_tiles(geo);
    const uint32_t xs = (uint32_t)geo.axis_stride;                      // < 2^32 elements for every supported shape
    auto locate = [&](long long w0) { return spass_locate(geo, w0 + tile_in_cta, total, c); };
    auto issue = [&](int f, const TileAt& a) {
        if (a.live && f < NIN) {
            const cd* base = fields.f[f] + a.off + (size_t)t * xs;
#pragma unroll
            for (int j = 0; j < P::EPT; ++j) cp_async16(own + f * P::TILE_CD + j * ROWSTEP, base + (size_t)(P::TPL * j) * xs);
        }
        cp_async_commit();
    };

    long long w0 = (long long)blockIdx.x * P::TPC;
    TileAt cur = locate(w0);
#pragma unroll
    for (int f = 0; f < NF; ++f) issue(f, cur);

    for (; w0 < total; w0 += (long long)gridDim.x * P::TPC) {
        const TileAt nxt = locate(w0 + (long long)gridDim.x * P::TPC);
        const size_t prow = ((size_t)cur.o) * kg.nzp_pad + cur.z;         // + kx n1_loc nzp  (tables cover the plan's own rows)
        const size_t kxs = (size_t)kg.n1_loc * kg.nzp_pad;
        cd v[P::EPT];
        if constexpr (NIN < NF) {
#pragma unroll
            for (int j = 0; j < P::EPT; ++j) v[j] = cd{0.0, 0.0};
        }
        // forward transforms; every spectrum but the last is parked in its thread-owned rows
#pragma unroll
        for (int f = 0; f < NIN; ++f) {
            if (f == 0) cp_async_wait<NF - 1>();
            else if (f == 1) cp_async_wait<(NF > 2 ? NF - 2 : 0)>();
            else if (f == 2) cp_async_wait<(NF > 3 ? NF - 3 : 0)>();
            else cp_async_wait<0>();
            cd* Bf = own + f * P::TILE_CD;
#pragma unroll
            for (int j = 0; j < P::EPT; ++j) v[j] = cur.live ? Bf[j * ROWSTEP] : cd{0.0, 0.0};
            tile_fft<L, -1, W>(v, B0 + f * P::TILE_CD, t, c, tw);
            if (f < NF - 1) {
#pragma unroll
                for (int s = 0; s < P::EPT; ++s) Bf[s * ROWSTEP] = v[s];
            }
        }
        // multiply; all mixed spectra go back to the parking rows.  The multiplier data comes from global
        // memory: keep PF slots of it in flight.
        {
            constexpr int PF = Mix::kRing;      // table-driven mixes: 8 slots in flight (measured at 256^3: 269 us; 4 + prefetch.L2 hints: 283 us); computed multipliers: 1
            const typename Mix::Line kl = mix.line(kg, cur.o + kg.j1_off, cur.z);      // global row index (slab plans: the rank's y range)
            typename Mix::Coef ring[PF];
#pragma unroll
            for (int s = 0; s < PF; ++s) {
                const int kx = spass_out_index<L, W>(t, s);
                ring[s] = mix.fetch(kl, kx, prow + (size_t)kx * kxs, cur.live);
            }
#pragma unroll
            for (int s = 0; s < P::EPT; ++s) {
                const typename Mix::Coef coef = ring[s % PF];
                if (s + PF < P::EPT) {
                    const int kxn = spass_out_index<L, W>(t, s + PF < P::EPT ? s + PF : s);
                    ring[s % PF] = mix.fetch(kl, kxn, prow + (size_t)kxn * kxs, cur.live);
                }
                cd q[NF];
#pragma unroll
                for (int f = 0; f < NF - 1; ++f) q[f] = f < NIN ? own[f * P::TILE_CD + s * ROWSTEP] : cd{0.0, 0.0};
                q[NF - 1] = v[s];
                if (cur.live) mix.apply(coef, q);
#pragma unroll
                for (int f = 0; f < NOUT; ++f) own[f * P::TILE_CD + s * ROWSTEP] = q[f];
            }
        }
        // inverse transforms; the buffer of field f is refilled with the next tile as soon as it is free
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            if (f >= NOUT) {          // not an output: its buffer only has to take the next tile
                issue(f, nxt);
                continue;
            }
            cd* Bf = own + f * P::TILE_CD;
            cd u[P::EPT];
#pragma unroll
            for (int j = 0; j < P::EPT; ++j) u[j] = Bf[spass_slot_of_input<L, W>(j) * ROWSTEP];
            tile_fft<L, +1, W>(u, B0 + f * P::TILE_CD, t, c, tw);
            issue(f, nxt);
            if (cur.live) {
                if constexpr (PUSH) {
                    const size_t rowz = (size_t)(push.y0 + cur.o) * kg.nzp_pad + cur.z;
#pragma unroll
                    for (int s = 0; s < P::EPT; ++s) {
                        const int x = spass_out_index<L, W>(t, s);
                        const int r = x >> push.n0_loc_log2, xl = x - (r << push.n0_loc_log2);
                        push.peer[f][r][(size_t)xl * push.n1 * kg.nzp_pad + rowz] = u[s];
                    }
                } else {
                    cd* base = fields.f[f] + cur.off;
#pragma unroll
                    for (int s = 0; s < P::EPT; ++s) base[(size_t)spass_out_index<L, W>(t, s) * xs] = u[s];
                }
            }
        }
        cur = nxt;
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
//  x pass of ONE field with a computed multiplier (-k^2, 4 pi / k^2): nothing has to be parked, so it is the plain pass with
//  two transforms per tile -- load, FFT, multiply in registers, inverse FFT straight from the forward's slots, store -- at the
//  plain pass's occupancy (4 CTAs / SM) instead of the persistent three-buffer machinery of xmix_kernel.
// ------------------------------------------------------------------------------------------------
template <int L, class Mix>
__global__ void __launch_bounds__(128, SPass<L>::CTAS_PER_SM) xone_kernel(cd* __restrict__ field, SPassGeom geo, KGeom kg, Mix mix) {
    using P = SPass<L>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* tw = reinterpret_cast<cd*>(smem_raw);
    spass_load_twiddles<L>(tw);
    const int tile_in_cta = threadIdx.x / P::TILE_THREADS;
    const int tid = threadIdx.x % P::TILE_THREADS;
    const int t = tid / P::ZC, c = tid % P::ZC;
    cd* S = tw + L + (size_t)tile_in_cta * P::TILE_CD;
    const long long total = spass_tiles(geo);
    for (long long w0 = (long long)blockIdx.x * P::TPC; w0 < total; w0 += (long long)gridDim.x * P::TPC) {
        const TileAt a = spass_locate(geo, w0 + tile_in_cta, total, c);
        cd* base = field + a.off;
        cd v[P::EPT];
#pragma unroll
        for (int j = 0; j < P::EPT; ++j) v[j] = a.live ? base[(long long)(t + P::TPL * j) * geo.axis_stride] : cd{0.0, 0.0};
        tile_fft<L, -1>(v, S, t, c, tw);
        const typename Mix::Line kl = mix.line(kg, a.o + kg.j1_off, a.z);
#pragma unroll
        for (int s = 0; s < P::EPT; ++s) {
            cd q[1] = {v[s]};
            mix.apply(mix.fetch(kl, spass_out_index<L>(t, s), 0, a.live), q);
            v[s] = q[0];
        }
        cd u[P::EPT];
#pragma unroll
        for (int j = 0; j < P::EPT; ++j) u[j] = v[spass_slot_of_input<L>(j)];
        tile_fft<L, +1>(u, S, t, c, tw);
        if (a.live) {
#pragma unroll
            for (int s = 0; s < P::EPT; ++s) base[(long long)spass_out_index<L>(t, s) * geo.axis_stride] = u[s];
        }
    }
}

template <int L, bool WIDE = false>
constexpr int spass_smem_bytes(int tile_buffers) {
    return (L + SPass<L, WIDE>::TPC * SPass<L, WIDE>::TILE_CD * tile_buffers) * 16;
}
