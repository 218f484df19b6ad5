// Hand-written z-axis (contiguous axis) real<->half-complex FFT passes with the real-space
// elementwise work fused in, for power-of-two n2 in {128, 256, 512}.  The (x, y) axes are
// transformed in place by one batched 2-D cuFFT Z2Z plan over the padded half-spectrum layout
//      spec[x][y][nzp],  nzp = n2/2 + 8 (multiple of 8 complex = 128 B rows),  nzh = n2/2 + 1 used.
//
// Why: in the plain cuFFT pipeline every real-space pre/post kernel is a separate HBM round trip and
// the r2c/c2r pass runs at ~3.7 TB/s.  Here a warp owns 32/TPL lines; TPL lanes share one line:
// they load the line ONCE, generate every field that has to be transformed from it (e.g. n^beta,
// n^beta theta, n^beta theta^2/2, sqrt(n) for WGC99), run the M = n2/2 point packed complex FFT in
// registers (one warp-synchronous shared-memory exchange), do the real post-processing and write
// the half-spectrum lines.  The inverse pass mirrors it and ends in the real-space post-op (energy
// densities, potential assembly) without the fields ever touching HBM.
//
// Conventions match cuFFT: forward and inverse are unnormalised (the 1/N lives in the multipliers).
#include "common.cuh"
#include "fft_core.cuh"

namespace {

constexpr double k3Pi2 = 29.608813203268074;

// ------------------------------------------------------------------------------------------------
//  twiddle table
// ------------------------------------------------------------------------------------------------
bool g_tw_ready[64] = {false};

int ensure_twiddles(int device) {
    if (g_tw_ready[device & 63]) return PAD_OK;
    static double2 host[FFT_TW_N];
    for (int j = 0; j < FFT_TW_N; ++j) {
        // exact symmetries first, so that the table is bit-symmetric
        const double a = -2.0 * kPi * (double)j / (double)FFT_TW_N;
        host[j].x = cos(a);
        host[j].y = sin(a);
    }
    host[0] = {1.0, 0.0};
    host[FFT_TW_N / 4] = {0.0, -1.0};
    host[FFT_TW_N / 2] = {-1.0, 0.0};
    host[3 * FFT_TW_N / 4] = {0.0, 1.0};
    PAD_CUDA(cudaMemcpyToSymbol(g_fft_tw, host, sizeof(host)));
    g_tw_ready[device & 63] = true;
    return PAD_OK;
}

// ------------------------------------------------------------------------------------------------
//  shared-memory carving: per line  [ stage: NST * 2M doubles | S: M + M/TPL complex ]
// ------------------------------------------------------------------------------------------------
template <int M, int TPL, int NST>
struct ZLayout {
    static constexpr int kStageDoubles = NST * 2 * M;
    static constexpr int kScratchCplx = M + M / TPL + 1;           // also holds M + 1 natural-order values
    static constexpr int kLineBytes = kStageDoubles * 8 + kScratchCplx * 16;
    static constexpr int kLinesPerWarp = 32 / TPL;
};

// forward: NF fields generated from staged per-point values, r2c along z, written as padded half-spectra
//   Gen::NST                      staged doubles per point
//   gen.stage(gidx, out[NST])     values to keep for the point with global index gidx
//   gen.field(f, st[NST])         value of field f at a point
template <int M, int TPL, int NF, class Gen>
__global__ void __launch_bounds__(128) zfwd_kernel(Gen gen, cd* __restrict__ o0, cd* __restrict__ o1, cd* __restrict__ o2,
                                                  cd* __restrict__ o3, int nlines, int nzp) {
    using L = ZLayout<M, TPL, Gen::NST>;
    constexpr int EPT = M / TPL, LPW = L::kLinesPerWarp, NST = Gen::NST;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int sub = lane / TPL, t = lane % TPL;
    unsigned char* mine = smem_raw + (size_t)(warp * LPW + sub) * L::kLineBytes;
    double* st = reinterpret_cast<double*>(mine);
    cd* S = reinterpret_cast<cd*>(mine + L::kStageDoubles * 8);
    cd* outs[4] = {o0, o1, o2, o3};

    for (int line0 = (blockIdx.x * wpb + warp) * LPW; line0 < nlines; line0 += gridDim.x * wpb * LPW) {
        const int line = line0 + sub;
        const bool live = line < nlines;
        const size_t base = (size_t)line * (2 * M);
        // ---- load the line once, stage the per-point values
        if (live) {
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                const int z = 2 * (t + TPL * j);
                double a[NST], b[NST];
                gen.stage(base + z, base + z + 1, a, b);
#pragma unroll
                for (int s = 0; s < NST; ++s) {
                    st[s * 2 * M + z] = a[s];
                    st[s * 2 * M + z + 1] = b[s];
                }
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int f = 0; f < NF; ++f) {
            cd v[EPT];
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                const int z = 2 * (t + TPL * j);
                double a[NST], b[NST];
#pragma unroll
                for (int s = 0; s < NST; ++s) {
                    a[s] = st[s * 2 * M + z];
                    b[s] = st[s * 2 * M + z + 1];
                }
                v[j] = cd{gen.field(f, a), gen.field(f, b)};
            }
            line_fft<M, TPL, -1>(v, S, t);
            // natural order into S (all stage-2 reads are done after line_fft's trailing __syncwarp)
#pragma unroll
            for (int sl = 0; sl < EPT; ++sl) S[line_fft_out_index<M, TPL>(t, sl)] = v[sl];
            __syncwarp();
            // real post-processing: X[k] = Ev + w^k Od, X[M-k] = conj(Ev - w^k Od)
            cd* out = outs[f] + (size_t)line * nzp;
            for (int k = t; k <= M / 2; k += TPL) {
                const cd A = S[k], B = cconj(S[(M - k) & (M - 1)]);
                const cd Ev = cscale(A + B, 0.5);
                const cd D = A - B;                          // 2 i Od
                const cd Od = cd{0.5 * D.y, -0.5 * D.x};     // -i/2 * D
                const cd T = cmul(Od, twiddle<-1>(k, FFT_TW_N / (2 * M)));
                if (live) {
                    if (k == 0) {
                        out[0] = cd{A.x + A.y, 0.0};
                        out[M] = cd{A.x - A.y, 0.0};
                    } else {
                        out[k] = Ev + T;
                        out[M - k] = cconj(Ev - T);
                    }
                }
            }
            __syncwarp();
        }
    }
}

// inverse: NF padded half-spectra -> c2r along z -> post(gidx, values[NF], acc)
//   post.apply(gidx0, gidx1, u0[NF], u1[NF], acc)   for the two points of a packed pair
template <int M, int TPL, int NF, int NRED, class Post>
__global__ void __launch_bounds__(128) zinv_kernel(Post post, const cd* __restrict__ i0, const cd* __restrict__ i1,
                                                  const cd* __restrict__ i2, const cd* __restrict__ i3, int nlines, int nzp,
                                                  double* __restrict__ partials) {
    using L = ZLayout<M, TPL, NF>;
    constexpr int EPT = M / TPL, LPW = L::kLinesPerWarp;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int sub = lane / TPL, t = lane % TPL;
    unsigned char* mine = smem_raw + (size_t)(warp * LPW + sub) * L::kLineBytes;
    double* st = reinterpret_cast<double*>(mine);
    cd* S = reinterpret_cast<cd*>(mine + L::kStageDoubles * 8);
    const cd* ins[4] = {i0, i1, i2, i3};
    double acc[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int r = 0; r < (NRED > 0 ? NRED : 1); ++r) acc[r] = 0.0;

    for (int line0 = (blockIdx.x * wpb + warp) * LPW; line0 < nlines; line0 += gridDim.x * wpb * LPW) {
        const int line = line0 + sub;
        const bool live = line < nlines;
        const size_t base = (size_t)line * (2 * M);
#pragma unroll 1
        for (int f = 0; f < NF; ++f) {
            const cd* in = ins[f] + (size_t)(live ? line : 0) * nzp;
            // pre-processing: Z[k] = Ev + i Od, Z[M-k] = conj(Ev - i Od);  Ev = (X[k] + conj X[M-k])/2,
            // w^k Od = (X[k] - conj X[M-k])/2.  Imaginary parts of X[0], X[M] are ignored (c2r semantics).
            for (int k = t; k <= M / 2; k += TPL) {
                cd Xa = in[k], Xb = in[M - k];
                if (k == 0) { Xa.y = 0.0; Xb.y = 0.0; }
                const cd B = cconj(Xb);
                const cd Ev = cscale(Xa + B, 0.5);
                const cd T = cscale(Xa - B, 0.5);
                const cd Od = cmul(T, twiddle<+1>(k, FFT_TW_N / (2 * M)));      // conj(w^k) T
                const cd iOd = cd{-Od.y, Od.x};
                S[k] = Ev + iOd;
                if (k != 0 && k != M - k) S[M - k] = cconj(Ev - iOd);
            }
            __syncwarp();
            cd v[EPT];
#pragma unroll
            for (int j = 0; j < EPT; ++j) v[j] = S[t + TPL * j];
            __syncwarp();
            line_fft<M, TPL, +1>(v, S, t);
            // packed complex n -> reals 2n, 2n+1; factor 2 makes it the unnormalised c2r of length 2M
#pragma unroll
            for (int sl = 0; sl < EPT; ++sl) {
                const int n = line_fft_out_index<M, TPL>(t, sl);
                st[f * 2 * M + 2 * n] = 2.0 * v[sl].x;
                st[f * 2 * M + 2 * n + 1] = 2.0 * v[sl].y;
            }
            __syncwarp();
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                const int z = 2 * (t + TPL * j);
                double u0[NF], u1[NF];
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    u0[f] = st[f * 2 * M + z];
                    u1[f] = st[f * 2 * M + z + 1];
                }
                post.apply(base + z, base + z + 1, u0, u1, acc);
            }
        }
        __syncwarp();
    }
    if constexpr (NRED > 0) {
        // 128-thread block reduction (4 warps)
        __shared__ double red[NRED][4];
#pragma unroll
        for (int r = 0; r < NRED; ++r) {
            const double w = warp_sum(acc[r]);
            if (lane == 0) red[r][warp] = w;
        }
        __syncthreads();
        if (threadIdx.x < NRED) {
            double s = 0.0;
            for (int w = 0; w < wpb; ++w) s += red[threadIdx.x][w];
            partials[(size_t)threadIdx.x * PAD_MAX_BLOCKS + blockIdx.x] = s;
        }
    }
}

template <int M, int TPL, int NST>
constexpr int zsmem_bytes(int warps) {
    return warps * ZLayout<M, TPL, NST>::kLinesPerWarp * ZLayout<M, TPL, NST>::kLineBytes;
}

inline int zgrid(int nlines, int lines_per_block, int blocks_per_sm) {
    int b = (nlines + lines_per_block - 1) / lines_per_block;
    const int cap = 148 * blocks_per_sm;
    if (b > cap) b = cap;
    if (b > PAD_MAX_BLOCKS) b = PAD_MAX_BLOCKS;
    return b < 1 ? 1 : b;
}

template <int M, int TPL, int NF, class Gen>
int launch_zfwd(pad_plan* p, cudaStream_t s, Gen gen, cd* o0, cd* o1, cd* o2, cd* o3) {
    constexpr int warps = 4;
    constexpr int smem = zsmem_bytes<M, TPL, Gen::NST>(warps);
    static bool attr = false;
    auto kern = zfwd_kernel<M, TPL, NF, Gen>;
    if (!attr) {
        PAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    const int nlines = p->n0 * p->n1;
    const int lpb = warps * (32 / TPL);
    const int bps = smem > 0 ? (227 * 1024) / smem : 4;
    kern<<<zgrid(nlines, lpb, bps < 1 ? 1 : bps), warps * 32, smem, s>>>(gen, o0, o1, o2, o3, nlines, p->nzp);
    ++g_pad_launches;
    PAD_CUDA(cudaGetLastError());
    return PAD_OK;
}

template <int M, int TPL, int NF, int NRED, class Post>
int launch_zinv(pad_plan* p, cudaStream_t s, Post post, const cd* i0, const cd* i1, const cd* i2, const cd* i3, int* grid_out) {
    constexpr int warps = 4;
    constexpr int smem = zsmem_bytes<M, TPL, NF>(warps);
    static bool attr = false;
    auto kern = zinv_kernel<M, TPL, NF, NRED, Post>;
    if (!attr) {
        PAD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    const int nlines = p->n0 * p->n1;
    const int lpb = warps * (32 / TPL);
    const int bps = (227 * 1024) / smem;
    const int grid = zgrid(nlines, lpb, bps < 1 ? 1 : bps);
    kern<<<grid, warps * 32, smem, s>>>(post, i0, i1, i2, i3, nlines, p->nzp, p->partials);
    ++g_pad_launches;
    if (grid_out) *grid_out = grid;
    PAD_CUDA(cudaGetLastError());
    return PAD_OK;
}

// ------------------------------------------------------------------------------------------------
//  batched 2-D (x, y) transform over the padded layout, in place
// ------------------------------------------------------------------------------------------------
int ensure_xy(pad_plan* p, cudaStream_t s) {
    if (!p->xy_ready) {
        p->nzp = p->n2 / 2 + 8;
        int n[2] = {p->n0, p->n1};
        size_t w = 0;
        PAD_CUFFT(cufftCreate(&p->xy));
        PAD_CUFFT(cufftSetAutoAllocation(p->xy, 0));
        PAD_CUFFT(cufftMakePlanMany(p->xy, 2, n, n, p->nzp, 1, n, p->nzp, 1, CUFFT_Z2Z, p->nzh, &w));
        if (w > 0) {
            PAD_CUDA(cudaMalloc(&p->xy_work, w));
            p->bytes_allocated += w;
        }
        PAD_CUFFT(cufftSetWorkArea(p->xy, p->xy_work));
        p->xy_stream = (cudaStream_t)(-1);
        p->xy_ready = true;
    }
    if (p->xy_stream != s) {
        PAD_CUFFT(cufftSetStream(p->xy, s));
        p->xy_stream = s;
    }
    return PAD_OK;
}

int xy_exec(pad_plan* p, cudaStream_t s, cd* spec, int dir) {
    PAD_TRY(ensure_xy(p, s));
    PAD_CUFFT(cufftExecZ2Z(p->xy, reinterpret_cast<cufftDoubleComplex*>(spec), reinterpret_cast<cufftDoubleComplex*>(spec),
                           dir < 0 ? CUFFT_FORWARD : CUFFT_INVERSE));
    ++g_pad_fft_execs;
    return PAD_OK;
}

int get_zbuf(pad_plan* p, int i, cd** out) {
    if (i < 0 || i >= 4) { pad_set_error("zbuf index %d", i); return PAD_ERR_ARG; }
    const size_t bytes = sizeof(cd) * (size_t)p->n0 * p->n1 * p->nzp;
    if (!p->zbuf[i]) {
        PAD_CUDA(cudaMalloc(&p->zbuf[i], bytes));
        PAD_CUDA(cudaMemset(p->zbuf[i], 0, bytes));       // padding columns stay zero forever
        p->bytes_allocated += bytes;
    }
    *out = reinterpret_cast<cd*>(p->zbuf[i]);
    return PAD_OK;
}

bool fast_shape(const pad_plan* p) { return p->n2 == 128 || p->n2 == 256 || p->n2 == 512; }

// dispatch on n2: M = n2/2; (M, TPL) in {(64, 8), (128, 8), (256, 16)}
#define ZDISPATCH(p, CALL)                                             \
    do {                                                               \
        if ((p)->n2 == 256) { constexpr int M = 128, TPL = 8; CALL; }  \
        else if ((p)->n2 == 128) { constexpr int M = 64, TPL = 8; CALL; } \
        else { constexpr int M = 256, TPL = 16; CALL; }                \
    } while (0)

// ------------------------------------------------------------------------------------------------
//  functors
// ------------------------------------------------------------------------------------------------
struct GenCopy {                       // plain r2c of one real field
    static constexpr int NST = 1;
    const double* f;
    __device__ void stage(size_t g0, size_t g1, double* a, double* b) const { a[0] = f[g0]; b[0] = f[g1]; }
    __device__ double field(int, const double* s) const { return s[0]; }
};

struct PostStore {                     // plain c2r of one real field
    double* out;
    __device__ void apply(size_t g0, size_t g1, const double* u0, const double* u1, double*) const {
        out[g0] = u0[0];
        out[g1] = u1[0];
    }
};

// WGC99, first forward pass: a = n^beta, a theta, a theta^2 / 2, chi = sqrt(n)   (functionals.py:974-981, :242-243)
struct GenWgcA {
    static constexpr int NST = 2;      // n, n^beta
    const double* den;
    const double* scal;
    double beta;
    __device__ void stage(size_t g0, size_t g1, double* a, double* b) const {
        const double n0 = den[g0], n1 = den[g1];
        a[0] = n0; a[1] = exp(beta * log(n0));
        b[0] = n1; b[1] = exp(beta * log(n1));
    }
    __device__ double field(int f, const double* s) const {
        const double th = s[0] - scal[S_NREF];
        switch (f) {
            case 0: return s[1];
            case 1: return s[1] * th;
            case 2: return 0.5 * s[1] * th * th;
            default: return s[0] != 0.0 ? sqrt(s[0]) : 0.0;
        }
    }
};

// WGC99, second forward pass: P = n^alpha (stored by the mid pass), P theta, P theta^2 / 2
struct GenWgcP {
    static constexpr int NST = 2;      // n, P
    const double* den;
    const double* P;
    const double* scal;
    __device__ void stage(size_t g0, size_t g1, double* a, double* b) const {
        a[0] = den[g0]; a[1] = P[g0];
        b[0] = den[g1]; b[1] = P[g1];
    }
    __device__ double field(int f, const double* s) const {
        const double th = s[0] - scal[S_NREF];
        return f == 0 ? s[1] : (f == 1 ? s[1] * th : 0.5 * s[1] * th * th);
    }
};

// WGC99 mid pass: u1, u2, u3, lap(chi) -> energy densities, first half of the potential, P
struct PostWgcMid {
    const double* den;
    const double* scal;
    double* v_out;
    double* P_out;
    double alpha;
    int accumulate, want_v;
    __device__ void one(size_t g, const double* u, double* acc) const {
        const double n = den[g];
        const double th = n - scal[S_NREF];
        const double P = exp(alpha * log(n));
        const double conv = u[0] + th * (u[1] + 0.5 * th * u[2]);
        const double c = cbrt(n);
        const double chi = n != 0.0 ? sqrt(n) : 0.0;
        acc[0] += kCTF * n * c * c;
        acc[1] += chi * u[3];
        acc[2] += P * conv;
        if (want_v) {
            double v = (5.0 / 3.0) * kCTF * c * c;
            if (n != 0.0) v += -0.5 * u[3] / chi;
            v += kCTF * (alpha * P / n * conv + P * (u[1] + th * u[2]));
            v_out[g] = accumulate ? v_out[g] + v : v;
            P_out[g] = P;
        }
    }
    __device__ void apply(size_t g0, size_t g1, const double* u0, const double* u1, double* acc) const {
        one(g0, u0, acc);
        one(g1, u1, acc);
    }
};

// WGC99 final pass: g1, g2, g3 -> second half of the potential
struct PostWgcFin {
    const double* den;
    const double* scal;
    double* v_out;
    double beta;
    __device__ void one(size_t g, const double* u) const {
        const double n = den[g];
        const double th = n - scal[S_NREF];
        const double a = exp(beta * log(n));
        const double da = beta * a / n;
        v_out[g] += kCTF * (da * u[0] + (da * th + a) * u[1] + (0.5 * da * th * th + a * th) * u[2]);
    }
    __device__ void apply(size_t g0, size_t g1, const double* u0, const double* u1, double*) const {
        one(g0, u0);
        one(g1, u1);
    }
};

// reciprocal-space kernel over the padded layout: f(unpadded index, padded index, kpoint)
template <class F>
__global__ void __launch_bounds__(PAD_THREADS) ksp_kernel(KGeom g, uint32_t nk, int nzp, F f) {
    const uint32_t stride = gridDim.x * PAD_THREADS;
    for (uint32_t idx = blockIdx.x * PAD_THREADS + threadIdx.x; idx < nk; idx += stride) {
        const KPoint k = make_kpoint(g, idx);
        const uint32_t row = idx / (uint32_t)g.nzh;
        f(idx, (size_t)row * nzp + (uint32_t)k.j2, k);
    }
}

template <class F>
void launch_ksp(pad_plan* p, cudaStream_t s, F f) {
    ksp_kernel<F><<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(p->geom, (uint32_t)p->Nk, p->nzp, f);
    ++g_pad_launches;
}

}  // namespace

// =================================================================================================
//  public: custom 3-D r2c / c2r (used by tests and by pad_gradient-like helpers)
// =================================================================================================
extern "C" int pad_fast_fft_supported(const pad_plan* p) { return p && fast_shape(p) ? 1 : 0; }

// out: padded half-spectrum (n0, n1, nzp) complex; returns nzp through *nzp_out
extern "C" int pad_rfft3_fast(pad_plan* p, const double* in, double* out_cplx_padded, int* nzp_out, void* stream) {
    if (!p || !in || !out_cplx_padded) { pad_set_error("pad_rfft3_fast: null argument"); return PAD_ERR_ARG; }
    if (!fast_shape(p)) { pad_set_error("pad_rfft3_fast: n2 = %d not supported (128, 256, 512)", p->n2); return PAD_ERR_ARG; }
    PAD_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    PAD_TRY(ensure_twiddles(p->device));
    PAD_TRY(ensure_xy(p, s));
    if (nzp_out) *nzp_out = p->nzp;
    cd* o = reinterpret_cast<cd*>(out_cplx_padded);
    GenCopy gen{in};
    ZDISPATCH(p, PAD_TRY((launch_zfwd<M, TPL, 1>(p, s, gen, o, nullptr, nullptr, nullptr))));
    PAD_TRY(xy_exec(p, s, o, -1));
    return PAD_OK;
}

// in: padded half-spectrum (destroyed); out: real field, unnormalised (N x the inverse)
extern "C" int pad_irfft3_fast(pad_plan* p, double* in_cplx_padded, double* out, void* stream) {
    if (!p || !in_cplx_padded || !out) { pad_set_error("pad_irfft3_fast: null argument"); return PAD_ERR_ARG; }
    if (!fast_shape(p)) { pad_set_error("pad_irfft3_fast: n2 = %d not supported (128, 256, 512)", p->n2); return PAD_ERR_ARG; }
    PAD_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    PAD_TRY(ensure_twiddles(p->device));
    cd* i = reinterpret_cast<cd*>(in_cplx_padded);
    PAD_TRY(xy_exec(p, s, i, +1));
    PostStore post{out};
    ZDISPATCH(p, PAD_TRY((launch_zinv<M, TPL, 1, 0>(p, s, post, i, nullptr, nullptr, nullptr, nullptr))));
    return PAD_OK;
}

// =================================================================================================
//  WGC99 on the fused pipeline.  Called by pad_eval_wgc99 after the kernel cache and S_NREF are ready.
// =================================================================================================
int pad_wgc99_fast_supported(const pad_plan* p) { return fast_shape(p) ? 1 : 0; }

int pad_wgc99_fast(pad_plan* p, const double* den, double alpha, double beta, const double* kern, double* E_out,
                   double* v_out, int accumulate, cudaStream_t s) {
    PAD_TRY(ensure_twiddles(p->device));
    PAD_TRY(ensure_xy(p, s));
    cd* B[4];
    for (int i = 0; i < 4; ++i) PAD_TRY(get_zbuf(p, i, &B[i]));
    double* Pbuf;
    PAD_TRY(pad_get_rbuf(p, 7, &Pbuf));
    const double* scal = p->scal;
    const size_t nk = p->Nk;
    const double inv_n = p->geom.inv_n;
    const double *W0 = kern, *K1 = kern + nk, *K2 = kern + 2 * nk, *K3 = kern + 3 * nk;
    const bool want_v = v_out != nullptr;

    GenWgcA genA{den, scal, beta};
    ZDISPATCH(p, PAD_TRY((launch_zfwd<M, TPL, 4>(p, s, genA, B[0], B[1], B[2], B[3]))));
    for (int i = 0; i < 4; ++i) PAD_TRY(xy_exec(p, s, B[i], -1));
    {
        cd *CA = B[0], *CB = B[1], *CC = B[2], *CX = B[3];
        launch_ksp(p, s, [=] __device__(uint32_t idx, size_t pidx, const KPoint& k) {
            const double w0 = W0[idx], k1 = K1[idx], k2 = K2[idx], k3 = K3[idx];
            const cd A = CA[pidx], Bb = CB[pidx], Cc = CC[pidx];
            CA[pidx] = cd{w0 * A.x + k1 * Bb.x + k2 * Cc.x, w0 * A.y + k1 * Bb.y + k2 * Cc.y};
            CB[pidx] = cd{k1 * A.x + k3 * Bb.x, k1 * A.y + k3 * Bb.y};
            CC[pidx] = cd{k2 * A.x, k2 * A.y};
            const double m = -inv_n * sym_even(k, [](double kx, double ky, double kz) { return kx * kx + ky * ky + kz * kz; });
            const cd X = CX[pidx];
            CX[pidx] = cd{X.x * m, X.y * m};
        });
        PAD_CUDA(cudaGetLastError());
    }
    for (int i = 0; i < 4; ++i) PAD_TRY(xy_exec(p, s, B[i], +1));
    int grid = 1;
    PostWgcMid mid{den, scal, v_out, Pbuf, alpha, accumulate, want_v ? 1 : 0};
    ZDISPATCH(p, PAD_TRY((launch_zinv<M, TPL, 4, 3>(p, s, mid, B[0], B[1], B[2], B[3], &grid))));
    if (E_out) {
        FinalizeArgs a;
        a.nblocks = grid; a.nterms = 3; a.accumulate = accumulate;
        for (int t = 0; t < PAD_MAX_RED; ++t) a.coef[t] = 0.0;
        a.coef[0] = p->dV; a.coef[1] = -0.5 * p->dV; a.coef[2] = kCTF * p->dV;
        a.sums_out = nullptr;
        a.E_out = E_out;
        pad_launch_finalize(p, a, s);
    }
    if (!want_v) return PAD_OK;

    GenWgcP genP{den, Pbuf, scal};
    ZDISPATCH(p, PAD_TRY((launch_zfwd<M, TPL, 3>(p, s, genP, B[0], B[1], B[2], nullptr))));
    for (int i = 0; i < 3; ++i) PAD_TRY(xy_exec(p, s, B[i], -1));
    {
        cd *CA = B[0], *CB = B[1], *CC = B[2];
        launch_ksp(p, s, [=] __device__(uint32_t idx, size_t pidx, const KPoint&) {
            const double w0 = W0[idx], k1 = K1[idx], k2 = K2[idx], k3 = K3[idx];
            const cd A = CA[pidx], Bb = CB[pidx], Cc = CC[pidx];
            CA[pidx] = cd{w0 * A.x + k1 * Bb.x + k2 * Cc.x, w0 * A.y + k1 * Bb.y + k2 * Cc.y};
            CB[pidx] = cd{k1 * A.x + k3 * Bb.x, k1 * A.y + k3 * Bb.y};
            CC[pidx] = cd{k2 * A.x, k2 * A.y};
        });
        PAD_CUDA(cudaGetLastError());
    }
    for (int i = 0; i < 3; ++i) PAD_TRY(xy_exec(p, s, B[i], +1));
    PostWgcFin fin{den, scal, v_out, beta};
    ZDISPATCH(p, PAD_TRY((launch_zinv<M, TPL, 3, 0>(p, s, fin, B[0], B[1], B[2], nullptr, nullptr))));
    return PAD_OK;
}
