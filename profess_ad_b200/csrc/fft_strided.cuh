// Strided-axis (x or y) complex FFT passes over the padded half-spectrum layout spec[x][y][nzp],
// with the reciprocal-space multiply fused between the forward and the inverse x transform.
//
// Tile = L points along the transformed axis x 8 consecutive z columns (128-byte row segments).
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

template <int L>
struct SPass {
    static constexpr int TPL = (L == 256) ? 16 : 8;      // threads per line
    static constexpr int EPT = L / TPL;                  // points per thread: 16 (L = 256, 128), 8 (L = 64)
    static constexpr int G = EPT / TPL;                  // second-stage FFTs per thread: 1, 2, 1
    static constexpr int ZC = 8;                         // z columns per tile
    static constexpr int TILE_THREADS = TPL * ZC;        // 128, 64, 64
    static constexpr int THREADS = 128;
    static constexpr int TPC = THREADS / TILE_THREADS;   // tiles per CTA
    static constexpr int TILE_CD = L * ZC;               // complex numbers per tile
    static_assert(L == 256 || L == 128 || L == 64, "strided pass: L must be 64, 128 or 256");
};

// natural index along the line held by register slot s after tile_fft
template <int L>
__device__ __forceinline__ int spass_out_index(int t, int s) {
    using P = SPass<L>;
    return (t + P::TPL * (s / P::TPL)) + P::EPT * fft_nat<P::TPL>(s % P::TPL);
}
// register slot (after tile_fft) that holds the input j of a following tile_fft: k = t + TPL j
template <int L>
__device__ __forceinline__ constexpr int spass_slot_of_input(int j) {
    using P = SPass<L>;
    // k = t + TPL g + EPT k2 = t + TPL (g + G k2)  ->  j = g + G k2
    return (j % P::G) * P::TPL + fft_slot<P::TPL>(j / P::G);
}

// In: v[j] = z[t + TPL j].  Out: v[s] = Z[spass_out_index(t, s)].  S: tile scratch of L * 8 cd.  tw[e] = exp(-2 pi i e / L).
template <int L, int DIR>
__device__ __forceinline__ void tile_fft(cd* v, cd* S, int t, int c, const cd* __restrict__ tw) {
    using P = SPass<L>;
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
    long long outer_stride;    // complex elements between consecutive lines groups (the non-transformed, non-z axis)
    int n_outer;               // extent of that axis
    int zc0, nzc;              // first z chunk and number of z chunks handled by this launch
    int nzh;                   // live z columns (n2/2 + 1); columns >= nzh are padding and are not touched
};

struct SPassFields {
    cd* f[4];
};

// ------------------------------------------------------------------------------------------------
//  plain pass: in-place FFT along the strided axis for nf fields
// ------------------------------------------------------------------------------------------------
template <int L, int DIR>
__global__ void __launch_bounds__(128) spass_kernel(SPassFields fields, int nf, SPassGeom geo) {
    using P = SPass<L>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* tw = reinterpret_cast<cd*>(smem_raw);
    spass_load_twiddles<L>(tw);
    const int tile_in_cta = threadIdx.x / P::TILE_THREADS;
    const int tid = threadIdx.x % P::TILE_THREADS;
    const int t = tid / P::ZC, c = tid % P::ZC;
    cd* S = tw + L + (size_t)tile_in_cta * P::TILE_CD;
    const long long per_field = (long long)geo.n_outer * geo.nzc;
    const long long total = per_field * nf;
    for (long long w0 = (long long)blockIdx.x * P::TPC; w0 < total; w0 += (long long)gridDim.x * P::TPC) {
        const long long w = w0 + tile_in_cta;
        const bool live_tile = w < total;
        const long long wf = live_tile ? w : 0;
        const int fi = (int)(wf / per_field);
        const int rem = (int)(wf - (long long)fi * per_field);
        const int o = rem / geo.nzc;
        const int z = (geo.zc0 + (rem - o * geo.nzc)) * P::ZC + c;
        const bool live = live_tile && z < geo.nzh;
        cd* base = fields.f[fi] + (long long)o * geo.outer_stride + z;
        cd v[P::EPT];
#pragma unroll
        for (int j = 0; j < P::EPT; ++j) v[j] = live ? base[(long long)(t + P::TPL * j) * geo.axis_stride] : cd{0.0, 0.0};
        tile_fft<L, DIR>(v, S, t, c, tw);
        if (live) {
#pragma unroll
            for (int s = 0; s < P::EPT; ++s) base[(long long)spass_out_index<L>(t, s) * geo.axis_stride] = v[s];
        }
    }
}

// ------------------------------------------------------------------------------------------------
//  x pass with fused multiply:  NF coupled fields.  mix(kpoint, idx, f[NF]) edits the NF spectral
//  values of one k-point in place; idx = (kx n1 + ky) nzh + kz is the unpadded half-spectrum index.
//  The transformed axis is axis 0, the outer axis is axis 1.
// ------------------------------------------------------------------------------------------------
template <int L, int NF, class Mix>
__global__ void __launch_bounds__(128) xmix_kernel(SPassFields fields, SPassGeom geo, KGeom kg, Mix mix) {
    using P = SPass<L>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* tw = reinterpret_cast<cd*>(smem_raw);
    spass_load_twiddles<L>(tw);
    const int tile_in_cta = threadIdx.x / P::TILE_THREADS;
    const int tid = threadIdx.x % P::TILE_THREADS;
    const int t = tid / P::ZC, c = tid % P::ZC;
    cd* S = tw + L + (size_t)tile_in_cta * P::TILE_CD;
    // park[f][slot][thread]: thread-private, conflict-free (consecutive threads -> consecutive 16-byte words)
    cd* park = tw + L + (size_t)P::TPC * P::TILE_CD + threadIdx.x;
    constexpr int PARK_FIELD = P::EPT * P::THREADS;
    const long long total = (long long)geo.n_outer * geo.nzc;
    for (long long w0 = (long long)blockIdx.x * P::TPC; w0 < total; w0 += (long long)gridDim.x * P::TPC) {
        const long long w = w0 + tile_in_cta;
        const bool live_tile = w < total;
        const int rem = (int)(live_tile ? w : 0);
        const int o = rem / geo.nzc;
        const int z = (geo.zc0 + (rem - o * geo.nzc)) * P::ZC + c;
        const bool live = live_tile && z < geo.nzh;
        const long long off = (long long)o * geo.outer_stride + z;
        cd v[P::EPT];
        // forward transforms; fields NF-1 .. 1 are parked, field 0 stays in registers
#pragma unroll
        for (int f = NF - 1; f >= 0; --f) {
            const cd* base = fields.f[f] + off;
#pragma unroll
            for (int j = 0; j < P::EPT; ++j) v[j] = live ? base[(long long)(t + P::TPL * j) * geo.axis_stride] : cd{0.0, 0.0};
            tile_fft<L, -1>(v, S, t, c, tw);
            if (f > 0) {
#pragma unroll
                for (int s = 0; s < P::EPT; ++s) park[(f - 1) * PARK_FIELD + s * P::THREADS] = v[s];
            }
        }
        // multiply
#pragma unroll
        for (int s = 0; s < P::EPT; ++s) {
            cd q[NF];
            q[0] = v[s];
#pragma unroll
            for (int f = 1; f < NF; ++f) q[f] = park[(f - 1) * PARK_FIELD + s * P::THREADS];
            if (live) {
                const int kx = spass_out_index<L>(t, s);
                const KPoint kp = make_kpoint_at(kg, kx, o, z);
                mix(kp, ((uint32_t)kx * (uint32_t)kg.n1 + (uint32_t)o) * (uint32_t)kg.nzh + (uint32_t)z, q);
            }
            v[s] = q[0];
#pragma unroll
            for (int f = 1; f < NF; ++f) park[(f - 1) * PARK_FIELD + s * P::THREADS] = q[f];
        }
        // inverse transforms
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            cd u[P::EPT];
#pragma unroll
            for (int j = 0; j < P::EPT; ++j) {
                const int s = spass_slot_of_input<L>(j);
                u[j] = (f == 0) ? v[s] : park[(f - 1) * PARK_FIELD + s * P::THREADS];
            }
            tile_fft<L, +1>(u, S, t, c, tw);
            if (live) {
                cd* base = fields.f[f] + off;
#pragma unroll
                for (int s = 0; s < P::EPT; ++s) base[(long long)spass_out_index<L>(t, s) * geo.axis_stride] = u[s];
            }
        }
    }
}

template <int L>
constexpr int spass_smem_bytes(int parked_fields) {
    return (L + SPass<L>::TPC * SPass<L>::TILE_CD + parked_fields * SPass<L>::EPT * SPass<L>::THREADS) * 16;
}
