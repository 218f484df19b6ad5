// Register-level fp64 FFT building blocks for the hand-written transform passes (sm_100a).
//
//  * fft_reg<R, DIR>: R-point complex FFT (R = 2, 4, 8, 16) entirely in registers, radix-4/2
//    decimation; the output with natural index k ends up in register slot fft_slot<R>(k).
//  * line_fft<M, TPL, DIR>: M-point complex FFT of one line shared by TPL consecutive lanes of a
//    warp (M / TPL elements per lane), four-step: radix-(M/TPL) in registers, twiddle, one
//    shared-memory exchange (warp-synchronous, no block barrier), radix-TPL in registers.
//  * real <-> half-complex post/pre-processing for lines of 2M reals packed as M complex numbers.
//
// DIR = -1: forward (exp(-i...)), DIR = +1: inverse, both unnormalised.
#pragma once
#include <cuda_runtime.h>

struct __align__(16) cd {
    double x, y;
};
__device__ __forceinline__ cd operator+(cd a, cd b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cd operator-(cd a, cd b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cd cmul(cd a, cd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd cconj(cd a) { return {a.x, -a.y}; }
__device__ __forceinline__ cd cscale(cd a, double s) { return {a.x * s, a.y * s}; }

// multiply by DIR * i  (forward butterflies need -i)
template <int DIR>
__device__ __forceinline__ cd mul_i(cd a) {
    return DIR < 0 ? cd{a.y, -a.x} : cd{-a.y, a.x};
}

// twiddle table: c_tw[j] = exp(-2 pi i j / FFT_TW_N), j < FFT_TW_N; forward sign, conjugate for inverse
// (defined here: the header is included by exactly one translation unit, fftz.cu -- the library is built
//  without relocatable device code, so device symbols cannot be shared across .cu files)
#define FFT_TW_N 2048
__device__ double2 g_fft_tw[FFT_TW_N];

template <int DIR>
__device__ __forceinline__ cd twiddle(int num, int den_log2_shift /* FFT_TW_N / den */) {
    const double2 w = g_fft_tw[(num * den_log2_shift) & (FFT_TW_N - 1)];
    return DIR < 0 ? cd{w.x, w.y} : cd{w.x, -w.y};
}

template <int DIR>
__device__ __forceinline__ void fft2(cd& a, cd& b) {
    const cd t = a - b;
    a = a + b;
    b = t;
}

template <int DIR>
__device__ __forceinline__ void fft4(cd& a0, cd& a1, cd& a2, cd& a3) {
    const cd t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = mul_i<DIR>(a1 - a3);
    a0 = t0 + t2;
    a1 = t1 + t3;
    a2 = t0 - t2;
    a3 = t1 - t3;
}

// constant twiddles exp(DIR * 2 pi i m / 16)
template <int DIR>
__device__ __forceinline__ cd w16(int m) {
    constexpr double c1 = 0.92387953251128674, s1 = 0.38268343236508977, h = 0.70710678118654752;
    double c, s;
    switch (m & 15) {
        case 0: c = 1; s = 0; break;
        case 1: c = c1; s = s1; break;
        case 2: c = h; s = h; break;
        case 3: c = s1; s = c1; break;
        case 4: c = 0; s = 1; break;
        case 5: c = -s1; s = c1; break;
        case 6: c = -h; s = h; break;
        case 7: c = -c1; s = s1; break;
        case 8: c = -1; s = 0; break;
        case 9: c = -c1; s = -s1; break;
        case 10: c = -h; s = -h; break;
        case 11: c = -s1; s = -c1; break;
        case 12: c = 0; s = -1; break;
        case 13: c = s1; s = -c1; break;
        case 14: c = h; s = -h; break;
        default: c = c1; s = -s1; break;
    }
    return cd{c, DIR < 0 ? -s : s};
}

// constant twiddles exp(DIR * 2 pi i m / 32), m < 16
template <int DIR>
__device__ __forceinline__ cd w32(int m) {
    constexpr double c[9] = {1.0, 0.98078528040323043, 0.92387953251128674, 0.83146961230254524, 0.70710678118654752,
                             0.55557023301960218, 0.38268343236508977, 0.19509032201612825, 0.0};
    const int q = m & 15;
    const double co = q <= 8 ? c[q] : -c[16 - q];
    const double si = q <= 8 ? c[8 - q] : c[q - 8];
    return cd{co, DIR < 0 ? -si : si};
}

// register slot that holds natural output index k after fft_reg<R>
template <int R>
__device__ __forceinline__ constexpr int fft_slot(int k) {
    return R == 32 ? (((k & 15) >> 2) + 4 * (k & 3)) + 16 * (k >> 4)
         : R == 16 ? (k >> 2) + 4 * (k & 3) : R == 8 ? (k >> 2) + 2 * (k & 3) : k;
}
// natural output index held by register slot r (inverse of fft_slot)
template <int R>
__device__ __forceinline__ constexpr int fft_nat(int r) {
    return R == 32 ? (((r & 15) >> 2) + 4 * (r & 3)) + 16 * (r >> 4)
         : R == 16 ? (r >> 2) + 4 * (r & 3) : R == 8 ? (r >> 1) + 4 * (r & 1) : r;
}

template <int R, int DIR>
__device__ __forceinline__ void fft_reg(cd* v) {
    if constexpr (R == 2) {
        fft2<DIR>(v[0], v[1]);
    } else if constexpr (R == 4) {
        fft4<DIR>(v[0], v[1], v[2], v[3]);
    } else if constexpr (R == 8) {
        // n = n1 + 2 n2 ; k = k2 + 4 k1
#pragma unroll
        for (int n1 = 0; n1 < 2; ++n1) fft4<DIR>(v[n1], v[n1 + 2], v[n1 + 4], v[n1 + 6]);
        v[3] = cmul(v[3], w16<DIR>(2));      // W8^1
        v[5] = mul_i<DIR>(v[5]);             // W8^2
        v[7] = cmul(v[7], w16<DIR>(6));      // W8^3
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) fft2<DIR>(v[2 * k2], v[2 * k2 + 1]);
    } else if constexpr (R == 32) {
        // n = n1 + 2 n2: two 16-point transforms (even, odd inputs), twiddle, radix-2; slot s holds A[nat16(s)],
        // so X[nat16(s)] goes to slot s and X[nat16(s) + 16] to slot 16 + s
        cd a[16], b[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { a[j] = v[2 * j]; b[j] = v[2 * j + 1]; }
        fft_reg<16, DIR>(a);
        fft_reg<16, DIR>(b);
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            const int k = fft_nat<16>(s);
            const cd t = k == 0 ? b[s] : (k == 8 ? mul_i<DIR>(b[s]) : cmul(b[s], w32<DIR>(k)));
            v[s] = a[s] + t;
            v[16 + s] = a[s] - t;
        }
    } else {
        static_assert(R == 16, "fft_reg: R must be 2, 4, 8, 16 or 32");
        // n = n1 + 4 n2 ; k = k2 + 4 k1
#pragma unroll
        for (int n1 = 0; n1 < 4; ++n1) fft4<DIR>(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
#pragma unroll
        for (int n1 = 1; n1 < 4; ++n1)
#pragma unroll
            for (int k2 = 1; k2 < 4; ++k2) v[n1 + 4 * k2] = cmul(v[n1 + 4 * k2], w16<DIR>(n1 * k2));
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) fft4<DIR>(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
    }
}

// ---------------------------------------------------------------------------------------------
//  M-point complex FFT of one line spread over TPL lanes.  On entry lane t holds
//  v[j] = z[t + TPL j], j < EPT = M / TPL.  On exit v[g * TPL + r] = Z[k] with
//  k = (t + TPL g) + EPT * fft_nat<TPL>(r),  g < EPT / TPL.
//  S: shared scratch of the line, M * (1 + 1/TPL) complex (padded rows of TPL + 1).
// ---------------------------------------------------------------------------------------------
template <int DIR>
__device__ __forceinline__ cd tw_dir(cd w) {
    return DIR < 0 ? w : cd{w.x, -w.y};
}

// tw1: table in SHARED memory, tw1[e] = exp(-2 pi i e / M), e < M (global-memory twiddles stall the
// FFT on L2 latency: with most of the L1 carved out as shared memory the table does not stay resident)
template <int M, int TPL, int DIR>
__device__ __forceinline__ void line_fft(cd* v, cd* S, int t, const cd* __restrict__ tw1) {
    constexpr int EPT = M / TPL;
    constexpr int G = EPT / TPL;
    static_assert(EPT % TPL == 0, "line_fft: M / TPL must be a multiple of TPL");
    fft_reg<EPT, DIR>(v);
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        const int k1 = fft_nat<EPT>(r);
        cd a = v[r];
        if (k1 != 0) a = cmul(a, tw_dir<DIR>(tw1[t * k1]));
        S[k1 * (TPL + 1) + t] = a;
    }
    __syncwarp();
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int k1 = t + TPL * g;
        cd u[TPL];
#pragma unroll
        for (int t2 = 0; t2 < TPL; ++t2) u[t2] = S[k1 * (TPL + 1) + t2];
        fft_reg<TPL, DIR>(u);
#pragma unroll
        for (int r = 0; r < TPL; ++r) v[g * TPL + r] = u[r];
    }
    __syncwarp();
}

template <int M, int TPL>
__device__ __forceinline__ int line_fft_out_index(int t, int slot) {
    constexpr int EPT = M / TPL;
    const int g = slot / TPL, r = slot % TPL;
    return (t + TPL * g) + EPT * fft_nat<TPL>(r);
}

// ---------------------------------------------------------------------------------------------
//  8-elements-per-lane variant: M = 64 (TPL = 8) or M = 128 (TPL = 16).  On entry lane t holds
//  v[j] = z[t + TPL j], j < 8; on exit slot r holds Z[t + TPL * fft_nat<8>(r)]: the output is
//  distributed over the lanes exactly like the input, so real-space data never needs re-staging.
//  For M = 128 the second stage is a 16-point FFT per k1 shared by the two lanes (k1, h):
//  lane h computes the outputs k2 = 2 m + h as an 8-point FFT of (y[t'] +- y[t' + 8]) W16^{t' h}.
//  S: shared scratch of the line, 8 rows of TPL + 1 complex.
// ---------------------------------------------------------------------------------------------
template <int M, int TPL, int DIR>
__device__ __forceinline__ void line_fft8(cd* v, cd* S, int t, const cd* __restrict__ tw1) {
    static_assert((M == 64 && TPL == 8) || (M == 128 && TPL == 16) || (M == 256 && TPL == 32),
                  "line_fft8: (M, TPL) must be (64, 8), (128, 16) or (256, 32)");
    fft_reg<8, DIR>(v);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int k1 = fft_nat<8>(r);
        cd a = v[r];
        if (k1 != 0) a = cmul(a, tw_dir<DIR>(tw1[t * k1]));
        S[k1 * (TPL + 1) + t] = a;
    }
    __syncwarp();
    if constexpr (TPL == 8) {
#pragma unroll
        for (int t2 = 0; t2 < 8; ++t2) v[t2] = S[t * (TPL + 1) + t2];
    } else if constexpr (TPL == 32) {
        // 32-point FFT per k1 shared by the four lanes (k1, h): lane h computes the outputs k2 = 4 m + h as an 8-point
        // FFT of u[t'] = (sum_q y[t' + 8 q] W4^{q h}) W32^{t' h}
        const int k1 = t & 7, h = t >> 3;
        const bool odd = h & 1, neg = h & 2;
#pragma unroll
        for (int t2 = 0; t2 < 8; ++t2) {
            const cd y0 = S[k1 * (TPL + 1) + t2], y1 = S[k1 * (TPL + 1) + t2 + 8];
            const cd y2 = S[k1 * (TPL + 1) + t2 + 16], y3 = S[k1 * (TPL + 1) + t2 + 24];
            const cd A = odd ? y0 - y2 : y0 + y2;
            const cd Bq = y1 - y3, Bs = y1 + y3;
            const cd B = odd ? mul_i<DIR>(Bq) : Bs;
            cd u = neg ? A - B : A + B;
            if (t2 != 0) u = cmul(u, tw_dir<DIR>(tw1[(M / 32) * t2 * h]));      // W32^{t2 h}; h = 0: tw1[0] = 1
            v[t2] = u;
        }
    } else {
        const int k1 = t & 7, h = t >> 3;
#pragma unroll
        for (int t2 = 0; t2 < 8; ++t2) {
            const cd lo = S[k1 * (TPL + 1) + t2], hi = S[k1 * (TPL + 1) + t2 + 8];
            cd u = h ? lo - hi : lo + hi;
            if (h && t2 != 0) u = cmul(u, w16<DIR>(t2));
            v[t2] = u;
        }
    }
    fft_reg<8, DIR>(v);
    __syncwarp();
}
