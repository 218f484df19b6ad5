// Table-driven fp64 log / exp / rsqrt for the real-space WGC99 kernels (sm_100a).
//
// The B200 executes 64 fp64 operations per clock and SM; the library pow / cbrt / sqrt / division cost
// ~190 instructions per grid point in the mid pass, as much as all the FFT butterflies of that pass.  These
// versions are for POSITIVE, NORMAL, FINITE arguments only (a density and its powers); callers branch to the
// library functions otherwise.  Accuracy: |error| <= ~2 ulp (checked on the GPU against the library
// functions, tests/test_gpu_fastfft.py::test_fast_math_accuracy).
//
//   log x   : x = 2^e m, m in [1, 2); m = c_i (1 + r), c_i = 1 + (i + 1/2) / 128 from the top 7 mantissa bits,
//             |r| <= 2^-8;  log x = e ln2 + log c_i + log1p(r), degree-7 Taylor polynomial
//   exp y   : y = (64 k + j) ln2 / 64 + r, |r| <= ln2 / 128;  exp y = 2^k 2^(j/64) exp(r), degree-6 polynomial
//   rsqrt x : MUFU.RSQ64H seed (2^-20), two Newton steps
#pragma once
#include <cuda_runtime.h>

// g_fm_log[i] = { fl(1 / c_i), -log(fl(1 / c_i)) };  g_fm_exp[j] = 2^(j / 64)
// (defined here: this header is included by exactly one translation unit, like fft_core.cuh)
__device__ double2 g_fm_log[128];
__device__ double g_fm_exp[64];
// The lookups go to SHARED copies of the two tables (2.5 KB per CTA; every kernel that calls fm_log / fm_exp starts with
// fm_load_tables()).  Round 1 read them with __ldg: in the 8-warp inverse z kernel the first FMA after the log-table load
// alone held 10 % of all stall samples (long scoreboard, ncu source page) -- an L1/L2 round trip on the critical path of a
// dependent chain that only two warps per scheduler were there to hide.
__shared__ double2 s_fm_log[128];
__shared__ double s_fm_exp[64];

__device__ __forceinline__ void fm_load_tables() {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_fm_log[i] = g_fm_log[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_fm_exp[i] = g_fm_exp[i];
    __syncthreads();
}

// Polynomial coefficients and split constants live in the constant bank: as literals every one of them costs two UMOV
// (32-bit halves into a uniform register) right before its use -- 19-28 % of the instructions of the per-point functions
// in the round-1 SASS profile; from c[3][..] two of them arrive per LDCU.128.
__constant__ double c_fm[18] = {
    1.0 / 7.0, -1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, -0.5,           // 0..5   log1p(r) Taylor
    0.6931471803691238, 1.9082149292705877e-10,                  // 6, 7   ln2 hi (11 trailing zero bits), lo
    92.33248261689366, 6755399441055744.0,                       // 8, 9   64 / ln2, 1.5 * 2^52
    -0.01083042469326756, -2.9815858269852933e-12,               // 10, 11 -ln2/64 hi (21 trailing zero bits), lo
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0    // 12..17 exp(r) - 1 Taylor
};

__device__ __forceinline__ double fm_log(double x) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const int idx = (hi >> 13) & 127;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 t = s_fm_log[idx];
    const double r = fma(m, t.x, -1.0);
    double p = fma(r, c_fm[0], c_fm[1]);
    p = fma(r, p, c_fm[2]);
    p = fma(r, p, c_fm[3]);
    p = fma(r, p, c_fm[4]);
    p = fma(r, p, c_fm[5]);
    const double l1p = fma(r * r, p, r);
    const double ed = (double)e;
    // ln2 = hi + lo, hi has 11 trailing zero bits so that e * hi is exact
    return fma(ed, c_fm[6], (t.y + l1p) + ed * c_fm[7]);
}

__device__ __forceinline__ double fm_exp(double y) {
    const double kMagic = c_fm[9];                                // 1.5 * 2^52: round to nearest integer
    const double z = fma(y, c_fm[8] /* 64 / ln2 */, kMagic);
    const int ki = __double2loint(z);
    const double kd = z - kMagic;
    double r = fma(kd, c_fm[10] /* -ln2/64 hi, 21 trailing zero bits */, y);
    r = fma(kd, c_fm[11] /* -ln2/64 lo */, r);
    const double T = s_fm_exp[ki & 63];
    double p = fma(r, c_fm[12], c_fm[13]);
    p = fma(r, p, c_fm[14]);
    p = fma(r, p, c_fm[15]);
    p = fma(r, p, c_fm[16]);
    p = fma(r * r, p, r);                                         // exp(r) - 1
    const double v = fma(T, p, T);
    return __hiloint2double(__double2hiint(v) + ((ki >> 6) << 20), __double2loint(v));
}

// y = 1 / sqrt(x)
__device__ __forceinline__ double fm_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double h = x * y;
    double t = fma(-h, y, 1.0);
    y = fma(0.5 * y, t, y);
    h = x * y;
    t = fma(-h, y, 1.0);
    y = fma(0.5 * y, t, y);
    h = x * y;
    t = fma(-h, y, 1.0);
    y = fma(0.5 * y, t, y);
    return y;
}

// s = sqrt(x) from y ~ 1 / sqrt(x), one correction step
__device__ __forceinline__ double fm_sqrt_from_rsqrt(double x, double y) {
    const double s = x * y;
    const double e = fma(-s, s, x);
    return fma(e, 0.5 * y, s);
}

// arguments the fast paths accept: positive, normal, far from the exponent limits
__device__ __forceinline__ bool fm_ok(double x) { return x > 1e-280 && x < 1e280; }
