// 1-D cubic Hermite lookup on a (uniform) table with finite-difference slopes: restates interpolate()
// (functional_tools.py:292-334) per query point.  Shared by the Huang-Carter kernel omega(eta) (hc.cu) and the
// local pseudopotential v(|k|) of the ionic potential (ions.cu).
#pragma once
#include "common.cuh"

struct UniformTable {
    const double* eta;     // n abscissae, uniform
    const double* w;       // n values
    const double* m;       // n Hermite slopes (functional_tools.py:309-310)
    int n;
    double eta_max, inv_d; // last abscissa, 1 / spacing
};

// end slopes one-sided, interior slopes = mean of the adjacent secants
static __global__ void k_table_slopes(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ m, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto sec = [&](int j) { return (y[j + 1] - y[j]) / (x[j + 1] - x[j]); };
    double v;
    if (i == 0) v = sec(0);
    else if (i == n - 1) v = sec(n - 2);
    else v = 0.5 * (sec(i) + sec(i - 1));
    m[i] = v;
}

__device__ __forceinline__ void hermite(double t, double& h00, double& h10, double& h01, double& h11) {
    const double t2 = t * t, t3 = t2 * t;
    h00 = 1.0 - 3.0 * t2 + 2.0 * t3; h10 = t - 2.0 * t2 + t3; h01 = 3.0 * t2 - 2.0 * t3; h11 = t3 - t2;
}

// value at eta (clamped to the table end); interval = searchsorted(x[1:], eta) (left), as the reference
__device__ __forceinline__ double table_lookup(const UniformTable& T, double eta) {
    eta = fmin(eta, T.eta_max);
    int i = (int)ceil(eta * T.inv_d) - 1;
    i = max(0, min(i, T.n - 2));
    while (i < T.n - 2 && T.eta[i + 1] < eta) ++i;
    while (i > 0 && T.eta[i] >= eta) --i;
    const double x0 = T.eta[i], dx = T.eta[i + 1] - x0;
    double h00, h10, h01, h11;
    hermite((eta - x0) / dx, h00, h10, h01, h11);
    return h00 * T.w[i] + h10 * T.m[i] * dx + h01 * T.w[i + 1] + h11 * T.m[i + 1] * dx;
}

// value and d/d eta of the interpolant; the clamp makes the slope zero beyond the table end (torch.minimum)
__device__ __forceinline__ void table_lookup_slope(const UniformTable& T, double eta, double& val, double& slope) {
    const bool beyond = eta > T.eta_max;
    eta = fmin(eta, T.eta_max);
    int i = (int)ceil(eta * T.inv_d) - 1;
    i = max(0, min(i, T.n - 2));
    while (i < T.n - 2 && T.eta[i + 1] < eta) ++i;
    while (i > 0 && T.eta[i] >= eta) --i;
    const double x0 = T.eta[i], dx = T.eta[i + 1] - x0;
    const double t = (eta - x0) / dx, t2 = t * t;
    double h00, h10, h01, h11;
    hermite(t, h00, h10, h01, h11);
    const double y0 = T.w[i], y1 = T.w[i + 1], m0 = T.m[i] * dx, m1 = T.m[i + 1] * dx;
    val = h00 * y0 + h10 * m0 + h01 * y1 + h11 * m1;
    slope = beyond ? 0.0
                   : ((-6.0 * t + 6.0 * t2) * y0 + (1.0 - 4.0 * t + 3.0 * t2) * m0 + (6.0 * t - 6.0 * t2) * y1 + (3.0 * t2 - 2.0 * t) * m1) / dx;
}
