// Ion-ion interaction: real-space damped pairwise electrostatic sum in a neutralising background
// (Pickard, Phys. Rev. Materials 2, 013806), restating ion_interaction_sum (ion_utils.py:293-333) with its
// derivatives in closed form instead of a materialised pair list + autograd:
//
//   E = 1/2 sum_i Z_i sum_{(j,s): 0 < r < Rc} Z_j erfc(r / Rd) / r  +  sum_i T_i,        r = |r_j + s B - r_i|
//   T_i = -pi Z_i rho Ra_i^2 + pi Z_i rho (Ra_i^2 - Rd^2 / 2) erf(Ra_i / Rd) + sqrt(pi) Z_i rho Ra_i Rd exp(-Ra_i^2 / Rd^2)
//         - Z_i^2 / (sqrt(pi) Rd),     rho = sum Z / vol,   Ra_i = (3 Q_i / (4 pi rho))^(1/3),   Q_i = Z_i + sum_{(j,s)} Z_j
//
// The reference builds the (i, j, shift) list with torch_nl and lets autograd differentiate; Q_i comes out of an index
// operation, so it carries no gradient there either.  Here one CTA per (ion, slice of the image shifts) sweeps the
// candidates (j, s) directly -- nothing is materialised -- and accumulates, next to the energy,
//   dE/dr_i        = -sum_{(j,s)} Z_i Z_j phi'(r) d / r                      (forces)
//   dE/dB_ka |_r   = 1/2 sum_i sum_{(j,s)} Z_i Z_j phi'(r) d_a s_k / r + dE_corr/dvol * vol * (B^-1)_ak
// (the partial derivative with respect to the lattice vectors at FIXED Cartesian coordinates: what torch.autograd needs
// for ion_interaction_sum(box_vecs, coords, ...) as a function of two tensors; the stress follows by the chain rule through
// coords = frac @ box_vecs), with phi(r) = erfc(r / Rd) / r and
//   dT_i/dvol = [ (T_i + Z_i^2 / (sqrt(pi) Rd)) + (2 pi / 3) Z_i rho Ra_i^2 erfc(Ra_i / Rd) ] / vol     (rho ~ 1/vol, Ra ~ vol^(1/3)).
// Sums are taken in a fixed order (thread-strided, tree reduction, slices in order): deterministic.
#include <vector>

#include "common.cuh"

namespace {

constexpr int kIonThreads = 256;
constexpr int kIonVals = 14;        // e, Q, f[3], g[9]

struct IonBox {
    double B[9];
    int r0, r1, r2;                 // image shifts -r_a .. r_a
    long long nshift;
};

__global__ void __launch_bounds__(kIonThreads) k_ion_pairs(IonBox bx, const double* __restrict__ cart, const double* __restrict__ Z,
                                                          int n, double Rc2, double inv_Rd, int nslice, int want_grad,
                                                          double* __restrict__ part /* [n][nslice][kIonVals] */) {
    const int i = blockIdx.x, sl = blockIdx.y;
    const double xi = cart[3 * i], yi = cart[3 * i + 1], zi = cart[3 * i + 2];
    const int w1 = 2 * bx.r1 + 1, w2 = 2 * bx.r2 + 1;
    const long long per = (bx.nshift + nslice - 1) / nslice;
    const long long s_begin = (long long)sl * per, s_end = s_begin + per < bx.nshift ? s_begin + per : bx.nshift;
    const long long total = (s_end > s_begin ? s_end - s_begin : 0) * n;
    double acc[kIonVals];
#pragma unroll
    for (int q = 0; q < kIonVals; ++q) acc[q] = 0.0;
    const double two_over_sqrt_pi = 1.1283791670955126;
    for (long long c = threadIdx.x; c < total; c += kIonThreads) {
        const long long sidx = s_begin + c / n;
        const int j = (int)(c - (c / n) * n);
        const int s0 = (int)(sidx / ((long long)w1 * w2)) - bx.r0;
        const int rem = (int)(sidx % ((long long)w1 * w2));
        const int s1 = rem / w2 - bx.r1, s2 = rem % w2 - bx.r2;
        if (j == i && s0 == 0 && s1 == 0 && s2 == 0) continue;
        const double tx = cart[3 * j] + ((double)s0 * bx.B[0] + (double)s1 * bx.B[3] + (double)s2 * bx.B[6]);
        const double ty = cart[3 * j + 1] + ((double)s0 * bx.B[1] + (double)s1 * bx.B[4] + (double)s2 * bx.B[7]);
        const double tz = cart[3 * j + 2] + ((double)s0 * bx.B[2] + (double)s1 * bx.B[5] + (double)s2 * bx.B[8]);
        const double dx = tx - xi, dy = ty - yi, dz = tz - zi;
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (!(r2 < Rc2)) continue;
        const double r = sqrt(r2), zj = Z[j];
        const double x = r * inv_Rd;
        const double ec = erfc(x), ir = 1.0 / r;
        acc[0] += zj * ec * ir;
        acc[1] += zj;
        if (want_grad) {
            // phi'(r) / r
            const double dphi = -(ec * ir + two_over_sqrt_pi * inv_Rd * exp(-x * x)) * ir * ir;
            const double w = zj * dphi;
            acc[2] += w * dx; acc[3] += w * dy; acc[4] += w * dz;
            const double a0 = w * (double)s0, a1 = w * (double)s1, a2 = w * (double)s2;
            acc[5] += a0 * dx; acc[6] += a0 * dy; acc[7] += a0 * dz;
            acc[8] += a1 * dx; acc[9] += a1 * dy; acc[10] += a1 * dz;
            acc[11] += a2 * dx; acc[12] += a2 * dy; acc[13] += a2 * dz;
        }
    }
    __shared__ double sm[kIonVals][kIonThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < kIonVals; ++q) {
        const double v = warp_sum(acc[q]);
        if (lane == 0) sm[q][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < kIonVals) {
        double v = 0.0;
        for (int w = 0; w < kIonThreads / 32; ++w) v += sm[threadIdx.x][w];
        part[((size_t)i * nslice + sl) * kIonVals + threadIdx.x] = v;
    }
}

// one CTA: per-ion totals over the slices (in order), correction terms, sums over the ions
__global__ void __launch_bounds__(kIonThreads) k_ion_finish(const double* __restrict__ part, const double* __restrict__ Z, int n, int nslice,
                                                           double Rd, double vol, double Ztot, IonBox bx, double binv0, double binv1,
                                                           double binv2, double binv3, double binv4, double binv5, double binv6,
                                                           double binv7, double binv8, double* __restrict__ E_out,
                                                           double* __restrict__ dcart /* n x 3 or null */,
                                                           double* __restrict__ dbox /* 9 or null */) {
    const double rho = Ztot / vol;
    const double sqrt_pi = 1.7724538509055160;
    double acc[11];       // E, g[9], dEcorr/dvol * vol
#pragma unroll
    for (int q = 0; q < 11; ++q) acc[q] = 0.0;
    for (int i = threadIdx.x; i < n; i += kIonThreads) {
        double tot[kIonVals];
#pragma unroll
        for (int q = 0; q < kIonVals; ++q) tot[q] = 0.0;
        for (int sl = 0; sl < nslice; ++sl)
#pragma unroll
            for (int q = 0; q < kIonVals; ++q) tot[q] += part[((size_t)i * nslice + sl) * kIonVals + q];
        const double z = Z[i];
        const double Q = z + tot[1];
        const double aux = (0.75 / kPi) * Q / rho;
        const double Ra = aux < 0.0 ? -cbrt(-aux) : cbrt(aux);
        const double x = Ra / Rd;
        const double ex = exp(-x * x), ef = erf(x);
        const double T = -kPi * z * rho * Ra * Ra + kPi * z * rho * (Ra * Ra - 0.5 * Rd * Rd) * ef + sqrt_pi * z * rho * Ra * Rd * ex -
                         z * z / (sqrt_pi * Rd);
        acc[0] += 0.5 * z * tot[0] + T;
        // dT/dvol * vol  (rho ~ 1/vol at fixed charges, Ra ~ vol^(1/3) at fixed Q)
        acc[10] += -(T + z * z / (sqrt_pi * Rd)) - (2.0 * kPi / 3.0) * z * rho * Ra * Ra * erfc(x);
        if (dcart) {
            // dE/dr_i = -Z_i sum Z_j phi'(r) d / r
            dcart[3 * i] = -z * tot[2]; dcart[3 * i + 1] = -z * tot[3]; dcart[3 * i + 2] = -z * tot[4];
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) acc[1 + q] += 0.5 * z * tot[5 + q];
    }
    __shared__ double sm[11][kIonThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 11; ++q) {
        const double v = warp_sum(acc[q]);
        if (lane == 0) sm[q][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[11];
        for (int q = 0; q < 11; ++q) {
            double v = 0.0;
            for (int w = 0; w < kIonThreads / 32; ++w) v += sm[q][w];
            t[q] = v;
        }
        E_out[0] = t[0];
        if (dbox) {
            // dE/dB_ka = G_ka + (dE_corr/dvol vol) (B^-1)_ak      (d vol / d B_ka = vol (B^-1)_ak)
            const double bi[9] = {binv0, binv1, binv2, binv3, binv4, binv5, binv6, binv7, binv8};
            for (int k = 0; k < 3; ++k)
                for (int a = 0; a < 3; ++a) dbox[3 * k + a] = t[1 + 3 * k + a] + t[10] * bi[3 * a + k];
        }
    }
}

int invert3h(const double* m, double* inv, double* det_out) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C;
    *det_out = det;
    if (det == 0.0 || !isfinite(det)) return 1;
    const double id = 1.0 / det;
    inv[0] = A * id;  inv[1] = -(b * i - c * h) * id; inv[2] = (b * f - c * e) * id;
    inv[3] = B * id;  inv[4] = (a * i - c * g) * id;  inv[5] = -(a * f - c * d) * id;
    inv[6] = C * id;  inv[7] = -(a * h - b * g) * id; inv[8] = (a * e - b * d) * id;
    return 0;
}

}  // namespace

// box_host: 9 doubles (rows = lattice vectors, bohr); cart_dev: n x 3 Cartesian coordinates; charges_dev: n; charge_total: sum of
// the charges (host).  E_out_dev: 1 double.  dcart_dev (n x 3) / dbox_dev (9): dE/dcoords and dE/dbox_vecs at fixed coords, or null.
// work_dev: at least pad_ion_ion_work_doubles(n, ...) doubles of device scratch.
extern "C" size_t pad_ion_ion_work_doubles(const double* box_host, int n, double Rc) {
    double inv[9], det;
    if (!box_host || n < 1 || invert3h(box_host, inv, &det)) return 0;
    long long ns = 1;
    for (int a = 0; a < 3; ++a) {
        // interplanar spacing h_a = 1 / |column a of B^-1|
        const double h = 1.0 / sqrt(inv[a] * inv[a] + inv[3 + a] * inv[3 + a] + inv[6 + a] * inv[6 + a]);
        ns *= 2 * ((long long)ceil(Rc / h) + 1) + 1;
    }
    long long nslice = (148LL * 8 + n - 1) / n;
    if (nslice > ns) nslice = ns;
    if (nslice < 1) nslice = 1;
    return (size_t)n * (size_t)nslice * kIonVals;
}

extern "C" int pad_ion_ion(const double* box_host, const double* cart_dev, const double* charges_dev, int n, double charge_total,
                           double Rc, double Rd, double* E_out_dev, double* dcart_dev, double* dbox_dev, double* work_dev, int device,
                           void* stream) {
    if (!box_host || !cart_dev || !charges_dev || !E_out_dev || !work_dev || n < 1) { pad_set_error("pad_ion_ion: bad argument"); return PAD_ERR_ARG; }
    if (!(Rc > 0.0) || !(Rd > 0.0)) { pad_set_error("pad_ion_ion: Rc and Rd must be positive"); return PAD_ERR_ARG; }
    double inv[9], det;
    if (invert3h(box_host, inv, &det)) { pad_set_error("Lattice vector matrix is not invertible."); return PAD_ERR_ARG; }
    PAD_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    IonBox bx;
    memcpy(bx.B, box_host, sizeof(double) * 9);
    int r[3];
    for (int a = 0; a < 3; ++a) {
        const double h = 1.0 / sqrt(inv[a] * inv[a] + inv[3 + a] * inv[3 + a] + inv[6 + a] * inv[6 + a]);
        r[a] = (int)ceil(Rc / h) + 1;
    }
    bx.r0 = r[0]; bx.r1 = r[1]; bx.r2 = r[2];
    bx.nshift = (long long)(2 * r[0] + 1) * (2 * r[1] + 1) * (2 * r[2] + 1);
    long long nslice = (148LL * 8 + n - 1) / n;
    if (nslice > bx.nshift) nslice = bx.nshift;
    if (nslice < 1) nslice = 1;
    if (nslice > 65535) nslice = 65535;
    const int want_grad = (dcart_dev || dbox_dev) ? 1 : 0;
    dim3 grid((unsigned)n, (unsigned)nslice);
    k_ion_pairs<<<grid, kIonThreads, 0, s>>>(bx, cart_dev, charges_dev, n, Rc * Rc, 1.0 / Rd, (int)nslice, want_grad, work_dev);
    k_ion_finish<<<1, kIonThreads, 0, s>>>(work_dev, charges_dev, n, (int)nslice, Rd, fabs(det), charge_total, bx, inv[0], inv[1], inv[2],
                                         inv[3], inv[4], inv[5], inv[6], inv[7], inv[8], E_out_dev, dcart_dev, dbox_dev);
    g_pad_launches += 2;
    PAD_CUDA(cudaGetLastError());
    return PAD_OK;
}
