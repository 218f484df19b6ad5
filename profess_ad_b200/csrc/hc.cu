// Huang-Carter / revised Huang-Carter functionals: field-dependent convolution by a spline over
// xi-nodes, with the analytic potential.
//
// Replaces HuangCarter.forward / RevisedHuangCarter.forward (functionals.py:1232-1269, 1331-1365),
// field_dependent_convolution (functional_tools.py:381-423), interpolate_kernel (:337-378) and the
// 1-D Hermite table lookup interpolate (:292-334).
//
//   xi(r)   = 2 kF(n) (1 + lambda |grad n|^2 / (n^{8/3} + 1e-30))              (HC)
//           = 2 kF(n) (1 + a s^2 / (1 + b s^2)),  s^2 = |grad n|^2 / (4 (3 pi^2)^{2/3} n^{8/3})   (revHC)
//   K(r)    = sum_j w_j(xi(r)) [omega(|k| / xi_j) * n^beta](r)      cubic Hermite in xi over the nodes xi_j
//   T_NL    = C_HC  int n^{8/3 - beta} K / xi^3
//
// Potential, with F = n^{8/3-beta} / xi^3 and W_j = F w_j (at most four
// nodes per voxel carry weight):
//   v_NL = C_HC [ dF/dn|_xi K  +  beta n^{beta-1} irfft( sum_j omega_j rfft(W_j) )
//                + E_xi xi_n  - 2 div(E_xi xi_sigma grad n) ],    E_xi = -3 F K / xi + F dK/dxi
// FFT count: 12 + 2 n_xi including the 2 of vW.
//
// Stress (pad_stress_hc_nl, formula validated against autograd in tests/analytic_model.py:stress_huang_carter_nonlocal):
//   sigma_ab = (1/vol) [ delta_ab (E_NL - int v_NL n) - sum_r (dE/dg_a) g_b
//                        - sum_j sum_k w (omega'(|k|/xi_j) / xi_j) (k_a k_b / |k|) Re[conj(W_j^) g^] / N ]
// with g = grad n, dE/dg_a = 2 E_xi xi_sigma g_a dV, W_j = C_HC dV F w_j, g^ = rfftn(n^beta): the evaluation below run
// once more with three reductions hooked in.
//
// The node list depends on min/max of xi, which the reference reads on the host
// (functional_tools.py:408-416); this entry point does the same single device->host read of two doubles.
#include <vector>

#include "common.cuh"
#include "table.cuh"

#define HC_MAX_NODES 512

namespace {

constexpr double k3Pi2 = 29.608813203268074;
constexpr double kCHC = 680.1106493955546;          // 0.3 (3 pi^2)^{2/3} * 8 * 3 pi^2
constexpr double kCS = 0.026121172985233605;        // (1/4)(3 pi^2)^{-2/3}

typedef UniformTable HcTable;

__device__ __forceinline__ double kabs3(double kx, double ky, double kz) {
    const double k2 = kx * kx + ky * ky + kz * kz;
    return k2 != 0.0 ? sqrt(k2) : 0.0;
}

// xi and its partial derivatives w.r.t. n and sigma = |grad n|^2
__device__ __forceinline__ void xi_of(int variant, double p0, double p1, double n, double sig, double& xi, double& xi_n,
                                      double& xi_s) {
    const double c13 = cbrt(n);
    const double kF2 = 2.0 * cbrt(k3Pi2) * c13;            // 2 kF(n)
    const double dkF2 = kF2 / (3.0 * n);
    const double r83 = n * n * c13 * c13;
    if (variant == 0) {
        const double den = r83 + 1e-30;
        const double s2 = sig / den;
        const double g = 1.0 + p0 * s2;
        xi = kF2 * g;
        const double ds2_n = -sig * (8.0 / 3.0) * (r83 / n) / (den * den);
        xi_n = dkF2 * g + kF2 * p0 * ds2_n;
        xi_s = kF2 * p0 / den;
    } else {
        const double s2 = kCS * sig / r83;
        const double q = 1.0 + p1 * s2;
        const double g = 1.0 + p0 * s2 / q;
        const double dg = p0 / (q * q);
        xi = kF2 * g;
        xi_n = dkF2 * g + kF2 * dg * (-(8.0 / 3.0) * s2 / n);
        xi_s = kF2 * dg * kCS / r83;
    }
}

__global__ void __launch_bounds__(PAD_THREADS) k_xi_minmax(size_t n, const double* __restrict__ den, const double* __restrict__ gx,
                                                         const double* __restrict__ gy, const double* __restrict__ gz, int variant,
                                                         double p0, double p1, double* __restrict__ xi_out,
                                                         unsigned long long* minmax /*[0]=~min bits, [1]=max bits*/) {
    double lo = INFINITY, hi = 0.0;
    const size_t stride = (size_t)gridDim.x * PAD_THREADS;
    for (size_t i = (size_t)blockIdx.x * PAD_THREADS + threadIdx.x; i < n; i += stride) {
        const double a = gx[i], b = gy[i], c = gz[i];
        double xi, xn, xs;
        xi_of(variant, p0, p1, den[i], a * a + b * b + c * c, xi, xn, xs);
        xi_out[i] = xi;
        lo = fmin(lo, xi);
        hi = fmax(hi, xi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // positive doubles order like their bit patterns; min via max of the complement
        atomicMax(&minmax[0], ~(unsigned long long)__double_as_longlong(lo));
        atomicMax(&minmax[1], (unsigned long long)__double_as_longlong(hi));
    }
}

// interval index and Hermite data for xi among the nodes: x[i] < xi <= x[i+1]  (searchsorted(x[1:], xi), left)
__device__ __forceinline__ int node_interval(const double* __restrict__ x, int nn, double xi) {
    int lo = 0, hi = nn - 1;          // find first j in [1, nn-1] with x[j] >= xi, minus 1
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x[mid] < xi) lo = mid; else hi = mid;
    }
    return min(lo, nn - 2);
}

struct NodeWeights {
    int i;                 // interval
    double w[4];           // weights of conv_{i-1..i+2} in K
    double dw[4];          // d/dxi of those weights
};

__device__ __forceinline__ NodeWeights node_weights(const double* __restrict__ x, int nn, double xi) {
    NodeWeights r;
    const int i = node_interval(x, nn, xi);
    r.i = i;
    const double dx = x[i + 1] - x[i];
    const double t = (xi - x[i]) / dx;
    double h00, h10, h01, h11;
    hermite(t, h00, h10, h01, h11);
    const double t2 = t * t;
    const double g00 = (-6.0 * t + 6.0 * t2) / dx, g10 = (1.0 - 4.0 * t + 3.0 * t2) / dx, g01 = (6.0 * t - 6.0 * t2) / dx,
                 g11 = (3.0 * t2 - 2.0 * t) / dx;
    // slopes m_i = a sec_{i-1} + b sec_i,  m_{i+1} = c sec_i + e sec_{i+1}
    const double a = i == 0 ? 0.0 : 0.5, b = i == 0 ? 1.0 : 0.5;
    const double c = i + 1 == nn - 1 ? 1.0 : 0.5, e = i + 1 == nn - 1 ? 0.0 : 0.5;
    const double dm = i == 0 ? 1.0 : x[i] - x[i - 1];
    const double dp = i + 1 == nn - 1 ? 1.0 : x[i + 2] - x[i + 1];
    auto weights = [&](double H00, double H10, double H01, double H11, double* w) {
        w[0] = -H10 * dx * a / dm;
        w[1] = H00 + H10 * dx * (a / dm - b / dx) - H11 * dx * c / dx;
        w[2] = H01 + H10 * dx * b / dx + H11 * dx * (c / dx - e / dp);
        w[3] = H11 * dx * e / dp;
    };
    weights(h00, h10, h01, h11, r.w);
    weights(g00, g10, g01, g11, r.dw);
    return r;
}

// slab plans: mm[0] holds ~bits(min xi) and mm[1] bits(max xi); both order like non-negative doubles only after the
// complement is undone, so the exchange goes through (-min, max) as doubles
__global__ void k_mm_pack(unsigned long long* mm, double* scratch, int to_scratch) {
    if (to_scratch) {
        scratch[0] = -__longlong_as_double((long long)(~mm[0]));
        scratch[1] = __longlong_as_double((long long)mm[1]);
    } else {
        mm[0] = ~(unsigned long long)__double_as_longlong(-scratch[0]);
        mm[1] = (unsigned long long)__double_as_longlong(scratch[1]);
    }
}

int pad_allreduce_max_bits_raw(pad_plan* p, unsigned long long* mm, cudaStream_t s) {
    if (!p->dist) return PAD_OK;
    k_mm_pack<<<1, 1, 0, s>>>(mm, p->comm_scratch, 1);
    PAD_TRY(pad_slab_comm(p, PAD_COMM_ALL_REDUCE_MAX, 2, s));
    k_mm_pack<<<1, 1, 0, s>>>(mm, p->comm_scratch, 0);
    g_pad_launches += 2;
    PAD_CUDA(cudaGetLastError());
    return PAD_OK;
}

}  // namespace

__global__ void k_add_iso(double* sig, const double* E, double c) {
    const double v = c * E[0];
    sig[0] += v; sig[4] += v; sig[8] += v;
}

// sig != null: stress mode -- only the non-local term is evaluated (E_out / v_out must then be scratch: a device double and
// a ZEROED field), and its stress is ADDED to the 9 device doubles sig
static int hc_impl(pad_plan* p, const double* den, int variant, double p0, double p1, double beta, double kappa,
                   int geometric, const double* table_dev, int n_eta, double* E_out, double* v_out,
                   int accumulate, int* n_nodes_out, void* stream, double* sig) {
    if (!p || !den || !table_dev) { pad_set_error("pad_eval_hc: null argument"); return PAD_ERR_ARG; }
    if (variant != 0 && variant != 1) { pad_set_error("pad_eval_hc: variant must be 0 (HC) or 1 (revHC)"); return PAD_ERR_ARG; }
    if (n_eta < 3) { pad_set_error("pad_eval_hc: kernel table too short"); return PAD_ERR_ARG; }
    if (geometric && !(kappa > 1.0)) { pad_set_error("pad_eval_hc: kappa > 1 required for the geometric node progression"); return PAD_ERR_ARG; }
    PAD_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t N = p->N, nk = p->Nk;
    const size_t Ns = (N + 31) & ~(size_t)31;     // node-field stride: cuFFT operands must stay 16-byte aligned
    const int grid = pad_grid_for(N), gridk = pad_grid_for(nk);
    const double inv_n = p->geom.inv_n;
    const KGeom geom = p->geom;

    // TF + vW first (they also own rbuf 0/cbuf 0 while they run)
    if (!sig) PAD_TRY(pad_eval_wt(p, den, 1.0, 1.0, PAD_PART_TF | PAD_PART_VW, E_out, v_out, accumulate, stream));

    double* R[6];
    cufftDoubleComplex* C[4];
    for (int i = 0; i < 6; ++i) PAD_TRY(pad_get_rbuf(p, i, &R[i]));
    for (int i = 0; i < 4; ++i) PAD_TRY(pad_get_cbuf(p, i, &C[i]));
    double *Gx = R[0], *Gy = R[1], *Gz = R[2], *Xi = R[3], *Gb = R[4], *Ex = R[5];

    // ---- table slopes, grad n, xi and its range --------------------------------------------------
    if (!p->hc_scratch) {
        PAD_CUDA(cudaMalloc(&p->hc_scratch, sizeof(double) * (HC_MAX_NODES + 8)));
        p->bytes_allocated += sizeof(double) * (HC_MAX_NODES + 8);
    }
    if (p->hc_slopes_n < n_eta) {
        if (p->hc_slopes) cudaFree(p->hc_slopes);
        PAD_CUDA(cudaMalloc(&p->hc_slopes, sizeof(double) * n_eta));
        p->hc_slopes_n = n_eta;
    }
    k_table_slopes<<<(n_eta + 255) / 256, 256, 0, s>>>(table_dev, table_dev + n_eta, p->hc_slopes, n_eta);
    ++g_pad_launches;
    PAD_TRY(pad_gradient(p, den, Gx, Gy, Gz, stream));
    unsigned long long* mm = reinterpret_cast<unsigned long long*>(p->hc_scratch + HC_MAX_NODES);
    PAD_CUDA(cudaMemsetAsync(mm, 0, 2 * sizeof(unsigned long long), s));
    k_xi_minmax<<<grid, PAD_THREADS, 0, s>>>(N, den, Gx, Gy, Gz, variant, p0, p1, Xi, mm);
    ++g_pad_launches;
    unsigned long long mm_h[2];
    double eta_ends[2];
    // slab plans: the node list must be the same on every rank -> global min / max (both are maxima of bit patterns)
    PAD_TRY(pad_allreduce_max_bits_raw(p, mm, s));
    PAD_CUDA(cudaMemcpyAsync(mm_h, mm, sizeof(mm_h), cudaMemcpyDeviceToHost, s));
    PAD_CUDA(cudaMemcpyAsync(&eta_ends[0], table_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
    PAD_CUDA(cudaMemcpyAsync(&eta_ends[1], table_dev + n_eta - 1, sizeof(double), cudaMemcpyDeviceToHost, s));
    PAD_CUDA(cudaStreamSynchronize(s));        // the one host read the reference also does (functional_tools.py:408-416)
    double xi_min, xi_max;
    {
        unsigned long long b0 = ~mm_h[0], b1 = mm_h[1];
        memcpy(&xi_min, &b0, 8);
        memcpy(&xi_max, &b1, 8);
    }
    if (!(xi_min > 0.0) || !isfinite(xi_max)) { pad_set_error("pad_eval_hc: xi field is not positive/finite (min %g, max %g)", xi_min, xi_max); return PAD_ERR_ARG; }

    // ---- node list (functional_tools.py:406-417) ---------------------------------------------------
    std::vector<double> nodes;
    if (!geometric) {
        const double lower = (floor(xi_min / kappa) - 3.0) * kappa, upper = (ceil(xi_max / kappa) + 3.0) * kappa;
        const long cnt = (long)ceil((upper - lower) / kappa);       // torch.arange(lower, upper, kappa)
        for (long j = 0; j < cnt; ++j) {
            double v = lower + (double)j * kappa;
            if (v == 0.0) v = xi_min;
            nodes.push_back(v);
        }
    } else {
        const double lower = pow(kappa, -(ceil(-log(xi_min) / log(kappa)) + 3.0));
        const long cnt = (long)(ceil(log((xi_max + 1.0) / lower) / log(kappa)) + 3.0);
        for (long j = 0; j < cnt; ++j) nodes.push_back(lower * pow(kappa, (double)j));
    }
    const int nn = (int)nodes.size();
    if (nn < 2 || nn > HC_MAX_NODES) { pad_set_error("pad_eval_hc: %d xi-nodes (supported: 2..%d); increase kappa", nn, HC_MAX_NODES); return PAD_ERR_ARG; }
    if (n_nodes_out) *n_nodes_out = nn;
    double* nodes_dev = p->hc_scratch;
    PAD_CUDA(cudaMemcpyAsync(nodes_dev, nodes.data(), sizeof(double) * nn, cudaMemcpyHostToDevice, s));
    if (p->hc_conv_nodes < nn) {
        if (p->hc_conv) { cudaFree(p->hc_conv); p->bytes_allocated -= sizeof(double) * (size_t)p->hc_conv_nodes * Ns; }
        PAD_CUDA(cudaMalloc(&p->hc_conv, sizeof(double) * (size_t)nn * Ns));
        p->hc_conv_nodes = nn;
        p->bytes_allocated += sizeof(double) * (size_t)nn * Ns;
    }
    double* conv = p->hc_conv;
    HcTable T;
    T.eta = table_dev; T.w = table_dev + n_eta; T.m = p->hc_slopes; T.n = n_eta;
    T.eta_max = eta_ends[1];
    T.inv_d = (double)(n_eta - 1) / (eta_ends[1] - eta_ends[0]);

    // ---- g = n^beta, its spectrum, one convolution per node --------------------------------------
    {
        auto f = [=] __device__(size_t i, double(&)[1]) { Gb[i] = exp(beta * log(den[i])); };
        ew_kernel<0, decltype(f)><<<grid, PAD_THREADS, 0, s>>>(N, f, p->partials);
        ++g_pad_launches;
    }
    PAD_TRY(pad_fft_forward(p, Gb, C[0], s));
    const cufftDoubleComplex* G = C[0];
    for (int j = 0; j < nn; ++j) {
        cufftDoubleComplex* Cj = C[1];
        const double xi_j = nodes[j];
        auto f = [=] __device__(uint32_t idx, const KPoint& k) {
            const double m = inv_n * sym_even(k, [&](double kx, double ky, double kz) { return table_lookup(T, kabs3(kx, ky, kz) / xi_j); });
            const cufftDoubleComplex c = G[idx];
            Cj[idx] = make_cuDoubleComplex(c.x * m, c.y * m);
        };
        ks_kernel<decltype(f)><<<gridk, PAD_THREADS, 0, s>>>(geom, (uint32_t)nk, f);
        ++g_pad_launches;
        PAD_TRY(pad_fft_inverse(p, Cj, conv + (size_t)j * Ns, s));
    }

    // ---- K, energy density, local parts of the potential, E_xi, weights W_j ----------------------
    const bool want_v = v_out != nullptr;
    {
        auto f = [=] __device__(size_t i, double(&acc)[1]) {
            const double n = den[i], xi = Xi[i];
            const NodeWeights nw = node_weights(nodes_dev, nn, xi);
            double K = 0.0, dK = 0.0, cv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = nw.i - 1 + q;
                cv[q] = (j >= 0 && j < nn && (nw.w[q] != 0.0 || nw.dw[q] != 0.0)) ? conv[(size_t)j * Ns + i] : 0.0;
                K += nw.w[q] * cv[q];
                dK += nw.dw[q] * cv[q];
            }
            const double ln = log(n);
            const double pw = exp((8.0 / 3.0 - beta) * ln);
            const double xi3 = xi * xi * xi;
            const double F = pw / xi3;
            acc[0] += F * K;
            if (want_v) {
                v_out[i] += kCHC * (8.0 / 3.0 - beta) * pw / n / xi3 * K;
                Ex[i] = kCHC * (-3.0 * F * K / xi + F * dK);
                Gb[i] = kCHC * beta * exp(beta * ln) / n;        // beta n^{beta-1}, applied to the adjoint convolution
            }
        };
        ew_kernel<1, decltype(f)><<<grid, PAD_THREADS, 0, s>>>(N, f, p->partials);
        ++g_pad_launches;
    }
    {
        FinalizeArgs a;
        a.nblocks = grid; a.nterms = 1; a.accumulate = 1;
        for (int t = 0; t < PAD_MAX_RED; ++t) a.coef[t] = 0.0;
        a.coef[0] = kCHC * p->dV;
        a.sums_out = nullptr;
        a.E_out = E_out;
        if (E_out) pad_launch_finalize(p, a, s);
    }
    PAD_CUDA(cudaGetLastError());
    if (!want_v) return PAD_OK;

    // ---- adjoint convolutions: sum_j omega_j rfft(W_j), W_j = F w_j ----------------------------------
    PAD_CUDA(cudaMemsetAsync(C[2], 0, sizeof(cufftDoubleComplex) * nk, s));
    // every conv_j is dead once K and dK/dxi are formed, so W_j is written in its place -- all n_xi fields in ONE sweep:
    // the node search, the Hermite weights and n^(8/3 - beta) / xi^3 are evaluated once per voxel instead of once per node
    {
        auto f = [=] __device__(size_t i, double(&)[1]) {
            const double n = den[i], xi = Xi[i];
            const NodeWeights nw = node_weights(nodes_dev, nn, xi);
            const double F = exp((8.0 / 3.0 - beta) * log(n)) / (xi * xi * xi);
            for (int j = 0; j < nn; ++j) {
                const int q = j - (nw.i - 1);
                const double w = q == 0 ? nw.w[0] : q == 1 ? nw.w[1] : q == 2 ? nw.w[2] : q == 3 ? nw.w[3] : 0.0;
                conv[(size_t)j * Ns + i] = w != 0.0 ? w * F : 0.0;
            }
        };
        ew_kernel<0, decltype(f)><<<grid, PAD_THREADS, 0, s>>>(N, f, p->partials);
        ++g_pad_launches;
    }
    for (int j = 0; j < nn; ++j) {
        double* Wj = conv + (size_t)j * Ns;
        PAD_TRY(pad_fft_forward(p, Wj, C[1], s));
        cufftDoubleComplex *Cj = C[1], *Acc = C[2];
        const double xi_j = nodes[j];
        if (sig) {      // - sum_k w (omega'(eta_j) / xi_j) (k_a k_b / |k|) Re[conj(W_j^) g^] C_HC dV / (N vol)
            const cufftDoubleComplex* Gs = C[0];
            auto fs = [=] __device__(size_t i, double(&acc)[7]) {
                const KPoint k = make_kpoint(geom, (uint32_t)i);
                const double ka = kabs3(k.kx, k.ky, k.kz);
                if (ka == 0.0) return;
                double val, slope;
                table_lookup_slope(T, ka / xi_j, val, slope);
                const cufftDoubleComplex w = Cj[i], gq = Gs[i];
                const bool edge = k.j2 == 0 || (geom.e2 && k.j2 == geom.n2 / 2);
                const double t = (edge ? 1.0 : 2.0) * slope / xi_j * (w.x * gq.x + w.y * gq.y) / ka;
                acc[1] += t * k.kx * k.kx; acc[2] += t * k.ky * k.ky; acc[3] += t * k.kz * k.kz;
                acc[4] += t * k.kx * k.ky; acc[5] += t * k.kx * k.kz; acc[6] += t * k.ky * k.kz;
            };
            ew_kernel<7, decltype(fs)><<<gridk, PAD_THREADS, 0, s>>>(nk, fs, p->partials);
            ++g_pad_launches;
            PAD_TRY(pad_stress_accumulate(p, s, gridk, 0.0, -kCHC * inv_n * inv_n, sig));
        }
        auto fk = [=] __device__(uint32_t idx, const KPoint& k) {
            const double m = inv_n * sym_even(k, [&](double kx, double ky, double kz) { return table_lookup(T, kabs3(kx, ky, kz) / xi_j); });
            const cufftDoubleComplex c = Cj[idx];
            cufftDoubleComplex a = Acc[idx];
            a.x += c.x * m; a.y += c.y * m;
            Acc[idx] = a;
        };
        ks_kernel<decltype(fk)><<<gridk, PAD_THREADS, 0, s>>>(geom, (uint32_t)nk, fk);
        ++g_pad_launches;
    }
    double* Adj;
    PAD_TRY(pad_get_rbuf(p, 6, &Adj));
    PAD_TRY(pad_fft_inverse(p, C[2], Adj, s));

    // ---- xi-dependence: E_xi xi_n - 2 div(E_xi xi_sigma grad n) ------------------------------------
    if (sig) {      // - sum_r (dE/dg_a) g_b / vol,  dE/dg_a = 2 E_xi xi_sigma g_a dV
        auto f = [=] __device__(size_t i, double(&acc)[7]) {
            const double gx = Gx[i], gy = Gy[i], gz = Gz[i];
            double xi, xn, xs;
            xi_of(variant, p0, p1, den[i], gx * gx + gy * gy + gz * gz, xi, xn, xs);
            const double w = 2.0 * Ex[i] * xs;
            acc[1] += w * gx * gx; acc[2] += w * gy * gy; acc[3] += w * gz * gz;
            acc[4] += w * gx * gy; acc[5] += w * gx * gz; acc[6] += w * gy * gz;
        };
        ew_kernel<7, decltype(f)><<<grid, PAD_THREADS, 0, s>>>(N, f, p->partials);
        ++g_pad_launches;
        PAD_TRY(pad_stress_accumulate(p, s, grid, 0.0, -inv_n, sig));
    }
    {
        auto f = [=] __device__(size_t i, double(&)[1]) {
            const double gx = Gx[i], gy = Gy[i], gz = Gz[i];
            double xi, xn, xs;
            xi_of(variant, p0, p1, den[i], gx * gx + gy * gy + gz * gz, xi, xn, xs);
            const double ex = Ex[i];
            v_out[i] += ex * xn + Gb[i] * Adj[i];
            const double w = 2.0 * ex * xs;
            Gx[i] = w * gx; Gy[i] = w * gy; Gz[i] = w * gz;
        };
        ew_kernel<0, decltype(f)><<<grid, PAD_THREADS, 0, s>>>(N, f, p->partials);
        ++g_pad_launches;
    }
    PAD_TRY(pad_fft_forward_many(p, R, C, 3, s));
    {
        cufftDoubleComplex *C0 = C[0], *C1 = C[1], *C2 = C[2];
        auto fk = [=] __device__(uint32_t idx, const KPoint& k) {
            double kx, ky, kz;
            sym_kvec(k, kx, ky, kz);
            const cufftDoubleComplex a = C0[idx], b = C1[idx], c = C2[idx];
            const double re = kx * a.x + ky * b.x + kz * c.x, im = kx * a.y + ky * b.y + kz * c.y;
            C0[idx] = make_cuDoubleComplex(-im * inv_n, re * inv_n);
        };
        ks_kernel<decltype(fk)><<<gridk, PAD_THREADS, 0, s>>>(geom, (uint32_t)nk, fk);
        ++g_pad_launches;
    }
    PAD_TRY(pad_fft_inverse(p, C[0], R[0], s));
    {
        const double* Dv = R[0];
        auto f = [=] __device__(size_t i, double(&)[1]) { v_out[i] -= Dv[i]; };
        ew_kernel<0, decltype(f)><<<grid, PAD_THREADS, 0, s>>>(N, f, p->partials);
        ++g_pad_launches;
    }
    if (sig) {      // delta_ab (E_NL - int v_NL n) / vol
        auto f = [=] __device__(size_t i, double(&acc)[7]) { acc[0] += v_out[i] * den[i]; };
        ew_kernel<7, decltype(f)><<<grid, PAD_THREADS, 0, s>>>(N, f, p->partials);
        ++g_pad_launches;
        PAD_TRY(pad_stress_accumulate(p, s, grid, -inv_n, 0.0, sig));
        k_add_iso<<<1, 1, 0, s>>>(sig, E_out, 1.0 / p->vol);
        ++g_pad_launches;
    }
    PAD_CUDA(cudaGetLastError());
    return PAD_OK;
}

extern "C" int pad_eval_hc(pad_plan* p, const double* den, int variant, double p0, double p1, double beta, double kappa,
                           int geometric, const double* table_dev, int n_eta, double* E_out, double* v_out,
                           int accumulate, int* n_nodes_out, void* stream) {
    return hc_impl(p, den, variant, p0, p1, beta, kappa, geometric, table_dev, n_eta, E_out, v_out, accumulate, n_nodes_out, stream,
                   nullptr);
}

// stress of the non-local Huang-Carter term, added to sig[9] (device).  Scratch: real field 7 (v_NL), scalar slot S_TMP0 + 4.
int pad_stress_hc_nl(pad_plan* p, const double* den, int variant, double p0, double p1, double beta, double kappa, int geometric,
                     const double* table_dev, int n_eta, double* sig, cudaStream_t s) {
    double* vtmp;
    PAD_TRY(pad_get_rbuf(p, 7, &vtmp));
    PAD_CUDA(cudaMemsetAsync(vtmp, 0, sizeof(double) * p->N, s));
    double* Etmp = p->scal + S_TMP0 + 4;
    PAD_CUDA(cudaMemsetAsync(Etmp, 0, sizeof(double), s));
    return hc_impl(p, den, variant, p0, p1, beta, kappa, geometric, table_dev, n_eta, Etmp, vtmp, 1, nullptr, (void*)s, sig);
}
