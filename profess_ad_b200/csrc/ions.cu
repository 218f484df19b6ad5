// Ionic (local pseudopotential) potential and the ion-electron forces, straight from the ion positions.
//
// Replaces System.__potential_from_ions (system.py:183-205) -> interpolate_recpot (ion_utils.py:49-81) ->
// lattice_sum (ion_utils.py:88-118) -> structure_factor (ion_utils.py:121-137):
//     v_ext(r) = irfftn( sum_s v_s(|k|) S_s(k) ) / vol,      S_s(k) = sum_{I in s} exp(-i k.R_I)
// and the IonElectron part of System.__compute_forces (system.py:913-925), which the reference obtains by autograd
// through the same expression:
//     F_I = -dE/dR_I = -(dV / vol) sum_k w_k v_s(|k|) k Im[ exp(-i k.R_I) conj(rho_hat(k)) ]
// (w_k = 1 on the self-conjugate planes of the half spectrum, 2 elsewhere).
//
// The reference materialises an N_k x N_ion phase tensor (34.6 GB for 256 atoms on 256^3).  Here
// exp(-i k.R) = e0[j0] e1[j1] e2[j2] with k.R = 2 pi sum_a f_a u_a (u = fractional coordinates, f = integer
// frequencies), so each ion needs three 1-D phase tables (n0 + n1 + n2/2 + 1 complex numbers, sincospi of an exactly
// reduced argument) and a k-point costs one complex multiply-add per ion instead of a sincos.  A CTA owns one
// (j0, j1) row of the half spectrum: e0 e1 of 128 ions is staged in shared memory, the threads run along j2 where
// the e2 loads coalesce.  On the self-conjugate planes the "Nyquist made positive" spectrum is not Hermitian and the
// reference's irfftn keeps only its Hermitian part; the kernel forms that part explicitly
// from the partner wave vector, so any c2r transform reproduces the reference.
#include "common.cuh"
#include "table.cuh"

namespace {

constexpr int ION_T = 128;       // threads per CTA of the force kernel
constexpr int ION_RB = 8;        // half-spectrum rows per CTA sweep in the structure-factor kernel
constexpr int ION_CH = 128;      // ions staged per shared-memory chunk
constexpr int ION_IPC = 8;       // ions per CTA in the force kernel

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// E_a[ion][j] = exp(-2 pi i f_a(j) u_a[ion]); the product f u is reduced modulo 1 exactly (fma remainder) first
__global__ void __launch_bounds__(PAD_THREADS) k_ion_phases(const double* __restrict__ frac, int n_ions, int n0, int n1, int nzh,
                                                          double2* __restrict__ E0, double2* __restrict__ E1,
                                                          double2* __restrict__ E2) {
    const int per = n0 + n1 + nzh;
    const size_t total = (size_t)n_ions * per;
    for (size_t e = (size_t)blockIdx.x * PAD_THREADS + threadIdx.x; e < total; e += (size_t)gridDim.x * PAD_THREADS) {
        const int ion = (int)(e / per);
        int j = (int)(e - (size_t)ion * per);
        int axis;
        double f;
        double2* dst;
        if (j < n0) { axis = 0; f = (double)(j <= n0 / 2 ? j : j - n0); dst = E0 + (size_t)ion * n0 + j; }
        else if (j < n0 + n1) { j -= n0; axis = 1; f = (double)(j <= n1 / 2 ? j : j - n1); dst = E1 + (size_t)ion * n1 + j; }
        else { j -= n0 + n1; axis = 2; f = (double)j; dst = E2 + (size_t)ion * nzh + j; }
        const double u = frac[3 * ion + axis];
        const double t = f * u, lo = fma(f, u, -t);
        const double fr = (t - rint(t)) + lo;
        double sn, cs;
        sincospi(-2.0 * fr, &sn, &cs);
        *dst = make_double2(cs, sn);
    }
}

// v_s(|k|): Hermite interpolation of the tail-free table, Coulomb tail removed again (ion_utils.py:71-80)
__device__ __forceinline__ double recpot_value(const UniformTable& T, double z, double kx, double ky, double kz) {
    const double k2 = kx * kx + ky * ky + kz * kz;
    if (k2 == 0.0) return table_lookup(T, 0.0);
    return table_lookup(T, sqrt(k2)) - 4.0 * kPi * z / k2;
}

// One CTA owns ION_RB consecutive (j0, j1) rows at a time; its threads run along j2.  Per ion the e2 phase of a column
// is loaded once and used for all ION_RB rows (the e0 e1 products of ION_CH ions x ION_RB rows are staged in shared
// memory), so the traffic from L2 is 1 / ION_RB of a row-at-a-time sweep and the loop is FMA-bound.
__global__ void __launch_bounds__(256) k_ion_spectrum(KGeom g, int nrows, UniformTable T, double z, int n_ions,
                                                    const double2* __restrict__ E0, const double2* __restrict__ E1,
                                                    const double2* __restrict__ E2, double inv_vol, int accumulate,
                                                    int raw /* 1: plain S(k), no potential factor, no symmetrisation */,
                                                    double2* __restrict__ out) {
    __shared__ double2 s01[ION_RB][ION_CH], s01b[ION_RB][ION_CH];
    const int nzh = g.nzh;
    const bool nyq2_any = g.e2;
    for (int row0 = blockIdx.x * ION_RB; row0 < nrows; row0 += gridDim.x * ION_RB) {
        int j0r[ION_RB], j1r[ION_RB];
#pragma unroll
        for (int r = 0; r < ION_RB; ++r) {
            const int row = min(row0 + r, nrows - 1);
            j0r[r] = row / g.n1_loc;
            j1r[r] = row - j0r[r] * g.n1_loc + g.j1_off;
        }
        for (int jbase = 0; jbase < nzh; jbase += blockDim.x) {
            const int j2 = jbase + threadIdx.x;
            const bool live = j2 < nzh;
            const bool zedge0 = live && j2 == 0, zedgeN = live && nyq2_any && j2 == g.n2 / 2;
            const bool maybe_sp = !raw && (zedge0 || zedgeN);
            double2 acc[ION_RB], accb[ION_RB];
#pragma unroll
            for (int r = 0; r < ION_RB; ++r) { acc[r] = make_double2(0.0, 0.0); accb[r] = make_double2(0.0, 0.0); }
            for (int base = 0; base < n_ions; base += ION_CH) {
                __syncthreads();
                for (int e = threadIdx.x; e < ION_CH * ION_RB; e += blockDim.x) {
                    const int r = e / ION_CH, i = e - r * ION_CH, ion = base + i;
                    if (ion < n_ions) {
                        const int rr = min(row0 + r, nrows - 1);
                        const int jj0 = rr / g.n1_loc, jj1 = rr - jj0 * g.n1_loc + g.j1_off;
                        double2 a = E0[(size_t)ion * g.n0 + jj0], b = E1[(size_t)ion * g.n1 + jj1];
                        s01[r][i] = cmul(a, b);
                        if (!(g.e0 && jj0 == g.n0 / 2)) a.y = -a.y;   // partner wave vector: -f unless f is the (positive) Nyquist frequency
                        if (!(g.e1 && jj1 == g.n1 / 2)) b.y = -b.y;
                        s01b[r][i] = cmul(a, b);
                    }
                }
                __syncthreads();
                const int cnt = min(ION_CH, n_ions - base);
                if (live) {
                    const double2* __restrict__ e2 = E2 + (size_t)base * nzh + j2;
                    for (int i = 0; i < cnt; ++i) {
                        const double2 ph = e2[(size_t)i * nzh];
#pragma unroll
                        for (int r = 0; r < ION_RB; ++r) {
                            const double2 e = s01[r][i];
                            acc[r].x = fma(e.x, ph.x, fma(-e.y, ph.y, acc[r].x));
                            acc[r].y = fma(e.x, ph.y, fma(e.y, ph.x, acc[r].y));
                        }
                        if (maybe_sp) {
#pragma unroll
                            for (int r = 0; r < ION_RB; ++r) {
                                const double2 eb = s01b[r][i];
                                accb[r].x = fma(eb.x, ph.x, fma(-eb.y, ph.y, accb[r].x));
                                accb[r].y = fma(eb.x, ph.y, fma(eb.y, ph.x, accb[r].y));
                            }
                        }
                    }
                }
            }
            if (live) {
#pragma unroll
                for (int r = 0; r < ION_RB; ++r) {
                    if (row0 + r >= nrows) continue;
                    const KPoint p = make_kpoint_at(g, j0r[r], j1r[r], j2);
                    const double f = raw ? 1.0 : recpot_value(T, z, p.kx, p.ky, p.kz);
                    double2 G = make_double2(f * acc[r].x, f * acc[r].y);
                    if (p.special && !raw) {
                        const double fb = recpot_value(T, z, p.px, p.py, p.pz);
                        G.x = 0.5 * (G.x + fb * accb[r].x);
                        G.y = 0.5 * (G.y - fb * accb[r].y);
                    }
                    const size_t idx = (size_t)(row0 + r) * nzh + j2;
                    double2 o = accumulate ? out[idx] : make_double2(0.0, 0.0);
                    o.x += G.x * inv_vol;
                    o.y += G.y * inv_vol;
                    out[idx] = o;
                }
            }
        }
    }
}

// partial force sums of ION_IPC ions over one chunk of half-spectrum rows
__global__ void __launch_bounds__(ION_T) k_ion_force_partial(KGeom g, int nrows, int rows_per_chunk, int n_ions,
                                                           const double2* __restrict__ E0, const double2* __restrict__ E1,
                                                           const double2* __restrict__ E2, const double2* __restrict__ B,
                                                           double* __restrict__ partial /* [ion][chunk][3] */) {
    __shared__ double2 s01[ION_IPC];
    __shared__ double red[3 * ION_IPC][ION_T / 32];
    const int nzh = g.nzh;
    const int chunk = blockIdx.x, nchunks = gridDim.x, ion0 = blockIdx.y * ION_IPC;
    const int r_lo = chunk * rows_per_chunk, r_hi = min(nrows, r_lo + rows_per_chunk);
    double acc[ION_IPC][3];
#pragma unroll
    for (int q = 0; q < ION_IPC; ++q) acc[q][0] = acc[q][1] = acc[q][2] = 0.0;
    for (int row = r_lo; row < r_hi; ++row) {
        const int j0 = row / g.n1_loc, j1 = row - j0 * g.n1_loc + g.j1_off;
        if (threadIdx.x < ION_IPC) {
            const int ion = ion0 + threadIdx.x;
            s01[threadIdx.x] = ion < n_ions ? cmul(E0[(size_t)ion * g.n0 + j0], E1[(size_t)ion * g.n1 + j1]) : make_double2(0.0, 0.0);
        }
        __syncthreads();
        for (int j2 = threadIdx.x; j2 < nzh; j2 += ION_T) {
            const KPoint p = make_kpoint_at(g, j0, j1, j2);
            const double2 b = B[(size_t)row * nzh + j2];
#pragma unroll
            for (int q = 0; q < ION_IPC; ++q) {
                const int ion = ion0 + q;
                if (ion < n_ions) {
                    const double2 e = cmul(s01[q], E2[(size_t)ion * nzh + j2]);
                    const double im = fma(e.x, b.y, e.y * b.x);
                    acc[q][0] = fma(p.kx, im, acc[q][0]);
                    acc[q][1] = fma(p.ky, im, acc[q][1]);
                    acc[q][2] = fma(p.kz, im, acc[q][2]);
                }
            }
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < ION_IPC; ++q)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v = warp_sum(acc[q][a]);
            if (lane == 0) red[3 * q + a][warp] = v;
        }
    __syncthreads();
    if (threadIdx.x < 3 * ION_IPC) {
        const int q = threadIdx.x / 3, a = threadIdx.x - 3 * q;
        if (ion0 + q < n_ions) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < ION_T / 32; ++w) v += red[threadIdx.x][w];
            partial[((size_t)(ion0 + q) * nchunks + chunk) * 3 + a] = v;
        }
    }
}

__global__ void k_ion_force_finish(const double* __restrict__ partial, int n_ions, int nchunks, double* __restrict__ forces) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * n_ions) return;
    const int ion = e / 3, a = e - 3 * ion;
    double v = 0.0;
    for (int c = 0; c < nchunks; ++c) v += partial[((size_t)ion * nchunks + c) * 3 + a];
    forces[e] = -v;
}

// block size (multiple of 32, <= 256) that wastes the fewest lanes on the n2/2 + 1 columns of a row
int spectrum_threads(int nzh) {
    int best = 256, best_waste = 1 << 30;
    for (int t = 96; t <= 256; t += 32) {
        const int waste = (nzh + t - 1) / t * t - nzh;
        if (waste < best_waste || (waste == best_waste && t > best)) { best = t; best_waste = waste; }
    }
    return best;
}
int spectrum_grid(int nrows) {
    const int blocks = (nrows + ION_RB - 1) / ION_RB;
    return blocks < 148 * 4 ? blocks : 148 * 4;
}

struct IonScratch {
    double2 *E0, *E1, *E2;
    double* slopes;
};

int check_species(const pad_plan* p, const pad_species* sp, int n_species, const char* who) {
    if (!p || !sp || n_species < 1) { pad_set_error("%s: null argument / no species", who); return PAD_ERR_ARG; }
    for (int s = 0; s < n_species; ++s) {
        if (!sp[s].table_dev || sp[s].n_table < 3 || !(sp[s].k_max > 0.0) || sp[s].n_ions < 0 || (sp[s].n_ions > 0 && !sp[s].frac_dev)) {
            pad_set_error("%s: bad species %d (table %p, n_table %d, k_max %g, n_ions %d)", who, s, (const void*)sp[s].table_dev,
                          sp[s].n_table, sp[s].k_max, sp[s].n_ions);
            return PAD_ERR_ARG;
        }
    }
    return PAD_OK;
}

// phase tables and table slopes of one species in the plan's (grown on demand) ion scratch
int prepare_species(pad_plan* p, const pad_species& sp, cudaStream_t s, IonScratch& W, UniformTable& T) {
    const size_t per = (size_t)p->n0 + p->n1 + p->nzh;
    const size_t need = sizeof(double2) * per * (size_t)(sp.n_ions > 0 ? sp.n_ions : 1) + sizeof(double) * (size_t)sp.n_table;
    if (p->ion_scratch_bytes < need) {
        if (p->ion_scratch) { PAD_CUDA(cudaStreamSynchronize(s)); cudaFree(p->ion_scratch); p->bytes_allocated -= p->ion_scratch_bytes; }
        PAD_CUDA(cudaMalloc(&p->ion_scratch, need));
        p->ion_scratch_bytes = need;
        p->bytes_allocated += need;
    }
    W.E0 = reinterpret_cast<double2*>(p->ion_scratch);
    W.E1 = W.E0 + (size_t)sp.n_ions * p->n0;
    W.E2 = W.E1 + (size_t)sp.n_ions * p->n1;
    W.slopes = reinterpret_cast<double*>(W.E2 + (size_t)sp.n_ions * p->nzh);
    k_table_slopes<<<(sp.n_table + 255) / 256, 256, 0, s>>>(sp.table_dev, sp.table_dev + sp.n_table, W.slopes, sp.n_table);
    ++g_pad_launches;
    if (sp.n_ions > 0) {
        k_ion_phases<<<pad_grid_for(per * sp.n_ions), PAD_THREADS, 0, s>>>(sp.frac_dev, sp.n_ions, p->n0, p->n1, p->nzh, W.E0, W.E1, W.E2);
        ++g_pad_launches;
    }
    T.eta = sp.table_dev; T.w = sp.table_dev + sp.n_table; T.m = W.slopes; T.n = sp.n_table;
    T.eta_max = sp.k_max;
    T.inv_d = (double)(sp.n_table - 1) / sp.k_max;
    PAD_CUDA(cudaGetLastError());
    return PAD_OK;
}

}  // namespace

extern "C" int pad_ionic_potential(pad_plan* p, const pad_species* species, int n_species, double* v_ext_out, void* stream) {
    PAD_TRY(check_species(p, species, n_species, "pad_ionic_potential"));
    if (!v_ext_out) { pad_set_error("pad_ionic_potential: null output"); return PAD_ERR_ARG; }
    PAD_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    cufftDoubleComplex* G;
    PAD_TRY(pad_get_cbuf(p, 0, &G));
    const int nrows = p->n0 * p->geom.n1_loc;
    const int threads = spectrum_threads(p->nzh);
    const int grid = spectrum_grid(nrows);
    for (int sI = 0; sI < n_species; ++sI) {
        IonScratch W;
        UniformTable T;
        PAD_TRY(prepare_species(p, species[sI], s, W, T));
        k_ion_spectrum<<<grid, threads, 0, s>>>(p->geom, nrows, T, species[sI].z, species[sI].n_ions, W.E0, W.E1, W.E2, 1.0 / p->vol,
                                            sI > 0, 0, reinterpret_cast<double2*>(G));
        ++g_pad_launches;
        PAD_CUDA(cudaGetLastError());
    }
    // unnormalised c2r = irfftn(..., norm='forward') (ion_utils.py:118)
    PAD_TRY(pad_fft_inverse(p, G, v_ext_out, s));
    return PAD_OK;
}

extern "C" int pad_ion_forces(pad_plan* p, const pad_species* species, int n_species, const double* den, double* forces_out,
                              void* stream) {
    PAD_TRY(check_species(p, species, n_species, "pad_ion_forces"));
    if (!den || !forces_out) { pad_set_error("pad_ion_forces: null argument"); return PAD_ERR_ARG; }
    PAD_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    cufftDoubleComplex *R, *Bs;
    PAD_TRY(pad_get_cbuf(p, 0, &R));
    PAD_TRY(pad_get_cbuf(p, 1, &Bs));
    PAD_TRY(pad_fft_forward(p, den, R, s));      // out-of-place r2c: the input is left untouched
    const int nrows = p->n0 * p->geom.n1_loc;
    const KGeom geom = p->geom;
    const double scale = p->dV / p->vol;
    const int gridk = pad_grid_for(p->Nk);
    int ion_off = 0;
    for (int sI = 0; sI < n_species; ++sI) {
        const pad_species& sp = species[sI];
        if (sp.n_ions == 0) continue;
        IonScratch W;
        UniformTable T;
        PAD_TRY(prepare_species(p, sp, s, W, T));
        {
            const cufftDoubleComplex* Rc = R;
            cufftDoubleComplex* Bc = Bs;
            const double z = sp.z;
            auto f = [=] __device__(uint32_t idx, const KPoint& k) {
                const bool edge = k.j2 == 0 || (geom.e2 && k.j2 == geom.n2 / 2);
                const double w = scale * (edge ? 1.0 : 2.0) * recpot_value(T, z, k.kx, k.ky, k.kz);
                const cufftDoubleComplex r = Rc[idx];
                Bc[idx] = make_cuDoubleComplex(w * r.x, -w * r.y);
            };
            ks_kernel<decltype(f)><<<gridk, PAD_THREADS, 0, s>>>(geom, (uint32_t)p->Nk, f);
            ++g_pad_launches;
        }
        const int groups = (sp.n_ions + ION_IPC - 1) / ION_IPC;
        int nchunks = (148 * 6 + groups - 1) / groups;
        if (nchunks > nrows) nchunks = nrows;
        if (nchunks < 1) nchunks = 1;
        const int rows_per_chunk = (nrows + nchunks - 1) / nchunks;
        nchunks = (nrows + rows_per_chunk - 1) / rows_per_chunk;
        const size_t pbytes = sizeof(double) * 3 * (size_t)sp.n_ions * nchunks;
        if (p->ion_partial_bytes < pbytes) {
            if (p->ion_partial) { PAD_CUDA(cudaStreamSynchronize(s)); cudaFree(p->ion_partial); }
            PAD_CUDA(cudaMalloc(&p->ion_partial, pbytes));
            p->ion_partial_bytes = pbytes;
        }
        k_ion_force_partial<<<dim3(nchunks, groups), ION_T, 0, s>>>(geom, nrows, rows_per_chunk, sp.n_ions, W.E0, W.E1, W.E2,
                                                                    reinterpret_cast<const double2*>(Bs), p->ion_partial);
        k_ion_force_finish<<<(3 * sp.n_ions + 127) / 128, 128, 0, s>>>(p->ion_partial, sp.n_ions, nchunks, forces_out + 3 * (size_t)ion_off);
        g_pad_launches += 2;
        PAD_CUDA(cudaGetLastError());
        ion_off += sp.n_ions;
    }
    return PAD_OK;
}

// IonElectron part of System.__compute_stress (system.py:927-935): ions at fixed fractional coordinates, so S(k) does
// not depend on the cell and only v_s(|k|) / vol does:
//   sigma_ij = -delta_ij E / vol - (1/vol) sum_k w v_s'(|k|) k_i k_j / |k| Re(S_s(k) conj c_k),   E = sum_k w v_s Re(S_s conj c_k)
extern "C" int pad_ion_stress(pad_plan* p, const pad_species* species, int n_species, const double* den, double* stress_out,
                              int accumulate, void* stream) {
    PAD_TRY(check_species(p, species, n_species, "pad_ion_stress"));
    if (!den || !stress_out) { pad_set_error("pad_ion_stress: null argument"); return PAD_ERR_ARG; }
    PAD_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) PAD_CUDA(cudaMemsetAsync(stress_out, 0, sizeof(double) * 9, s));
    cufftDoubleComplex *R, *Sk;
    PAD_TRY(pad_get_cbuf(p, 0, &R));
    PAD_TRY(pad_get_cbuf(p, 1, &Sk));
    PAD_TRY(pad_fft_forward(p, den, R, s));
    const int nrows = p->n0 * p->geom.n1_loc;
    const int threads = spectrum_threads(p->nzh);
    const int grid = spectrum_grid(nrows);
    const KGeom geom = p->geom;
    const double inv_n = geom.inv_n;
    for (int sI = 0; sI < n_species; ++sI) {
        const pad_species& sp = species[sI];
        if (sp.n_ions == 0) continue;
        IonScratch W;
        UniformTable T;
        PAD_TRY(prepare_species(p, sp, s, W, T));
        k_ion_spectrum<<<grid, threads, 0, s>>>(geom, nrows, T, sp.z, sp.n_ions, W.E0, W.E1, W.E2, 1.0, 0, 1, reinterpret_cast<double2*>(Sk));
        ++g_pad_launches;
        const cufftDoubleComplex *Rc = R, *Sc = Sk;
        const double z = sp.z;
        auto f = [=] __device__(size_t i, double(&acc)[7]) {
            const KPoint k = make_kpoint(geom, (uint32_t)i);
            const cufftDoubleComplex r = Rc[i], sk = Sc[i];
            const bool edge = k.j2 == 0 || (geom.e2 && k.j2 == geom.n2 / 2);
            const double re = (edge ? 1.0 : 2.0) * (sk.x * r.x + sk.y * r.y) * inv_n;
            const double k2 = k.kx * k.kx + k.ky * k.ky + k.kz * k.kz;
            double val, slope;
            if (k2 == 0.0) {
                table_lookup_slope(T, 0.0, val, slope);
                acc[0] -= val * re;
                return;
            }
            const double ka = sqrt(k2);
            table_lookup_slope(T, ka, val, slope);
            val -= 4.0 * kPi * z / k2;
            slope += 8.0 * kPi * z / (k2 * ka);
            acc[0] -= val * re;
            const double t = -slope / ka * re;
            acc[1] += t * k.kx * k.kx; acc[2] += t * k.ky * k.ky; acc[3] += t * k.kz * k.kz;
            acc[4] += t * k.kx * k.ky; acc[5] += t * k.kx * k.kz; acc[6] += t * k.ky * k.kz;
        };
        ew_kernel<7, decltype(f)><<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(p->Nk, f, p->partials);
        ++g_pad_launches;
        PAD_TRY(pad_stress_accumulate(p, s, pad_grid_for(p->Nk), 1.0 / p->vol, 1.0 / p->vol, stress_out));
    }
    PAD_CUDA(cudaGetLastError());
    return PAD_OK;
}

// =================================================================================================
//  Particle-mesh Ewald structure factor (Essmann et al. 1995): replaces structure_factor_spline (ion_utils.py:218-286)
//  -- cardinal B-spline spreading of the ions onto the grid (cardinal_b_spline_values, ion_utils.py:140-204), one r2c,
//  Euler exponential-spline factors b_a(m) (exponential_spline_b, ion_utils.py:207-215):
//      S(k) = conj( b_0(j0) b_1(j1) b_2(j2) rfftn(Q)(k) ),   Q[(i - floor(u_a)) mod N_a ...] += prod_a M_n(u_a - floor(u_a) + i_a)
//  O(N_ion order^3 + N log N) instead of O(N_k N_ion).  The spreading uses fp64 atomicAdd: the order of the (few)
//  additions into a grid point varies from run to run, i.e. S is reproducible to rounding, not bit for bit.
// =================================================================================================
namespace {

constexpr int PME_MAX_ORDER = 32;

// M_n(x + i), i < n, by the Cox-de Boor recursion of cardinal_b_spline_values (x in [0, 1))
__host__ __device__ inline void bspline_weights(double x, int order, double* M /* [PME_MAX_ORDER] */) {
    for (int i = 0; i < order; ++i) M[i] = 0.0;
    M[0] = x;
    M[1] = 1.0 - x;
    for (int n = 3; n <= order; ++n) {
        double prev = 0.0;          // M_{n-1}[i-1]
        for (int i = 0; i < n; ++i) {
            const double cur = i < n - 1 ? M[i] : 0.0;          // M_{n-1}[i] (zero beyond its support)
            const double left = (x + (double)i) * cur;
            const double right = i >= 1 ? ((double)n - x - (double)i) * prev : 0.0;
            M[i] = (left + right) / (double)(n - 1);
            prev = cur;
        }
    }
}

// (slab plans: every rank sweeps ALL ions and keeps the stencil points that fall on its own x-planes [x_lo, x_lo + n0_loc))
__global__ void __launch_bounds__(128) k_pme_spread(const double* __restrict__ frac, int n_ions, int order, int N0, int N1, int N2,
                                                   int x_lo, int n0_loc, double* __restrict__ Q) {
    __shared__ double w[3][PME_MAX_ORDER];
    __shared__ int fl[3];
    for (int ion = blockIdx.x; ion < n_ions; ion += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 3) {
            const int a = threadIdx.x;
            const int N = a == 0 ? N0 : (a == 1 ? N1 : N2);
            double f = frac[3 * ion + a];
            f -= floor(f);
            f -= floor(f);
            const double u = f * (double)N;
            const double fu = floor(u);
            fl[a] = (int)fu;
            bspline_weights(u - fu, order, w[a]);
        }
        __syncthreads();
        const int total = order * order * order;
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int i = e / (order * order), j = (e / order) % order, k = e % order;
            int a0 = (i - fl[0]) % N0, a1 = (j - fl[1]) % N1, a2 = (k - fl[2]) % N2;
            if (a0 < 0) a0 += N0;
            if (a1 < 0) a1 += N1;
            if (a2 < 0) a2 += N2;
            a0 -= x_lo;
            if (a0 < 0 || a0 >= n0_loc) continue;
            atomicAdd(Q + ((size_t)a0 * N1 + a1) * N2 + a2, w[0][i] * w[1][j] * w[2][k]);
        }
    }
}

// host: b(m) = exp(2 pi i m (n - 1) / N) / sum_i M_n(i) exp(2 pi i m (i - 1) / N),  m < count
void pme_b_table(int count, int N, int order, std::vector<double2>& out) {
    double M[PME_MAX_ORDER];
    bspline_weights(0.0, order, M);
    out.resize(count);
    for (int m = 0; m < count; ++m) {
        double dr = 0.0, di = 0.0;
        for (int i = 0; i < order; ++i) {
            const double ph = 2.0 * kPi * (double)m * (double)(i - 1) / (double)N;
            dr += M[i] * cos(ph);
            di += M[i] * sin(ph);
        }
        const double ph = 2.0 * kPi * (double)m * (double)(order - 1) / (double)N;
        const double nr = cos(ph), ni = sin(ph), d2 = dr * dr + di * di;
        out[m] = make_double2((nr * dr + ni * di) / d2, (ni * dr - nr * di) / d2);
    }
}

// G(k) (+)= v_s(|k|) conj(b0 b1 b2 Qhat(k)) / vol, Hermitian part on the special points (as k_ion_spectrum);
// raw: plain S(k) into out, no potential factor, no symmetrisation
__global__ void __launch_bounds__(PAD_THREADS) k_pme_spectrum(KGeom g, uint32_t nk, UniformTable T, double z, const double2* __restrict__ Qh,
                                                            const double2* __restrict__ b0, const double2* __restrict__ b1,
                                                            const double2* __restrict__ b2, double inv_vol, int accumulate, int raw,
                                                            double2* __restrict__ out) {
    const uint32_t stride = gridDim.x * PAD_THREADS;
    for (uint32_t idx = blockIdx.x * PAD_THREADS + threadIdx.x; idx < nk; idx += stride) {
        const KPoint p = make_kpoint(g, idx);
        // Qhat is the transform of a REAL field: on the self-conjugate planes Qhat(pbar) = conj Qhat(p), so the partner value
        // of a special point never has to be fetched (on slab plans it may live on another rank)
        const double2 qh = Qh[idx];
        auto S_at = [&](int j0, int j1, int j2, bool partner) {
            const double2 q = partner ? make_double2(qh.x, -qh.y) : qh;
            const double2 b = cmul(cmul(b0[j0], b1[j1]), b2[j2]);
            const double2 t = cmul(b, q);
            return make_double2(t.x, -t.y);
        };
        const double2 S = S_at(p.j0, p.j1, p.j2, false);
        double2 G;
        if (raw) {
            G = S;
        } else {
            const double f = recpot_value(T, z, p.kx, p.ky, p.kz);
            G = make_double2(f * S.x, f * S.y);
            if (p.special) {
                const double2 Sb = S_at((g.n0 - p.j0) % g.n0, (g.n1 - p.j1) % g.n1, p.j2, true);
                const double fb = recpot_value(T, z, p.px, p.py, p.pz);
                G.x = 0.5 * (G.x + fb * Sb.x);
                G.y = 0.5 * (G.y - fb * Sb.y);
            }
            G.x *= inv_vol; G.y *= inv_vol;
        }
        double2 o = accumulate ? out[idx] : make_double2(0.0, 0.0);
        o.x += G.x; o.y += G.y;
        out[idx] = o;
    }
}

int pme_b_tables(pad_plan* p, int order, cudaStream_t s, double2** btab);

// spread + r2c + b tables of one species; Qhat in cbuf 1, b tables in *btab (device, caller frees)
int pme_prepare(pad_plan* p, const double* frac_dev, int n_ions, int order, cudaStream_t s, cufftDoubleComplex** Qh, double2** btab) {
    if (order < 2 || order > PME_MAX_ORDER || (order & 1)) { pad_set_error("Requires even order 2 <= n <= %d", PME_MAX_ORDER); return PAD_ERR_ARG; }
    double* Q;
    PAD_TRY(pad_get_rbuf(p, 0, &Q));
    PAD_TRY(pad_get_cbuf(p, 1, Qh));
    PAD_CUDA(cudaMemsetAsync(Q, 0, sizeof(double) * p->N, s));
    if (n_ions > 0) {
        k_pme_spread<<<n_ions < 148 * 8 ? n_ions : 148 * 8, 128, 0, s>>>(frac_dev, n_ions, order, p->n0, p->n1, p->n2,
                                                                       p->dist ? p->rank * p->n0_loc : 0, p->dist ? p->n0_loc : p->n0, Q);
        ++g_pad_launches;
    }
    PAD_TRY(pad_fft_forward(p, Q, *Qh, s));
    return pme_b_tables(p, order, s, btab);
}

// weights and their derivatives with respect to x: d/dx M_n(x + i) = M_{n-1}(x + i) - M_{n-1}(x + i - 1)
__device__ inline void bspline_weights_derivs(double x, int order, double* M, double* dM) {
    for (int i = 0; i < order; ++i) M[i] = 0.0;
    if (order == 2) {
        M[0] = x; M[1] = 1.0 - x;
        dM[0] = 1.0; dM[1] = -1.0;
        return;
    }
    bspline_weights(x, order - 1, M);                    // M_{n-1}(x + i), i < n - 1 (M[n-1] = 0)
    double prev = 0.0;
    for (int i = 0; i < order; ++i) {
        const double cur = i < order - 1 ? M[i] : 0.0;
        dM[i] = cur - prev;
        const double left = (x + (double)i) * cur;
        const double right = i >= 1 ? ((double)order - x - (double)i) * prev : 0.0;
        M[i] = (left + right) / (double)(order - 1);
        prev = cur;
    }
}

// A(k) = v_s(|k|) conj(b0 b1 b2 rho_hat(k)), Hermitian part on the special points: the unnormalised c2r of A is the field
// Phi(r) = sum_k w_k Re[A(k) e^{ikr}] the spread charges are contracted with,  E = (dV / vol) sum_r Q(r) Phi(r)
__global__ void __launch_bounds__(PAD_THREADS) k_pme_force_spectrum(KGeom g, uint32_t nk, UniformTable T, double z, const double2* __restrict__ R,
                                                                  const double2* __restrict__ b0, const double2* __restrict__ b1,
                                                                  const double2* __restrict__ b2, double2* __restrict__ out) {
    const uint32_t stride = gridDim.x * PAD_THREADS;
    for (uint32_t idx = blockIdx.x * PAD_THREADS + threadIdx.x; idx < nk; idx += stride) {
        const KPoint p = make_kpoint(g, idx);
        const double2 rh = R[idx];
        auto A_at = [&](int j0, int j1, int j2, double kx, double ky, double kz, bool partner) {
            const double2 r = partner ? make_double2(rh.x, -rh.y) : rh;      // rho_hat(pbar) = conj rho_hat(p) on these planes
            const double2 t = cmul(cmul(cmul(b0[j0], b1[j1]), b2[j2]), r);
            const double f = recpot_value(T, z, kx, ky, kz);
            return make_double2(f * t.x, -f * t.y);
        };
        double2 A = A_at(p.j0, p.j1, p.j2, p.kx, p.ky, p.kz, false);
        if (p.special) {
            const double2 Ab = A_at((g.n0 - p.j0) % g.n0, (g.n1 - p.j1) % g.n1, p.j2, p.px, p.py, p.pz, true);
            A.x = 0.5 * (A.x + Ab.x);
            A.y = 0.5 * (A.y - Ab.y);
        }
        out[idx] = A;
    }
}

// one CTA per ion: F_c = -(dV / vol) sum_a N_a (B^-1)_{ca} sum_stencil dW_a/du_a prod_{b != a} W_b Phi(grid point)
__global__ void __launch_bounds__(128) k_pme_force_gather(const double* __restrict__ frac, int n_ions, int order, int N0, int N1, int N2,
                                                         int x_lo, int n0_loc, const double* __restrict__ Phi, double scale, double bi00, double bi01, double bi02,
                                                         double bi10, double bi11, double bi12, double bi20, double bi21, double bi22,
                                                         double* __restrict__ forces /* n_ions x 3 */) {
    __shared__ double w[3][PME_MAX_ORDER], dw[3][PME_MAX_ORDER];
    __shared__ int fl[3];
    __shared__ double red[3][4];
    for (int ion = blockIdx.x; ion < n_ions; ion += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 3) {
            const int a = threadIdx.x;
            const int N = a == 0 ? N0 : (a == 1 ? N1 : N2);
            double f = frac[3 * ion + a];
            f -= floor(f);
            f -= floor(f);
            const double u = f * (double)N;
            const double fu = floor(u);
            fl[a] = (int)fu;
            bspline_weights_derivs(u - fu, order, w[a], dw[a]);
        }
        __syncthreads();
        double g0 = 0.0, g1 = 0.0, g2 = 0.0;
        const int total = order * order * order;
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int i = e / (order * order), j = (e / order) % order, k = e % order;
            int a0 = (i - fl[0]) % N0, a1 = (j - fl[1]) % N1, a2 = (k - fl[2]) % N2;
            if (a0 < 0) a0 += N0;
            if (a1 < 0) a1 += N1;
            if (a2 < 0) a2 += N2;
            a0 -= x_lo;
            if (a0 < 0 || a0 >= n0_loc) continue;          // slab plans: this rank's planes; the caller adds the ranks' partial forces
            const double ph = Phi[((size_t)a0 * N1 + a1) * N2 + a2];
            g0 += dw[0][i] * w[1][j] * w[2][k] * ph;
            g1 += w[0][i] * dw[1][j] * w[2][k] * ph;
            g2 += w[0][i] * w[1][j] * dw[2][k] * ph;
        }
        g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { red[0][warp] = g0; red[1][warp] = g1; red[2][warp] = g2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            // dE/du_a, u_a = N_a frac_a, frac = R B^-1  =>  dE/dR_c = sum_a N_a (B^-1)_{ca} dE/du_a
            const double d0 = (red[0][0] + red[0][1] + red[0][2] + red[0][3]) * (double)N0;
            const double d1 = (red[1][0] + red[1][1] + red[1][2] + red[1][3]) * (double)N1;
            const double d2 = (red[2][0] + red[2][1] + red[2][2] + red[2][3]) * (double)N2;
            forces[3 * ion + 0] = -scale * (bi00 * d0 + bi01 * d1 + bi02 * d2);
            forces[3 * ion + 1] = -scale * (bi10 * d0 + bi11 * d1 + bi12 * d2);
            forces[3 * ion + 2] = -scale * (bi20 * d0 + bi21 * d1 + bi22 * d2);
        }
    }
}

int pme_b_tables(pad_plan* p, int order, cudaStream_t s, double2** btab) {
    std::vector<double2> h0, h1, h2, all;
    pme_b_table(p->n0, p->n0, order, h0);
    pme_b_table(p->n1, p->n1, order, h1);
    pme_b_table(p->nzh, p->n2, order, h2);
    all.insert(all.end(), h0.begin(), h0.end());
    all.insert(all.end(), h1.begin(), h1.end());
    all.insert(all.end(), h2.begin(), h2.end());
    PAD_CUDA(cudaMalloc(btab, sizeof(double2) * all.size()));
    PAD_CUDA(cudaMemcpyAsync(*btab, all.data(), sizeof(double2) * all.size(), cudaMemcpyHostToDevice, s));
    PAD_CUDA(cudaStreamSynchronize(s));       // `all` is pageable host memory that dies with this frame
    return PAD_OK;
}

}  // namespace

// S(k) of the particle-mesh scheme over the half spectrum (n0, n1, n2/2 + 1), complex: what structure_factor_spline returns
extern "C" int pad_pme_structure_factor(pad_plan* p, const double* frac_dev, int n_ions, int order, double* S_out_cplx, void* stream) {
    if (!p || !frac_dev || !S_out_cplx || n_ions < 1) { pad_set_error("pad_pme_structure_factor: bad argument"); return PAD_ERR_ARG; }
    PAD_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    cufftDoubleComplex* Qh;
    double2* bt = nullptr;
    PAD_TRY(pme_prepare(p, frac_dev, n_ions, order, s, &Qh, &bt));
    UniformTable T{};
    k_pme_spectrum<<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(p->geom, (uint32_t)p->Nk, T, 0.0, reinterpret_cast<const double2*>(Qh), bt,
                                                            bt + p->n0, bt + p->n0 + p->n1, 0.0, 0, 1,
                                                            reinterpret_cast<double2*>(S_out_cplx));
    ++g_pad_launches;
    PAD_CUDA(cudaGetLastError());
    PAD_CUDA(cudaStreamSynchronize(s));
    cudaFree(bt);
    return PAD_OK;
}

// v_ext with particle-mesh structure factors: System.__potential_from_ions with pme_order (system.py:183-205)
extern "C" int pad_ionic_potential_pme(pad_plan* p, const pad_species* species, int n_species, int order, double* v_ext_out,
                                       void* stream) {
    PAD_TRY(check_species(p, species, n_species, "pad_ionic_potential_pme"));
    if (!v_ext_out) { pad_set_error("pad_ionic_potential_pme: null output"); return PAD_ERR_ARG; }
    PAD_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    cufftDoubleComplex* G;
    PAD_TRY(pad_get_cbuf(p, 0, &G));
    for (int sI = 0; sI < n_species; ++sI) {
        IonScratch W;
        UniformTable T;
        pad_species tab_only = species[sI];
        tab_only.n_ions = 0;                                    // only the table slopes are needed, no phase tables
        PAD_TRY(prepare_species(p, tab_only, s, W, T));
        cufftDoubleComplex* Qh;
        double2* bt = nullptr;
        PAD_TRY(pme_prepare(p, species[sI].frac_dev, species[sI].n_ions, order, s, &Qh, &bt));
        k_pme_spectrum<<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(p->geom, (uint32_t)p->Nk, T, species[sI].z,
                                                                reinterpret_cast<const double2*>(Qh), bt, bt + p->n0, bt + p->n0 + p->n1,
                                                                1.0 / p->vol, sI > 0, 0, reinterpret_cast<double2*>(G));
        ++g_pad_launches;
        PAD_CUDA(cudaGetLastError());
        PAD_CUDA(cudaStreamSynchronize(s));
        cudaFree(bt);
    }
    PAD_TRY(pad_fft_inverse(p, G, v_ext_out, s));
    return PAD_OK;
}


// IonElectron forces with the particle-mesh structure factor: what the reference's autograd through structure_factor_spline gives
// (system.py:913-925 with pme_order set).  E = (dV / vol) sum_r Q(r) Phi(r) with Q the B-spline charges (fixed stencil, weights
// smooth in R) and Phi the unnormalised c2r of v_s conj(b rho_hat); one r2c of the density, and per species one k-space pass, one
// c2r and one gather over the order^3 stencil of every ion: O(N log N + N_ion order^3) instead of O(N_k N_ion).
extern "C" int pad_ion_forces_pme(pad_plan* p, const pad_species* species, int n_species, int order, const double* den,
                                  double* forces_out, void* stream) {
    PAD_TRY(check_species(p, species, n_species, "pad_ion_forces_pme"));
    if (!den || !forces_out) { pad_set_error("pad_ion_forces_pme: null argument"); return PAD_ERR_ARG; }
    if (order < 2 || order > PME_MAX_ORDER || (order & 1)) { pad_set_error("Requires even order 2 <= n <= %d", PME_MAX_ORDER); return PAD_ERR_ARG; }
    PAD_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    cufftDoubleComplex *R, *A;
    double* Phi;
    PAD_TRY(pad_get_cbuf(p, 0, &R));
    PAD_TRY(pad_get_cbuf(p, 1, &A));
    PAD_TRY(pad_get_rbuf(p, 0, &Phi));
    PAD_TRY(pad_fft_forward(p, den, R, s));
    double2* bt = nullptr;
    PAD_TRY(pme_b_tables(p, order, s, &bt));
    double inv[9], det;
    {   // B^-1 of the lattice (rows = lattice vectors)
        const double* m = p->box;
        const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
        const double Aa = e * i - f * h, Bb = -(d * i - f * g), Cc = d * h - e * g;
        det = a * Aa + b * Bb + c * Cc;
        const double id = 1.0 / det;
        inv[0] = Aa * id; inv[1] = -(b * i - c * h) * id; inv[2] = (b * f - c * e) * id;
        inv[3] = Bb * id; inv[4] = (a * i - c * g) * id;  inv[5] = -(a * f - c * d) * id;
        inv[6] = Cc * id; inv[7] = -(a * h - b * g) * id; inv[8] = (a * e - b * d) * id;
    }
    int ion_off = 0;
    for (int sI = 0; sI < n_species; ++sI) {
        const pad_species& sp = species[sI];
        if (sp.n_ions == 0) continue;
        IonScratch W;
        UniformTable T;
        pad_species tab_only = sp;
        tab_only.n_ions = 0;
        PAD_TRY(prepare_species(p, tab_only, s, W, T));
        k_pme_force_spectrum<<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(p->geom, (uint32_t)p->Nk, T, sp.z, reinterpret_cast<const double2*>(R), bt,
                                                                      bt + p->n0, bt + p->n0 + p->n1, reinterpret_cast<double2*>(A));
        PAD_CUDA(cudaGetLastError());
        PAD_TRY(pad_fft_inverse(p, A, Phi, s));
        k_pme_force_gather<<<sp.n_ions < 148 * 8 ? sp.n_ions : 148 * 8, 128, 0, s>>>(sp.frac_dev, sp.n_ions, order, p->n0, p->n1, p->n2,
                                                                                 p->dist ? p->rank * p->n0_loc : 0, p->dist ? p->n0_loc : p->n0, Phi,
                                                                                 p->dV / p->vol, inv[0], inv[1], inv[2], inv[3], inv[4],
                                                                                 inv[5], inv[6], inv[7], inv[8], forces_out + 3 * (size_t)ion_off);
        g_pad_launches += 2;
        PAD_CUDA(cudaGetLastError());
        ion_off += sp.n_ions;
    }
    PAD_CUDA(cudaStreamSynchronize(s));
    cudaFree(bt);
    return PAD_OK;
}

// IonElectron stress with the particle-mesh structure factor (system.py:927-935 with pme_order): the ions keep their fractional
// coordinates, so the spread charges and S(k) do not depend on the cell -- pad_ion_stress with S(k) from the mesh.
extern "C" int pad_ion_stress_pme(pad_plan* p, const pad_species* species, int n_species, int order, const double* den,
                                  double* stress_out, int accumulate, void* stream) {
    PAD_TRY(check_species(p, species, n_species, "pad_ion_stress_pme"));
    if (!den || !stress_out) { pad_set_error("pad_ion_stress_pme: null argument"); return PAD_ERR_ARG; }
    PAD_CUDA(cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) PAD_CUDA(cudaMemsetAsync(stress_out, 0, sizeof(double) * 9, s));
    cufftDoubleComplex *R, *Sk;
    PAD_TRY(pad_get_cbuf(p, 0, &R));
    PAD_TRY(pad_get_cbuf(p, 2, &Sk));
    PAD_TRY(pad_fft_forward(p, den, R, s));
    const KGeom geom = p->geom;
    const double inv_n = geom.inv_n;
    for (int sI = 0; sI < n_species; ++sI) {
        const pad_species& sp = species[sI];
        if (sp.n_ions == 0) continue;
        IonScratch W;
        UniformTable T;
        pad_species tab_only = sp;
        tab_only.n_ions = 0;
        PAD_TRY(prepare_species(p, tab_only, s, W, T));
        cufftDoubleComplex* Qh;
        double2* bt = nullptr;
        PAD_TRY(pme_prepare(p, sp.frac_dev, sp.n_ions, order, s, &Qh, &bt));      // (uses real buffer 0 and complex buffer 1)
        UniformTable T0{};
        k_pme_spectrum<<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(geom, (uint32_t)p->Nk, T0, 0.0, reinterpret_cast<const double2*>(Qh), bt,
                                                                bt + p->n0, bt + p->n0 + p->n1, 0.0, 0, 1, reinterpret_cast<double2*>(Sk));
        ++g_pad_launches;
        const cufftDoubleComplex *Rc = R, *Sc = Sk;
        const double z = sp.z;
        auto f = [=] __device__(size_t i, double(&acc)[7]) {
            const KPoint k = make_kpoint(geom, (uint32_t)i);
            const cufftDoubleComplex r = Rc[i], sk = Sc[i];
            const bool edge = k.j2 == 0 || (geom.e2 && k.j2 == geom.n2 / 2);
            const double re = (edge ? 1.0 : 2.0) * (sk.x * r.x + sk.y * r.y) * inv_n;
            const double k2 = k.kx * k.kx + k.ky * k.ky + k.kz * k.kz;
            double val, slope;
            if (k2 == 0.0) {
                table_lookup_slope(T, 0.0, val, slope);
                acc[0] -= val * re;
                return;
            }
            const double ka = sqrt(k2);
            table_lookup_slope(T, ka, val, slope);
            val -= 4.0 * kPi * z / k2;
            slope += 8.0 * kPi * z / (k2 * ka);
            acc[0] -= val * re;
            const double t = -slope / ka * re;
            acc[1] += t * k.kx * k.kx; acc[2] += t * k.ky * k.ky; acc[3] += t * k.kz * k.kz;
            acc[4] += t * k.kx * k.ky; acc[5] += t * k.kx * k.kz; acc[6] += t * k.ky * k.kz;
        };
        ew_kernel<7, decltype(f)><<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(p->Nk, f, p->partials);
        ++g_pad_launches;
        PAD_TRY(pad_stress_accumulate(p, s, pad_grid_for(p->Nk), 1.0 / p->vol, 1.0 / p->vol, stress_out));
        PAD_CUDA(cudaStreamSynchronize(s));
        cudaFree(bt);
    }
    PAD_CUDA(cudaGetLastError());
    return PAD_OK;
}
