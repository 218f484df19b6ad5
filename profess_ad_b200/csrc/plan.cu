// Plan management, scratch memory, cuFFT wrappers and the deterministic second-stage reduction.
// Replaces wavevecs() (functional_tools.py:135-162): the k-vectors are never materialised, the
// reciprocal lattice (9 doubles) travels to every reciprocal-space kernel by value.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_pad_launches = 0;
unsigned long long g_pad_fft_execs = 0;

extern "C" unsigned long long pad_launch_count(void) { return g_pad_launches; }
extern "C" unsigned long long pad_fft_exec_count(void) { return g_pad_fft_execs; }
int g_pad_fast_fft = 1;      // hand-written fused FFT pipeline where the grid allows it (0: plain cuFFT 3-D + separate elementwise kernels)
extern "C" int pad_set_fast_fft(int on) { const int old = g_pad_fast_fft; g_pad_fast_fft = on ? 1 : 0; return old; }
int g_pad_own_xy = 1;        // hand-written strided (x, y) passes with the fused multiply (n0, n1 in 64/128/256)
int g_pad_pipe = 0;          // (z, y) passes of a plane as items of one persistent kernel, handed over through the L2 (zy_pipe.cuh)
int g_pad_pipe_lpi = 0, g_pad_pipe_tpi = 0;
int g_pad_fuse_terms = 1;
int g_pad_fold_table = 1;
int g_pad_zinv_stream = 0;     // measured at 256^3: the streamed form (12 warps/SM, 168 registers) is 7-13 % slower than the batch form
int g_pad_local_tail = 1;      // measured at 256^3: locals in the mid pass 534 us + 4-field final pass 437 us; plain mid 264 + final 199 + tail
int g_pad_ywide = 0;
int g_pad_pbe_fast = 1;
int g_pad_xone = 0;       // measured at 256^3: 107 us against 77 us for the persistent prefetching kernel
int g_pad_graphs = 1;          // whole evaluations with unchanged arguments are captured once and replayed as ONE graph launch
unsigned long long g_pad_option_epoch = 0;
int g_pad_fuse_mid = 0;        // measured: mid + forward z in one 8-warp kernel 508 us, as two kernels 348 + 135 us
extern "C" int pad_set_option(const char* name, int value) {
    int* slot = nullptr;
    if (!name) { pad_set_error("pad_set_option: null name"); return -1; }
    if (!strcmp(name, "fast_fft")) slot = &g_pad_fast_fft;
    else if (!strcmp(name, "own_xy")) slot = &g_pad_own_xy;
    else if (!strcmp(name, "pipe")) slot = &g_pad_pipe;
    else if (!strcmp(name, "fuse_terms")) slot = &g_pad_fuse_terms;
    else if (!strcmp(name, "fold_table")) slot = &g_pad_fold_table;
    else if (!strcmp(name, "zinv_stream")) slot = &g_pad_zinv_stream;
    else if (!strcmp(name, "fuse_mid")) slot = &g_pad_fuse_mid;
    else if (!strcmp(name, "pipe_lpi")) slot = &g_pad_pipe_lpi;
    else if (!strcmp(name, "pipe_tpi")) slot = &g_pad_pipe_tpi;
    else if (!strcmp(name, "graphs")) slot = &g_pad_graphs;
    else if (!strcmp(name, "ywide")) slot = &g_pad_ywide;
    else if (!strcmp(name, "xone")) slot = &g_pad_xone;
    else if (!strcmp(name, "pbe_fast")) slot = &g_pad_pbe_fast;
    else if (!strcmp(name, "local_tail")) slot = &g_pad_local_tail;
    if (!slot) { pad_set_error("pad_set_option: unknown option %s", name); return -1; }
    const int old = *slot;
    *slot = value;
    ++g_pad_option_epoch;
    return old;
}

// ---- live per-stage timing ------------------------------------------------------------------------
int g_pad_profile = 0;
namespace {
constexpr int kMaxStages = 256;
struct StageProf {
    cudaEvent_t ev[kMaxStages + 1];
    const char* name[kMaxStages + 1];
    int n = 0;                 // events recorded in the evaluation in flight
    bool created = false;
    double sum_ms[kMaxStages + 1];
    const char* sum_name[kMaxStages + 1];
    int n_stages = 0, evals = 0;
} g_prof;

void prof_flush() {
    StageProf& P = g_prof;
    if (P.n < 2) { P.n = 0; return; }
    cudaEventSynchronize(P.ev[P.n - 1]);
    for (int i = 1; i < P.n; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, P.ev[i - 1], P.ev[i]);
        if (P.evals == 0) { P.sum_ms[i - 1] = 0.0; P.sum_name[i - 1] = P.name[i]; }
        P.sum_ms[i - 1] += ms;
    }
    if (P.evals == 0) P.n_stages = P.n - 1;
    ++P.evals;
    P.n = 0;
}
}  // namespace

void pad_stage_begin(cudaStream_t s) {
    if (!g_pad_profile) return;
    StageProf& P = g_prof;
    if (!P.created) {
        for (int i = 0; i <= kMaxStages; ++i) cudaEventCreate(&P.ev[i]);
        P.created = true;
    }
    prof_flush();
    cudaEventRecord(P.ev[0], s);
    P.name[0] = "begin";
    P.n = 1;
}

void pad_stage_mark(const char* name, cudaStream_t s) {
    if (!g_pad_profile) return;
    StageProf& P = g_prof;
    if (P.n < 1 || P.n > kMaxStages) return;
    cudaEventRecord(P.ev[P.n], s);
    P.name[P.n] = name;
    ++P.n;
}

extern "C" int pad_profile_begin(void) {
    g_prof.n = 0; g_prof.n_stages = 0; g_prof.evals = 0;
    g_pad_profile = 1;
    return PAD_OK;
}

extern "C" int pad_profile_end(char* names_out, double* ms_out, int cap, int* n_out, int* evals_out) {
    prof_flush();
    g_pad_profile = 0;
    const int n = g_prof.n_stages < cap ? g_prof.n_stages : cap;
    for (int i = 0; i < n; ++i) {
        if (ms_out) ms_out[i] = g_prof.evals ? g_prof.sum_ms[i] / g_prof.evals : 0.0;
        if (names_out) { strncpy(names_out + 48 * i, g_prof.sum_name[i], 47); names_out[48 * i + 47] = 0; }
    }
    if (n_out) *n_out = n;
    if (evals_out) *evals_out = g_prof.evals;
    return PAD_OK;
}

void pad_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* pad_last_error(void) { return g_err; }
extern "C" const char* pad_version(void) { return "professad_b200 0.1 (sm_100a, fp64)"; }

static int invert3(const double* m, double* inv, double* det_out) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C;
    *det_out = det;
    if (det == 0.0 || !isfinite(det)) return 1;
    const double id = 1.0 / det;
    inv[0] = A * id;               inv[1] = -(b * i - c * h) * id; inv[2] = (b * f - c * e) * id;
    inv[3] = B * id;               inv[4] = (a * i - c * g) * id;  inv[5] = -(a * f - c * d) * id;
    inv[6] = C * id;               inv[7] = -(a * h - b * g) * id; inv[8] = (a * e - b * d) * id;
    return 0;
}

static int set_box(pad_plan* p, const double* box) {
    // b = 2 pi inv(box^T)  (functional_tools.py:149); rows of b are the reciprocal vectors
    double bt[9], inv[9], det;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) bt[r * 3 + c] = box[c * 3 + r];
    if (invert3(bt, inv, &det)) {
        pad_set_error("Lattice vector matrix is not invertible.");
        return PAD_ERR_ARG;
    }
    memcpy(p->box, box, sizeof(double) * 9);
    for (int k = 0; k < 9; ++k) p->recip[k] = 2.0 * kPi * inv[k];
    p->vol = fabs(det);
    const double n_global = (double)p->n0 * (double)p->n1 * (double)p->n2;      // p->N is the local count on slab plans
    p->dV = p->vol / n_global;
    KGeom& g = p->geom;
    g.n0 = p->n0; g.n1 = p->n1; g.n2 = p->n2; g.nzh = p->nzh;
    g.e0 = (p->n0 % 2 == 0); g.e1 = (p->n1 % 2 == 0); g.e2 = (p->n2 % 2 == 0);
    memcpy(g.b, p->recip, sizeof(double) * 9);
    g.inv_n = 1.0 / n_global;
    g.nzp_pad = p->n2 / 2 + 8;
    g.n1_loc = p->dist ? p->n1_loc : p->n1;
    g.j1_off = p->dist ? p->rank * p->n1_loc : 0;
    p->box_generation++;
    return PAD_OK;
}

static int plan_create_common(pad_plan** out, const double* box, const int* shape, int device, int rank, int world,
                              void* send_buf, void* recv_buf, double* comm_scratch, pad_comm_fn fn, void* user) {
    if (!out || !box || !shape) {
        pad_set_error("pad_plan_create: null argument");
        return PAD_ERR_ARG;
    }
    if (shape[0] < 1 || shape[1] < 1 || shape[2] < 1) {
        pad_set_error("pad_plan_create: bad shape (%d,%d,%d)", shape[0], shape[1], shape[2]);
        return PAD_ERR_ARG;
    }
    const bool dist = world > 0;
    if (dist) {
        if (rank < 0 || rank >= world || !send_buf || !recv_buf || !comm_scratch || !fn) {
            pad_set_error("pad_plan_create_slab: bad rank/world (%d/%d) or null buffer / callback", rank, world);
            return PAD_ERR_ARG;
        }
        if (shape[0] % world || shape[1] % world) {
            pad_set_error("pad_plan_create_slab: n0 = %d and n1 = %d must be multiples of the world size %d",
                          shape[0], shape[1], world);
            return PAD_ERR_ARG;
        }
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        pad_set_error("professad_b200 needs a CUDA device (sm_100a); none is visible (%s). There is no CPU fallback.",
                      cudaGetErrorString(ce));
        return PAD_ERR_CUDA;
    }
    PAD_CUDA(cudaSetDevice(device));
    pad_plan* p = new pad_plan();
    memset(p, 0, sizeof(pad_plan));
    p->n0 = shape[0]; p->n1 = shape[1]; p->n2 = shape[2]; p->nzh = shape[2] / 2 + 1;
    p->dist = dist;
    if (dist) {
        p->rank = rank; p->world = world;
        p->n0_loc = p->n0 / world; p->n1_loc = p->n1 / world;
        p->send_buf = send_buf; p->recv_buf = recv_buf; p->comm_scratch = comm_scratch;
        p->comm_fn = fn; p->comm_user = user;
        p->N = (size_t)p->n0_loc * p->n1 * p->n2;
        p->Nk = (size_t)p->n0 * p->n1_loc * p->nzh;
    } else {
        p->N = (size_t)p->n0 * p->n1 * p->n2;
        p->Nk = (size_t)p->n0 * p->n1 * p->nzh;
    }
    p->device = device;
    p->nzp = p->n2 / 2 + 8;      // padded row length of the fused pipeline (multiple of 8 complex)
    if (p->Nk >= 0xffffffffull) {
        pad_set_error("grid too large for one device plan (%zu half-spectrum points)", p->Nk);
        delete p;
        return PAD_ERR_ARG;
    }
    int rc = set_box(p, box);
    if (rc != PAD_OK) { delete p; return rc; }
    PAD_CUDA(cudaMalloc(&p->partials, sizeof(double) * PAD_MAX_RED * PAD_MAX_BLOCKS));
    PAD_CUDA(cudaMalloc(&p->scal, sizeof(double) * PAD_N_SCAL));
    PAD_CUDA(cudaMemset(p->scal, 0, sizeof(double) * PAD_N_SCAL));
    p->bytes_allocated = sizeof(double) * (PAD_MAX_RED * PAD_MAX_BLOCKS + PAD_N_SCAL);
    *out = p;
    return PAD_OK;
}

extern "C" int pad_plan_create(pad_plan** out, const double* box, const int* shape, int device) {
    return plan_create_common(out, box, shape, device, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int pad_plan_create_slab(pad_plan** out, const double* box, const int* global_shape, int device, int rank,
                                    int world, void* send_buf, void* recv_buf, double* comm_scratch, pad_comm_fn fn,
                                    void* user) {
    if (world < 1) { pad_set_error("pad_plan_create_slab: world must be >= 1"); return PAD_ERR_ARG; }
    return plan_create_common(out, box, global_shape, device, rank, world, send_buf, recv_buf, comm_scratch, fn, user);
}

extern "C" int pad_plan_set_box(pad_plan* p, const double* box) {
    if (!p || !box) { pad_set_error("pad_plan_set_box: null argument"); return PAD_ERR_ARG; }
    return set_box(p, box);
}

extern "C" size_t pad_plan_workspace_bytes(const pad_plan* p) { return p ? p->bytes_allocated : 0; }

extern "C" int pad_plan_destroy(pad_plan* p) {
    if (!p) return PAD_OK;
    cudaSetDevice(p->device);
    if (p->fft_ready && !p->dist) { cufftDestroy(p->d2z); cufftDestroy(p->z2d); }
    if (p->fft_ready && p->dist) { cufftDestroy(p->d2z_yz); cufftDestroy(p->z2d_yz); cufftDestroy(p->z2z_x); }
    if (p->comm_ready) {
        cudaStreamDestroy(p->comm_stream);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(p->ev_ready[i]); cudaEventDestroy(p->ev_a2a[i]); cudaEventDestroy(p->ev_free[i]); }
    }
    if (p->fft_work) cudaFree(p->fft_work);
    for (int i = 0; i < PAD_N_RBUF; ++i) if (p->rbuf[i]) cudaFree(p->rbuf[i]);
    for (int i = 0; i < PAD_N_CBUF; ++i) if (p->cbuf[i]) cudaFree(p->cbuf[i]);
    if (p->wgc_kern) cudaFree(p->wgc_kern);
    if (p->wgc_kern4) cudaFree(p->wgc_kern4);
    if (p->wt_kern) cudaFree(p->wt_kern);
    if (p->hc_scratch) cudaFree(p->hc_scratch);
    if (p->hc_slopes) cudaFree(p->hc_slopes);
    if (p->hc_conv) cudaFree(p->hc_conv);
    if (p->ion_scratch) cudaFree(p->ion_scratch);
    if (p->ion_partial) cudaFree(p->ion_partial);
    if (p->xy_ready) cufftDestroy(p->xy);
    if (p->xy_work) cudaFree(p->xy_work);
    for (int i = 0; i < 4; ++i) if (p->zbuf[i]) cudaFree(p->zbuf[i]);
    for (int i = 0; i < 16; ++i)
        if (p->graphs[i].exec) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(p->graphs[i].exec));
    if (p->graph_stream) cudaStreamDestroy(p->graph_stream);
    if (p->pipe_ctl) cudaFree(p->pipe_ctl);
    if (p->pipe_part) cudaFree(p->pipe_part);
    cudaFree(p->partials);
    cudaFree(p->scal);
    delete p;
    return PAD_OK;
}

int pad_get_rbuf(pad_plan* p, int i, double** out) {
    if (i < 0 || i >= PAD_N_RBUF) { pad_set_error("rbuf index %d", i); return PAD_ERR_ARG; }
    if (!p->rbuf[i]) {
        PAD_CUDA(cudaMalloc(&p->rbuf[i], sizeof(double) * p->N));
        p->bytes_allocated += sizeof(double) * p->N;
    }
    *out = p->rbuf[i];
    return PAD_OK;
}

int pad_get_cbuf(pad_plan* p, int i, cufftDoubleComplex** out) {
    if (i < 0 || i >= PAD_N_CBUF) { pad_set_error("cbuf index %d", i); return PAD_ERR_ARG; }
    if (!p->cbuf[i]) {
        PAD_CUDA(cudaMalloc(&p->cbuf[i], sizeof(cufftDoubleComplex) * p->Nk));
        p->bytes_allocated += sizeof(cufftDoubleComplex) * p->Nk;
    }
    *out = p->cbuf[i];
    return PAD_OK;
}

static int ensure_fft(pad_plan* p, cudaStream_t s) {
    if (!p->fft_ready) {
        size_t w1 = 0, w2 = 0;
        PAD_CUFFT(cufftCreate(&p->d2z));
        PAD_CUFFT(cufftCreate(&p->z2d));
        PAD_CUFFT(cufftSetAutoAllocation(p->d2z, 0));
        PAD_CUFFT(cufftSetAutoAllocation(p->z2d, 0));
        PAD_CUFFT(cufftMakePlan3d(p->d2z, p->n0, p->n1, p->n2, CUFFT_D2Z, &w1));
        PAD_CUFFT(cufftMakePlan3d(p->z2d, p->n0, p->n1, p->n2, CUFFT_Z2D, &w2));
        size_t w = w1 > w2 ? w1 : w2;
        if (w > 0) {
            PAD_CUDA(cudaMalloc(&p->fft_work, w));
            p->bytes_allocated += w;
        }
        p->fft_work_bytes = w;
        PAD_CUFFT(cufftSetWorkArea(p->d2z, p->fft_work));
        PAD_CUFFT(cufftSetWorkArea(p->z2d, p->fft_work));
        p->fft_ready = true;
        p->fft_stream = (cudaStream_t)(-1);
    }
    if (p->fft_stream != s) {
        PAD_CUFFT(cufftSetStream(p->d2z, s));
        PAD_CUFFT(cufftSetStream(p->z2d, s));
        p->fft_stream = s;
    }
    return PAD_OK;
}

// ---- slab plans: 3-D transform = batched 2-D (y, z) cuFFT + all-to-all + strided 1-D (x) cuFFT ------------
static int ensure_fft_slab(pad_plan* p, cudaStream_t s) {
    if (!p->fft_ready) {
        size_t w[3] = {0, 0, 0};
        PAD_CUFFT(cufftCreate(&p->d2z_yz));
        PAD_CUFFT(cufftCreate(&p->z2d_yz));
        PAD_CUFFT(cufftCreate(&p->z2z_x));
        PAD_CUFFT(cufftSetAutoAllocation(p->d2z_yz, 0));
        PAD_CUFFT(cufftSetAutoAllocation(p->z2d_yz, 0));
        PAD_CUFFT(cufftSetAutoAllocation(p->z2z_x, 0));
        int nyz[2] = {p->n1, p->n2};
        PAD_CUFFT(cufftMakePlanMany(p->d2z_yz, 2, nyz, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, p->n0_loc, &w[0]));
        PAD_CUFFT(cufftMakePlanMany(p->z2d_yz, 2, nyz, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, p->n0_loc, &w[1]));
        int nx[1] = {p->n0};
        const int lines = p->n1_loc * p->nzh;             // (n0, n1_loc, nzh): x has stride `lines`, lines are contiguous
        PAD_CUFFT(cufftMakePlanMany(p->z2z_x, 1, nx, nx, lines, 1, nx, lines, 1, CUFFT_Z2Z, lines, &w[2]));
        size_t wmax = w[0] > w[1] ? w[0] : w[1];
        if (w[2] > wmax) wmax = w[2];
        if (wmax > 0) {
            PAD_CUDA(cudaMalloc(&p->fft_work, wmax));
            p->bytes_allocated += wmax;
        }
        p->fft_work_bytes = wmax;
        PAD_CUFFT(cufftSetWorkArea(p->d2z_yz, p->fft_work));
        PAD_CUFFT(cufftSetWorkArea(p->z2d_yz, p->fft_work));
        PAD_CUFFT(cufftSetWorkArea(p->z2z_x, p->fft_work));
        p->fft_ready = true;
        p->fft_stream = (cudaStream_t)(-1);
    }
    if (p->fft_stream != s) {
        PAD_CUFFT(cufftSetStream(p->d2z_yz, s));
        PAD_CUFFT(cufftSetStream(p->z2d_yz, s));
        PAD_CUFFT(cufftSetStream(p->z2z_x, s));
        p->fft_stream = s;
    }
    return PAD_OK;
}

// (n0_loc, n1, nzh) <-> (world, n0_loc, n1_loc, nzh): block r holds the y range of rank r
__global__ void __launch_bounds__(PAD_THREADS) slab_permute_kernel(const double2* __restrict__ src, double2* __restrict__ dst,
                                                                   int n0_loc, int n1, int n1_loc, int nzh, int to_blocks) {
    const size_t total = (size_t)n0_loc * n1 * nzh;
    for (size_t e = (size_t)blockIdx.x * PAD_THREADS + threadIdx.x; e < total; e += (size_t)gridDim.x * PAD_THREADS) {
        const int k = (int)(e % nzh);
        const size_t row = e / nzh;
        const int j = (int)(row % n1);
        const int i = (int)(row / n1);
        const int r = j / n1_loc, jl = j - r * n1_loc;
        const size_t blocked = (((size_t)r * n0_loc + i) * n1_loc + jl) * nzh + k;
        if (to_blocks) dst[blocked] = src[e];
        else dst[e] = src[blocked];
    }
}

// peer form of the two permutes: the forward pack stores into the owners' receive buffers, the inverse copies its x-range blocks
struct PeerRecv {
    double2* p[8];
};
__global__ void __launch_bounds__(PAD_THREADS) slab_permute_push_kernel(const double2* __restrict__ src, const __grid_constant__ PeerRecv dst,
                                                                        int n0_loc, int n1, int n1_loc, int nzh, int rank) {
    const size_t total = (size_t)n0_loc * n1 * nzh;
    for (size_t e = (size_t)blockIdx.x * PAD_THREADS + threadIdx.x; e < total; e += (size_t)gridDim.x * PAD_THREADS) {
        const int k = (int)(e % nzh);
        const size_t row = e / nzh;
        const int j = (int)(row % n1);
        const int i = (int)(row / n1);
        const int r = j / n1_loc, jl = j - r * n1_loc;
        dst.p[r][(((size_t)rank * n0_loc + i) * n1_loc + jl) * nzh + k] = src[e];     // block `rank` of rank r's (world, n0_loc, n1_loc, nzh)
    }
}
__global__ void __launch_bounds__(PAD_THREADS) slab_block_push_kernel(const double2* __restrict__ src, const __grid_constant__ PeerRecv dst,
                                                                      size_t block, int world, int rank) {
    const size_t total = block * world;
    for (size_t e = (size_t)blockIdx.x * PAD_THREADS + threadIdx.x; e < total; e += (size_t)gridDim.x * PAD_THREADS) {
        const int r = (int)(e / block);
        dst.p[r][(size_t)rank * block + (e - (size_t)r * block)] = src[e];
    }
}
static PeerRecv peer_recv_of(const pad_plan* p, int b) {
    PeerRecv d;
    for (int r = 0; r < 8; ++r) d.p[r] = r < p->world ? reinterpret_cast<double2*>(p->peer_recv[b][r]) : nullptr;
    return d;
}

extern "C" int pad_plan_set_slab_peer_recv(pad_plan* p, void* const* recv, void* const* recv2, int world) {
    if (!p || !p->dist || !recv || !recv2 || world != p->world || world > 8 || !p->recv_buf2) {
        pad_set_error("pad_plan_set_slab_peer_recv: needs a slab plan with overlap buffers and the receive buffers of all ranks (world <= 8)");
        return PAD_ERR_ARG;
    }
    for (int r = 0; r < world; ++r) { p->peer_recv[0][r] = recv[r]; p->peer_recv[1][r] = recv2[r]; }
    if (recv[p->rank] != p->recv_buf || recv2[p->rank] != p->recv_buf2) {
        pad_set_error("pad_plan_set_slab_peer_recv: this rank's entries must be its own receive buffers");
        return PAD_ERR_ARG;
    }
    p->recv_push = true;
    p->recv_parity = 0;
    return PAD_OK;
}

int pad_slab_comm(pad_plan* p, int op, long long count, cudaStream_t s) {
    const int rc = p->comm_fn(p->comm_user, op, count, (void*)s);
    if (rc != 0) {
        pad_set_error("slab plan: the communication callback failed (op %d, rc %d)", op, rc);
        return PAD_ERR_NCCL;
    }
    return PAD_OK;
}

static int fft_forward_slab(pad_plan* p, const double* in, cufftDoubleComplex* out, cudaStream_t s) {
    PAD_TRY(ensure_fft_slab(p, s));
    PAD_CUFFT(cufftExecD2Z(p->d2z_yz, const_cast<double*>(in), out));                       // (n0_loc, n1, nzh)
    if (p->recv_push) {
        if (p->recv_after_pipeline) {      // the pipelined batches use the pair in their own order: re-synchronise once
            PAD_TRY(pad_slab_comm(p, PAD_COMM_BARRIER, 0, s));
            p->recv_after_pipeline = false;
        }
        // Receive buffers alternate from transform to transform.  A rank enters the barrier of transform t after its x FFT of
        // transform t - 1 (stream order), so once the barrier of t has let this rank through, every rank has consumed the buffer
        // transform t + 1 is about to be pushed into: one barrier per transform is enough.
        const int b = p->recv_parity;
        p->recv_parity ^= 1;
        slab_permute_push_kernel<<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(reinterpret_cast<const double2*>(out), peer_recv_of(p, b),
                                                                            p->n0_loc, p->n1, p->n1_loc, p->nzh, p->rank);
        PAD_CUDA(cudaGetLastError());
        PAD_TRY(pad_slab_comm(p, PAD_COMM_BARRIER, 0, s));
        PAD_CUFFT(cufftExecZ2Z(p->z2z_x, reinterpret_cast<cufftDoubleComplex*>(b ? p->recv_buf2 : p->recv_buf), out, CUFFT_FORWARD));
        g_pad_fft_execs += 2;
        ++g_pad_launches;
        return PAD_OK;
    }
    slab_permute_kernel<<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(reinterpret_cast<const double2*>(out),
        reinterpret_cast<double2*>(p->send_buf), p->n0_loc, p->n1, p->n1_loc, p->nzh, 1);
    PAD_CUDA(cudaGetLastError());
    PAD_TRY(pad_slab_comm(p, PAD_COMM_ALL_TO_ALL, (long long)p->n0_loc * p->n1_loc * p->nzh, s)); // recv = (n0, n1_loc, nzh)
    PAD_CUFFT(cufftExecZ2Z(p->z2z_x, reinterpret_cast<cufftDoubleComplex*>(p->recv_buf), out, CUFFT_FORWARD));
    g_pad_fft_execs += 2;
    ++g_pad_launches;
    return PAD_OK;
}

static int fft_inverse_slab(pad_plan* p, cufftDoubleComplex* in, double* out, cudaStream_t s) {
    PAD_TRY(ensure_fft_slab(p, s));
    if (p->recv_push) {
        if (p->recv_after_pipeline) {
            PAD_TRY(pad_slab_comm(p, PAD_COMM_BARRIER, 0, s));
            p->recv_after_pipeline = false;
        }
        const int b = p->recv_parity;
        p->recv_parity ^= 1;
        // x inverse into the send buffer (blocked by x range), blocks copied into the owners' receive buffers, barrier, unpack
        PAD_CUFFT(cufftExecZ2Z(p->z2z_x, in, reinterpret_cast<cufftDoubleComplex*>(p->send_buf), CUFFT_INVERSE));
        const size_t block = (size_t)p->n0_loc * p->n1_loc * p->nzh;
        slab_block_push_kernel<<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(reinterpret_cast<const double2*>(p->send_buf), peer_recv_of(p, b),
                                                                          block, p->world, p->rank);
        PAD_CUDA(cudaGetLastError());
        PAD_TRY(pad_slab_comm(p, PAD_COMM_BARRIER, 0, s));
        slab_permute_kernel<<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(reinterpret_cast<const double2*>(b ? p->recv_buf2 : p->recv_buf),
            reinterpret_cast<double2*>(in), p->n0_loc, p->n1, p->n1_loc, p->nzh, 0);
        PAD_CUDA(cudaGetLastError());
        PAD_CUFFT(cufftExecZ2D(p->z2d_yz, in, out));
        g_pad_fft_execs += 2;
        g_pad_launches += 2;
        return PAD_OK;
    }
    // x inverse into the send buffer: (n0, n1_loc, nzh) is already blocked by x range
    PAD_CUFFT(cufftExecZ2Z(p->z2z_x, in, reinterpret_cast<cufftDoubleComplex*>(p->send_buf), CUFFT_INVERSE));
    PAD_TRY(pad_slab_comm(p, PAD_COMM_ALL_TO_ALL, (long long)p->n0_loc * p->n1_loc * p->nzh, s)); // recv = (world, n0_loc, n1_loc, nzh)
    slab_permute_kernel<<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(reinterpret_cast<const double2*>(p->recv_buf),
        reinterpret_cast<double2*>(in), p->n0_loc, p->n1, p->n1_loc, p->nzh, 0);
    PAD_CUDA(cudaGetLastError());
    PAD_CUFFT(cufftExecZ2D(p->z2d_yz, in, out));
    g_pad_fft_execs += 2;
    ++g_pad_launches;
    return PAD_OK;
}

// ---- pipelined batches ---------------------------------------------------------------------------------------
extern "C" int pad_plan_set_overlap_buffers(pad_plan* p, void* send_buf2, void* recv_buf2) {
    if (!p || !p->dist) { pad_set_error("pad_plan_set_overlap_buffers: needs a slab plan"); return PAD_ERR_ARG; }
    p->send_buf2 = send_buf2;
    p->recv_buf2 = recv_buf2;
    return PAD_OK;
}

extern "C" size_t pad_slab_fast_elements(const pad_plan* p) {
    return (p && p->dist) ? (size_t)p->n0_loc * p->n1 * p->nzp : 0;
}
extern "C" int pad_plan_set_slab_fast_buffers(pad_plan* p, void* const* six) {
    if (!p || !p->dist || !six) { pad_set_error("pad_plan_set_slab_fast_buffers: needs a slab plan and six buffers"); return PAD_ERR_ARG; }
    for (int i = 0; i < 6; ++i) {
        if (!six[i]) { pad_set_error("pad_plan_set_slab_fast_buffers: buffer %d is null", i); return PAD_ERR_ARG; }
        p->slab_fast[i] = six[i];
    }
    return PAD_OK;
}

extern "C" int pad_plan_set_slab_peer_buffers(pad_plan* p, void* const* base, int world) {
    if (!p || !p->dist || !base || world != p->world || world > 8) {
        pad_set_error("pad_plan_set_slab_peer_buffers: needs a slab plan, base pointers of all %d ranks and world <= 8", p ? p->world : 0);
        return PAD_ERR_ARG;
    }
    const size_t each = (size_t)p->n0_loc * p->n1 * p->nzp * sizeof(double) * 2;
    for (int r = 0; r < world; ++r) {
        if (!base[r]) { pad_set_error("pad_plan_set_slab_peer_buffers: rank %d has no buffer", r); return PAD_ERR_ARG; }
        for (int f = 0; f < 4; ++f) {
            p->peer_B[f][r] = static_cast<char*>(base[r]) + (size_t)f * each;
            p->peer_T[f][r] = static_cast<char*>(base[r]) + (size_t)(4 + f) * each;
        }
    }
    for (int f = 0; f < 4; ++f) p->slab_fast[f] = p->peer_B[f][p->rank];
    p->slab_fast[4] = p->slab_fast[5] = nullptr;
    p->slab_push = true;
    return PAD_OK;
}

int pad_ensure_comm_stream(pad_plan* p);
static int ensure_comm_stream(pad_plan* p) { return pad_ensure_comm_stream(p); }
int pad_ensure_comm_stream(pad_plan* p) {
    if (p->comm_ready) return PAD_OK;
    PAD_CUDA(cudaStreamCreateWithFlags(&p->comm_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        PAD_CUDA(cudaEventCreateWithFlags(&p->ev_ready[i], cudaEventDisableTiming));
        PAD_CUDA(cudaEventCreateWithFlags(&p->ev_a2a[i], cudaEventDisableTiming));
        PAD_CUDA(cudaEventCreateWithFlags(&p->ev_free[i], cudaEventDisableTiming));
    }
    p->comm_ready = true;
    return PAD_OK;
}

static bool can_overlap(const pad_plan* p, int n) { return p->dist && p->world > 1 && p->send_buf2 && p->recv_buf2 && n >= 2; }

// exchange of buffer pair b on the communication stream: after the producer on `s` (ev_ready[b]) and, from the third
// field on, after the consumer of the pair's previous contents (ev_free[b])
static int exchange_async(pad_plan* p, int b, bool wait_free, cudaStream_t s) {
    PAD_CUDA(cudaEventRecord(p->ev_ready[b], s));
    PAD_CUDA(cudaStreamWaitEvent(p->comm_stream, p->ev_ready[b], 0));
    if (wait_free) PAD_CUDA(cudaStreamWaitEvent(p->comm_stream, p->ev_free[b], 0));
    if (p->recv_push) {
        // peer form: barrier (every rank's consumer of receive buffer b is done -- each rank enters it after its own ev_free[b]),
        // blocks copied into the owners' receive buffers over NVLink, barrier (every rank's blocks have arrived)
        PAD_TRY(pad_slab_comm(p, PAD_COMM_BARRIER_2, 0, p->comm_stream));
        const size_t block = (size_t)p->n0_loc * p->n1_loc * p->nzh;
        slab_block_push_kernel<<<pad_grid_for(p->Nk), PAD_THREADS, 0, p->comm_stream>>>(
            reinterpret_cast<const double2*>(b ? p->send_buf2 : p->send_buf), peer_recv_of(p, b), block, p->world, p->rank);
        PAD_CUDA(cudaGetLastError());
        ++g_pad_launches;
        PAD_TRY(pad_slab_comm(p, PAD_COMM_BARRIER_2, 0, p->comm_stream));
        p->recv_after_pipeline = true;
    } else {
        PAD_TRY(pad_slab_comm(p, b ? PAD_COMM_ALL_TO_ALL_2 : PAD_COMM_ALL_TO_ALL, (long long)p->n0_loc * p->n1_loc * p->nzh, p->comm_stream));
    }
    PAD_CUDA(cudaEventRecord(p->ev_a2a[b], p->comm_stream));
    return PAD_OK;
}

int pad_fft_forward_many(pad_plan* p, const double* const* in, cufftDoubleComplex* const* out, int n, cudaStream_t s) {
    if (!can_overlap(p, n)) {
        for (int f = 0; f < n; ++f) PAD_TRY(pad_fft_forward(p, in[f], out[f], s));
        return PAD_OK;
    }
    PAD_TRY(ensure_fft_slab(p, s));
    PAD_TRY(ensure_comm_stream(p));
    void* sb[2] = {p->send_buf, p->send_buf2};
    void* rb[2] = {p->recv_buf, p->recv_buf2};
    auto local_then_exchange = [&](int f) -> int {
        const int b = f & 1;
        if (f >= 2) PAD_CUDA(cudaStreamWaitEvent(s, p->ev_a2a[b], 0));          // exchange f - 2 has left the send buffer
        PAD_CUFFT(cufftExecD2Z(p->d2z_yz, const_cast<double*>(in[f]), out[f]));
        slab_permute_kernel<<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(reinterpret_cast<const double2*>(out[f]),
            reinterpret_cast<double2*>(sb[b]), p->n0_loc, p->n1, p->n1_loc, p->nzh, 1);
        PAD_CUDA(cudaGetLastError());
        return exchange_async(p, b, f >= 2, s);
    };
    auto finish = [&](int f) -> int {
        const int b = f & 1;
        PAD_CUDA(cudaStreamWaitEvent(s, p->ev_a2a[b], 0));
        PAD_CUFFT(cufftExecZ2Z(p->z2z_x, reinterpret_cast<cufftDoubleComplex*>(rb[b]), out[f], CUFFT_FORWARD));
        PAD_CUDA(cudaEventRecord(p->ev_free[b], s));
        return PAD_OK;
    };
    PAD_TRY(local_then_exchange(0));
    for (int f = 1; f < n; ++f) {
        PAD_TRY(local_then_exchange(f));
        PAD_TRY(finish(f - 1));
    }
    PAD_TRY(finish(n - 1));
    g_pad_fft_execs += 2 * n;
    g_pad_launches += n;
    return PAD_OK;
}

int pad_fft_inverse_many(pad_plan* p, cufftDoubleComplex* const* in, double* const* out, int n, cudaStream_t s) {
    if (!can_overlap(p, n)) {
        for (int f = 0; f < n; ++f) PAD_TRY(pad_fft_inverse(p, in[f], out[f], s));
        return PAD_OK;
    }
    PAD_TRY(ensure_fft_slab(p, s));
    PAD_TRY(ensure_comm_stream(p));
    void* sb[2] = {p->send_buf, p->send_buf2};
    void* rb[2] = {p->recv_buf, p->recv_buf2};
    auto local_then_exchange = [&](int f) -> int {
        const int b = f & 1;
        if (f >= 2) PAD_CUDA(cudaStreamWaitEvent(s, p->ev_a2a[b], 0));
        PAD_CUFFT(cufftExecZ2Z(p->z2z_x, in[f], reinterpret_cast<cufftDoubleComplex*>(sb[b]), CUFFT_INVERSE));
        return exchange_async(p, b, f >= 2, s);
    };
    auto finish = [&](int f) -> int {
        const int b = f & 1;
        PAD_CUDA(cudaStreamWaitEvent(s, p->ev_a2a[b], 0));
        slab_permute_kernel<<<pad_grid_for(p->Nk), PAD_THREADS, 0, s>>>(reinterpret_cast<const double2*>(rb[b]),
            reinterpret_cast<double2*>(in[f]), p->n0_loc, p->n1, p->n1_loc, p->nzh, 0);
        PAD_CUDA(cudaGetLastError());
        PAD_CUDA(cudaEventRecord(p->ev_free[b], s));
        PAD_CUFFT(cufftExecZ2D(p->z2d_yz, in[f], out[f]));
        return PAD_OK;
    };
    PAD_TRY(local_then_exchange(0));
    for (int f = 1; f < n; ++f) {
        PAD_TRY(local_then_exchange(f));
        PAD_TRY(finish(f - 1));
    }
    PAD_TRY(finish(n - 1));
    g_pad_fft_execs += 2 * n;
    g_pad_launches += n;
    return PAD_OK;
}

int pad_fft_forward(pad_plan* p, const double* in, cufftDoubleComplex* out, cudaStream_t s) {
    if (p->dist) return fft_forward_slab(p, in, out, s);
    PAD_TRY(ensure_fft(p, s));
    PAD_CUFFT(cufftExecD2Z(p->d2z, const_cast<double*>(in), out));
    ++g_pad_fft_execs;
    return PAD_OK;
}

int pad_fft_inverse(pad_plan* p, cufftDoubleComplex* in, double* out, cudaStream_t s) {
    if (p->dist) return fft_inverse_slab(p, in, out, s);
    PAD_TRY(ensure_fft(p, s));
    PAD_CUFFT(cufftExecZ2D(p->z2d, in, out));
    ++g_pad_fft_execs;
    return PAD_OK;
}

// ---------------------------------------------------------------------------------------------
//  second-stage reduction: one CTA sums the per-block partials in a fixed order
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PAD_THREADS) finalize_kernel(const double* __restrict__ partials, FinalizeArgs a) {
    __shared__ double sm[PAD_THREADS / 32];
    __shared__ double total[PAD_MAX_RED];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = 0; t < a.nterms; ++t) {
        double v = 0.0;
        for (int b = threadIdx.x; b < a.nblocks; b += PAD_THREADS) v += partials[(size_t)t * PAD_MAX_BLOCKS + b];
        v = warp_sum(v);
        if (lane == 0) sm[warp] = v;
        __syncthreads();
        if (warp == 0) {
            double w = lane < PAD_THREADS / 32 ? sm[lane] : 0.0;
            w = warp_sum(w);
            if (lane == 0) total[t] = w;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double e = 0.0;
        for (int t = 0; t < a.nterms; ++t) {
            if (a.sums_out) a.sums_out[t] = total[t];
            e += a.coef[t] * total[t];
        }
        if (a.E_out) a.E_out[0] = (a.accumulate ? a.E_out[0] : 0.0) + e;
    }
}

// slab plans: the raw sums go through the caller's all-reduce before they are combined
__global__ void finalize_apply_kernel(const double* __restrict__ totals, FinalizeArgs a) {
    double e = 0.0;
    for (int t = 0; t < a.nterms; ++t) {
        if (a.sums_out) a.sums_out[t] = totals[t];
        e += a.coef[t] * totals[t];
    }
    if (a.E_out) a.E_out[0] = (a.accumulate ? a.E_out[0] : 0.0) + e;
}

__global__ void k_bits_to_scratch(const unsigned long long* bits, double* scratch, int n, int to_scratch) {
    const int i = threadIdx.x;
    if (i >= n) return;
    if (to_scratch) scratch[i] = __longlong_as_double((long long)bits[i]);
    else const_cast<unsigned long long*>(bits)[i] = (unsigned long long)__double_as_longlong(scratch[i]);
}

int pad_allreduce_max_bits(pad_plan* p, unsigned long long* bits, int n, cudaStream_t s) {
    if (!p->dist) return PAD_OK;
    if (n > PAD_COMM_SCRATCH) { pad_set_error("pad_allreduce_max_bits: %d values", n); return PAD_ERR_ARG; }
    k_bits_to_scratch<<<1, 64, 0, s>>>(bits, p->comm_scratch, n, 1);
    PAD_TRY(pad_slab_comm(p, PAD_COMM_ALL_REDUCE_MAX, n, s));
    k_bits_to_scratch<<<1, 64, 0, s>>>(bits, p->comm_scratch, n, 0);
    g_pad_launches += 2;
    PAD_CUDA(cudaGetLastError());
    return PAD_OK;
}

void pad_launch_finalize(pad_plan* p, const FinalizeArgs& a, cudaStream_t s) {
    if (!p->dist) {
        finalize_kernel<<<1, PAD_THREADS, 0, s>>>(p->partials, a);
        ++g_pad_launches;
        return;
    }
    FinalizeArgs raw = a;
    raw.sums_out = p->comm_scratch;
    raw.E_out = nullptr;
    finalize_kernel<<<1, PAD_THREADS, 0, s>>>(p->partials, raw);
    if (pad_slab_comm(p, PAD_COMM_ALL_REDUCE, a.nterms, s) != PAD_OK) return;      // error text is set; the caller's next CUDA check reports
    finalize_apply_kernel<<<1, 1, 0, s>>>(p->comm_scratch, a);
    g_pad_launches += 2;
}
