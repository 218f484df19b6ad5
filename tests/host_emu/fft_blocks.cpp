// Host emulation of the register / warp / tile FFT building blocks of csrc/fft_core.cuh and csrc/fft_strided.cuh:
// every lane (thread of a tile) is a host thread, __syncwarp / __syncthreads are std::barrier.  Checks fft_reg<8|16|32>,
// line_fft8 for (M, TPL) = (64, 8), (128, 16), (256, 32) and the two-stage tile FFT for L = 64 ... 512 against an
// O(n^2) DFT, together with the slot <-> index maps the kernels rely on.  No GPU needed.
#include "cuda_runtime.h"
thread_local std::barrier<>* g_warp_barrier = nullptr;
thread_local std::barrier<>* g_block_barrier = nullptr;
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#include "fft_core.cuh"
// the tile FFT part of fft_strided.cuh (SPass, spass_out_index, spass_slot_of_input, tile_fft), cut out of the real header by
// tests/test_host_emulation.py
#define PAD_Y256_CTAS 4      // (occupancy knob of the real header; irrelevant on host threads)
#include "tile_part.h"
#include <complex>
#include <cstdio>
#include <thread>
#include <vector>
#include <random>
typedef std::complex<double> C;
static void dft(const std::vector<C>& in, std::vector<C>& out, int dir) {
    int n = in.size(); out.assign(n, 0);
    for (int k = 0; k < n; ++k) { C s = 0; for (int j = 0; j < n; ++j) s += in[j] * std::polar(1.0, dir * 2 * M_PI * ((j * k) % n) / n); out[k] = s; }
}
template <int L, int DIR, bool WIDE = false> double test_tile() {
    using P = SPass<L, WIDE>;
    std::mt19937 g(L); std::uniform_real_distribution<double> u(-1, 1);
    std::vector<std::vector<C>> in(8, std::vector<C>(L)), ref(8);
    for (int c = 0; c < 8; ++c) { for (auto& x : in[c]) x = C(u(g), u(g)); dft(in[c], ref[c], DIR); }
    static cd S[L * 8]; static cd tw[L];
    for (int e = 0; e < L; ++e) tw[e] = cd{std::cos(-2 * M_PI * e / L), std::sin(-2 * M_PI * e / L)};
    const int NT = P::TILE_THREADS;
    std::barrier<> bar(NT);
    std::vector<double> err(NT, 0.0);
    std::vector<std::thread> th;
    for (int tid = 0; tid < NT; ++tid) th.emplace_back([&, tid] {
        g_block_barrier = &bar;
        const int t = tid / 8, c = tid % 8;
        cd v[P::EPT];
        for (int j = 0; j < P::EPT; ++j) v[j] = cd{in[c][t + P::TPL * j].real(), in[c][t + P::TPL * j].imag()};
        tile_fft<L, DIR, WIDE>(v, S, t, c, tw);
        for (int s = 0; s < P::EPT; ++s) err[tid] = std::max(err[tid], std::abs(C(v[s].x, v[s].y) - ref[c][spass_out_index<L, WIDE>(t, s)]));
        // inverse fed from the forward's slots: u[j] = v[slot_of_input(j)] must be Z[t + TPL j]
        for (int j = 0; j < P::EPT; ++j) {
            const int s = spass_slot_of_input<L, WIDE>(j);
            if (spass_out_index<L, WIDE>(t, s) != t + P::TPL * j) err[tid] = 1e9;
        }
    });
    for (auto& x : th) x.join();
    double e = 0; for (double x : err) e = std::max(e, x); return e;
}
template <int R, int DIR> double test_reg() {
    std::mt19937 g(R); std::uniform_real_distribution<double> u(-1, 1);
    std::vector<C> in(R), ref; cd v[R];
    for (int i = 0; i < R; ++i) { in[i] = C(u(g), u(g)); v[i] = cd{in[i].real(), in[i].imag()}; }
    dft(in, ref, DIR); fft_reg<R, DIR>(v);
    double e = 0; for (int r = 0; r < R; ++r) { int k = fft_nat<R>(r); e = std::max(e, std::abs(C(v[r].x, v[r].y) - ref[k])); if (fft_slot<R>(k) != r) e = 1e9; }
    return e;
}
template <int M, int TPL, int DIR> double test_line() {
    std::mt19937 g(M); std::uniform_real_distribution<double> u(-1, 1);
    std::vector<C> in(M), ref; for (auto& c : in) c = C(u(g), u(g));
    dft(in, ref, DIR);
    static cd S[8 * (TPL + 1)]; static cd tw1[M];
    for (int e = 0; e < M; ++e) tw1[e] = cd{std::cos(-2 * M_PI * e / M), std::sin(-2 * M_PI * e / M)};
    std::barrier<> bar(TPL);
    std::vector<double> err(TPL, 0.0);
    std::vector<std::thread> th;
    for (int t = 0; t < TPL; ++t) th.emplace_back([&, t] {
        g_warp_barrier = &bar;
        cd v[8];
        for (int j = 0; j < 8; ++j) v[j] = cd{in[t + TPL * j].real(), in[t + TPL * j].imag()};
        line_fft8<M, TPL, DIR>(v, S, t, tw1);
        for (int r = 0; r < 8; ++r) err[t] = std::max(err[t], std::abs(C(v[r].x, v[r].y) - ref[t + TPL * fft_nat<8>(r)]));
    });
    for (auto& x : th) x.join();
    double e = 0; for (double x : err) e = std::max(e, x); return e;
}
int main() {
    double worst = 0;
    auto chk = [&](const char* name, double e) { printf("%-14s %.3g\n", name, e); if (!(e < 2e-13)) worst = 1; };
    chk("reg8", std::max(test_reg<8, -1>(), test_reg<8, 1>()));
    chk("reg16", std::max(test_reg<16, -1>(), test_reg<16, 1>()));
    chk("reg32", std::max(test_reg<32, -1>(), test_reg<32, 1>()));
    chk("line 64/8", std::max(test_line<64, 8, -1>(), test_line<64, 8, 1>()));
    chk("line 128/16", std::max(test_line<128, 16, -1>(), test_line<128, 16, 1>()));
    chk("line 256/32", std::max(test_line<256, 32, -1>(), test_line<256, 32, 1>()));
    chk("tile 64", std::max(test_tile<64, -1>(), test_tile<64, 1>()));
    chk("tile 128", std::max(test_tile<128, -1>(), test_tile<128, 1>()));
    chk("tile 256", std::max(test_tile<256, -1>(), test_tile<256, 1>()));
    chk("tile 512", std::max(test_tile<512, -1>(), test_tile<512, 1>()));
    chk("tile 512 wide", std::max(test_tile<512, -1, true>(), test_tile<512, 1, true>()));
    chk("tile 256 wide", std::max(test_tile<256, -1, true>(), test_tile<256, 1, true>()));
    return worst != 0;
}
