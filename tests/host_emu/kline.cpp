// KLine (|k|^2 along an x line of the half spectrum from three per-line constants, csrc/common.cuh) against
// make_kpoint_at + sym_even for every point of small even / odd / skewed grids, including the special points.
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <random>
#define __device__
#define __forceinline__ inline
#include "kgeom_part.h"
int main() {
    std::mt19937 rng(7); std::uniform_real_distribution<double> u(-0.3, 0.3);
    double worst = 0;
    const int shapes[5][3] = {{8, 6, 10}, {7, 9, 5}, {6, 8, 7}, {4, 4, 4}, {16, 12, 8}};
    for (auto& sh : shapes) {
        KGeom g;
        g.n0 = sh[0]; g.n1 = sh[1]; g.n2 = sh[2]; g.nzh = sh[2] / 2 + 1;
        g.e0 = sh[0] % 2 == 0; g.e1 = sh[1] % 2 == 0; g.e2 = sh[2] % 2 == 0;
        for (int i = 0; i < 9; ++i) g.b[i] = (i % 4 == 0 ? 0.8 : 0.0) + u(rng);
        g.n1_loc = g.n1; g.j1_off = 0;
        for (int j1 = 0; j1 < g.n1; ++j1) for (int j2 = 0; j2 < g.nzh; ++j2) {
            const KLine L = make_kline(g, j1, j2);
            for (int j0 = 0; j0 < g.n0; ++j0) {
                const KPoint p = make_kpoint_at(g, j0, j1, j2);
                auto f1 = [](double x, double y, double z) { return x * x + y * y + z * z; };
                auto f2 = [](double x, double y, double z) { const double k2 = x * x + y * y + z * z; return k2 != 0.0 ? 1.0 / k2 : 0.0; };
                const double a1 = sym_even(p, f1), a2 = sym_even(p, f2);
                const double b1 = kline_sym_even(L, j0, [](double k2) { return k2; });
                const double b2 = kline_sym_even(L, j0, [](double k2) { return k2 > 0.0 ? 1.0 / k2 : 0.0; });
                const bool zero = (j0 == 0 && j1 == 0 && j2 == 0);
                if (zero) { if (b1 != 0.0 || b2 != 0.0) worst = 1e9; continue; }
                worst = std::fmax(worst, std::fabs(a1 - b1) / std::fabs(a1));
                worst = std::fmax(worst, std::fabs(a2 - b2) / std::fabs(a2));
            }
        }
    }
    printf("kline worst rel diff %.3g\n", worst);
    return !(worst < 1e-13);
}
