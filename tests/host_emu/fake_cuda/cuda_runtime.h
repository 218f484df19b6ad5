#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __align__(n) alignas(n)
#define __restrict__
#define __launch_bounds__(...)
struct alignas(16) double2 { double x, y; };
inline double2 make_double2(double x, double y) { return {x, y}; }
struct dim3 { unsigned x = 1, y = 1, z = 1; };
extern thread_local std::barrier<>* g_warp_barrier;
extern thread_local std::barrier<>* g_block_barrier;
extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
inline void __syncwarp() { g_warp_barrier->arrive_and_wait(); }
inline void __syncthreads() { g_block_barrier->arrive_and_wait(); }
