// tests/cuda_on_host.h — TEST INFRASTRUCTURE: the shim under which varscot_b200/csrc/vs_kernels.cuh compiles with g++, so that
// the device code of the scan can be exercised in the CPU test suite (tests/cpu_kernel_units.cpp, tests/cpu_scan_emulator.cpp).
// Two ways to run a kernel:
//   launch(grid, block, f)      one thread after the other — for kernels whose threads never cooperate
//                               (the mask kernels, k_resolve_hits); __syncthreads is a no-op
//   launch_cta(grid, block, f)  one OS thread per CUDA thread of a CTA, CTAs one after the other — for k_extract, k_score
//                               and the contig-start kernels (__syncthreads is a barrier, __shfl_up_sync exchanges through
//                               a buffer, __shared__ arrays are function-local statics); a warp vote sees only its own
//                               lane, which is exact for k_score: a lane whose own stage-A result is zero never hits
// Nothing in the product includes this file; the product has no CPU path.
#pragma once
#include <barrier>
#include <cstdint>
#include <thread>
#include <vector>

#define VS_HOST_UNIT_TEST
#define __host__
#define __device__
#define __forceinline__ inline
#define __global__
#define __constant__ static
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct Idx3 { unsigned x, y, z; };
static thread_local Idx3 threadIdx, blockIdx;
static Idx3 blockDim, gridDim;
static std::barrier<> *g_cta_barrier = nullptr;          // set by launch_cta
static uint32_t g_shfl[1024];

static inline void __syncthreads() { if (g_cta_barrier) g_cta_barrier->arrive_and_wait(); }
// every thread of the CTA executes the same shuffles (true for k_extract), so a CTA-wide exchange emulates the warp's
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned delta)
{
    g_shfl[threadIdx.x] = v;
    __syncthreads();
    const uint32_t r = (threadIdx.x & 31u) >= delta ? g_shfl[threadIdx.x - delta] : v;
    __syncthreads();
    return r;
}
// warp vote of a one-thread "warp" (launch): a lane whose own stage-A result is zero can never produce a hit, so voting alone is exact
static inline int __any_sync(unsigned, int pred) { return pred; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s)
{
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t sel)
{
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline uint32_t __ldg(const uint32_t *p) { return *p; }
// (k_score runs under launch_cta: its threads are concurrent OS threads)
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline uint16_t __ldg(const uint16_t *p) { return *p; }
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint64_t max(uint64_t a, uint64_t b) { return a > b ? a : b; }
static inline uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }

namespace vs { extern uint32_t sm[]; }                   // k_score's dynamic shared memory; defined by the including test

template <class F>
static void launch(unsigned grid_x, unsigned grid_y, unsigned block, F kernel)
{
    gridDim = Idx3{grid_x, grid_y, 1}; blockDim = Idx3{block, 1, 1};
    for (unsigned by = 0; by < grid_y; ++by)
        for (unsigned bx = 0; bx < grid_x; ++bx)
            for (unsigned t = 0; t < block; ++t) {
                blockIdx = Idx3{bx, by, 0}; threadIdx = Idx3{t, 0, 0};
                kernel();
            }
}

template <class F>
static void launch_cta(unsigned grid_x, unsigned block, F kernel, unsigned grid_y = 1)
{
    gridDim = Idx3{grid_x, grid_y, 1}; blockDim = Idx3{block, 1, 1};
    std::barrier<> bar((std::ptrdiff_t)block);
    g_cta_barrier = &bar;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < block; ++t)
        th.emplace_back([&, t] {
            for (unsigned by = 0; by < grid_y; ++by)
                for (unsigned bx = 0; bx < grid_x; ++bx) {
                    blockIdx = Idx3{bx, by, 0}; threadIdx = Idx3{t, 0, 0};
                    kernel();
                    bar.arrive_and_wait();                // the next CTA reuses the __shared__ statics
                }
        });
    for (auto &x : th) x.join();
    g_cta_barrier = nullptr;
}
