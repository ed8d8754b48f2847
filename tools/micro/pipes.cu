// tools/micro/pipes.cu — does the SM overlap alu-pipe (LOP3) and shared-memory (LDS) instructions of the same warps?
// Times kernels with NL LOP3 and NS LDS per inner iteration (independent chains, 8 warps per scheduler) and prints
// cycles per iteration per SM sub-partition, to compare t(mix) with t(lop3 only) + t(lds only) and max(...).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}

template <int NL, int NS, int NI>
__global__ void __launch_bounds__(256) k_mix(uint32_t *out, int iters)
{
    __shared__ uint32_t s[256 * 8];
    for (int i = threadIdx.x; i < 256 * 8; i += 256) s[i] = i * 2654435761u;
    __syncthreads();
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * (2 * i + 3) + i;
    uint32_t acc = 0, im = threadIdx.x | 1;
    const volatile uint32_t *p = s + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        uint32_t v[NS > 0 ? NS : 1];
#pragma unroll
        for (int u = 0; u < NS; ++u) v[u] = p[(u & 7) * 256];
#pragma unroll
        for (int u = 0; u < NL; ++u) a[u & 7] = lop3<0x96>(a[u & 7], a[(u + 1) & 7], a[(u + 2) & 7]);
#pragma unroll
        for (int u = 0; u < NI; ++u) im = im * 0x9E3779B1u + it;          // IMAD (fma pipe)
#pragma unroll
        for (int u = 0; u < NS; ++u) acc ^= v[u];                          // NS extra LOP (xor) ops to consume the loads
    }
    uint32_t r = acc ^ im;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NL, int NS, int NI>
static void run(const char *name, uint32_t *d_out, int sms)
{
    const int grid = sms * 8, iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_mix<NL, NS, NI><<<grid, 256>>>(d_out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    // 8 CTAs x 8 warps = 64 warps per SM = 16 per sub-partition; cycles per iteration per warp-slot at 1.965 GHz
    const double cyc = best * 1e-3 * 1.965e9 / iters / 16.0;
    printf("%-28s LOP3 %2d (+%d xor) LDS %2d IMAD %2d : %7.3f ms  %6.2f SMSP-cycles per warp-iteration\n", name, NL, NS, NS, NI, best, cyc);
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    uint32_t *d_out; cudaMalloc(&d_out, (size_t)prop.multiProcessorCount * 8 * 256 * 4);
    const int sms = prop.multiProcessorCount;
    run<32, 0, 0>("lop3 only", d_out, sms);
    run<0, 16, 0>("lds only", d_out, sms);
    run<32, 16, 0>("lop3 + lds", d_out, sms);
    run<16, 16, 0>("lop3/2 + lds", d_out, sms);
    run<32, 8, 0>("lop3 + lds/2", d_out, sms);
    run<0, 0, 32>("imad only", d_out, sms);
    run<32, 0, 32>("lop3 + imad", d_out, sms);
    run<32, 16, 16>("lop3 + lds + imad/2", d_out, sms);
    cudaFree(d_out);
    return 0;
}
