// tools/micro/h2d8.cu — microbenchmark: what is the box's ceiling for N concurrent pinned-host -> device copies?
// (the end-to-end 1 -> 8 GPU curve is bounded by 8 ranks uploading their text at the same time: 172 GB/s
// aggregate, 21.5 GB/s per GPU against 50 GB/s alone.  Is that the host's ceiling or a placement problem?)
// One host thread per GPU allocates and first-touches its own pinned buffer (so the pages sit on the node the thread runs
// on), then the copies are timed (CUDA events, best of 5) one GPU at a time and all GPUs together, for two buffer sizes
// (the whole config-3 text and one rank's 1/8 shard) and two allocation flavours (cudaMallocHost; malloc + first touch +
// cudaHostRegister).  Prints one JSON line.
// build: nvcc -O2 -std=c++17 -o build/h2d8 tools/micro/h2d8.cu -lpthread      run: build/h2d8 [n_gpus]
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static std::atomic<int> g_arrived{0};
static std::atomic<int> g_phase{0};
static void barrier(int n)
{
    const int phase = g_phase.load();
    if (g_arrived.fetch_add(1) + 1 == n) { g_arrived.store(0); g_phase.fetch_add(1); }
    else while (g_phase.load() == phase) std::this_thread::yield();
}

struct Res { double solo = 0, together = 0; };

int main(int argc, char **argv)
{
    int n = 0;
    cudaGetDeviceCount(&n);
    if (argc > 1) n = std::min(n, atoi(argv[1]));
    if (n <= 0) { printf("{\"error\": \"no device\"}\n"); return 1; }
    const size_t sizes[2] = {850ull << 20, 106ull << 20};
    const char *flavour[2] = {"cudaMallocHost", "malloc+touch+cudaHostRegister"};
    std::string out = "{\"n_gpus\": " + std::to_string(n) + ", \"runs\": [";
    bool first = true;
    for (int fl = 0; fl < 2; ++fl)
        for (int sz = 0; sz < 2; ++sz) {
            const size_t bytes = sizes[sz];
            std::vector<Res> res(n);
            std::vector<std::thread> th;
            g_arrived = 0; g_phase = 0;
            for (int g = 0; g < n; ++g)
                th.emplace_back([&, g] {
                    cudaSetDevice(g);
                    char *h = nullptr, *d = nullptr;
                    if (fl == 0) cudaMallocHost(&h, bytes);
                    else { h = (char *)aligned_alloc(4096, bytes); memset(h, 1, bytes); cudaHostRegister(h, bytes, cudaHostRegisterDefault); }
                    if (fl == 0) memset(h, 1, bytes);
                    cudaMalloc(&d, bytes);
                    cudaStream_t st; cudaStreamCreate(&st);
                    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
                    auto copy_ms = [&]() { float ms = 0; cudaEventRecord(a, st); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st); cudaEventRecord(b, st);
                                           cudaStreamSynchronize(st); cudaEventElapsedTime(&ms, a, b); return (double)ms; };
                    copy_ms();
                    barrier(n);
                    for (int turn = 0; turn < n; ++turn) {             // one GPU at a time
                        if (turn == g) { double best = 1e30; for (int r = 0; r < 5; ++r) best = std::min(best, copy_ms()); res[g].solo = bytes / best / 1e6; }
                        barrier(n);
                    }
                    double best = 1e30;                                // all together: every repetition starts at a barrier
                    for (int r = 0; r < 5; ++r) { barrier(n); best = std::min(best, copy_ms()); }
                    res[g].together = bytes / best / 1e6;
                    barrier(n);
                    if (fl == 0) cudaFreeHost(h); else { cudaHostUnregister(h); free(h); }
                    cudaFree(d);
                });
            for (auto &t : th) t.join();
            double agg = 0, solo_min = 1e30, solo_max = 0, tog_min = 1e30;
            for (auto &r : res) { agg += r.together; solo_min = std::min(solo_min, r.solo); solo_max = std::max(solo_max, r.solo); tog_min = std::min(tog_min, r.together); }
            char buf[512];
            snprintf(buf, sizeof buf, "%s{\"alloc\": \"%s\", \"mb\": %zu, \"solo_gbs_min\": %.1f, \"solo_gbs_max\": %.1f, \"together_gbs_sum\": %.1f, \"together_gbs_min\": %.1f}",
                     first ? "" : ", ", flavour[fl], bytes >> 20, solo_min, solo_max, agg, tog_min);
            out += buf; first = false;
        }
    out += "]}";
    printf("%s\n", out.c_str());
    return 0;
}
