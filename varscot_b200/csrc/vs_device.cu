// varscot_b200/csrc/vs_device.cu — device half of the C ABI in include/varscot_scan.h:
// context, packed-text upload, the scan (extract -> score) and the integer-pipe
// microbenchmarks.  Replaces the index-resident search loop of bidir_mapping.cpp:268,285-295.
// There is NO CPU fallback: every entry point fails with VS_ERR_CUDA / VS_ERR_NODEVICE when no
// sm_100 device is usable.
#include "vs_kernels.cuh"
#include "vs_internal.h"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace vs;

static thread_local std::string g_last_error;

struct vs_ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {};
    // text
    vs_word *d_words = nullptr;
    uint64_t words_cap = 0, n_words = 0, global_base = 0;
    // counters: [0] cand fwd, [1] cand rev, [2] blocks fwd, [3] blocks rev, [4] hits
    unsigned long long *d_cnt = nullptr, *h_cnt = nullptr;
    // candidate stores
    uint32_t *d_planes[2] = {nullptr, nullptr}, *d_pos[2] = {nullptr, nullptr};
    uint64_t blocks_cap[2] = {0, 0};
    // hits
    vs_hit *d_hits = nullptr;
    uint64_t hits_cap = 0, last_n_hits = 0;
    std::string err;
};

static int fail(vs_ctx *c, int code, const std::string &msg)
{
    g_last_error = msg;
    if (c) c->err = msg;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, VS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)

extern "C" const char *vs_last_error(const vs_ctx *ctx)
{
    if (ctx && !ctx->err.empty()) return ctx->err.c_str();
    return g_last_error.c_str();
}

void vs_set_last_error(const char *msg) { g_last_error = msg ? msg : ""; }

extern "C" int vs_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        g_last_error = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        (void)cudaGetLastError();
        return -VS_ERR_NODEVICE;
    }
    return n;
}

template <int K>
static cudaError_t set_score_attr()
{
    return cudaFuncSetAttribute(k_score<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, NPLANES * SCORE_THREADS * 4);
}

extern "C" int vs_ctx_create(int device, vs_ctx **out)
{
    vs_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, VS_ERR_ARG, "vs_ctx_create: out is NULL");
    *out = nullptr;
    int n = vs_device_count();
    if (n <= 0) return fail(nullptr, VS_ERR_NODEVICE, "no CUDA device visible (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(nullptr, VS_ERR_ARG, "vs_ctx_create: device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, VS_ERR_NODEVICE, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                                                  ", this build contains sm_100a code only");
    ctx = new vs_ctx();
    ctx->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_cnt, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_cnt, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = set_score_attr<0>();
    if (e == cudaSuccess) e = set_score_attr<1>();
    if (e == cudaSuccess) e = set_score_attr<2>();
    if (e == cudaSuccess) e = set_score_attr<3>();
    if (e == cudaSuccess) e = set_score_attr<4>();
    if (e == cudaSuccess) e = set_score_attr<5>();
    if (e == cudaSuccess) e = set_score_attr<6>();
    if (e == cudaSuccess) e = set_score_attr<7>();
    if (e == cudaSuccess) e = set_score_attr<8>();
    if (e != cudaSuccess) {
        std::string m = std::string("vs_ctx_create: ") + cudaGetErrorString(e);
        vs_ctx_destroy(ctx);
        return fail(nullptr, VS_ERR_CUDA, m);
    }
    *out = ctx;
    return VS_OK;
}

extern "C" void vs_ctx_destroy(vs_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->device >= 0) cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_words);
    cudaFree(ctx->d_cnt);
    if (ctx->h_cnt) cudaFreeHost(ctx->h_cnt);
    for (int s = 0; s < 2; ++s) { cudaFree(ctx->d_planes[s]); cudaFree(ctx->d_pos[s]); }
    cudaFree(ctx->d_hits);
    for (int i = 0; i < 6; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" void *vs_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void vs_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" int vs_text_upload(vs_ctx *ctx, const vs_word *words, uint64_t n_words, uint64_t global_base)
{
    if (!ctx || (!words && n_words)) return fail(ctx, VS_ERR_ARG, "vs_text_upload: bad arguments");
    if (n_words * 32 + global_base > (1ull << 32))
        return fail(ctx, VS_ERR_ARG, "vs_text_upload: text exceeds 4 Gbases (32-bit positions, as common.h:9-19)");
    CK(cudaSetDevice(ctx->device));
    if (n_words + 1 > ctx->words_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_words); ctx->d_words = nullptr; ctx->words_cap = 0;
        CK(cudaMalloc(&ctx->d_words, (n_words + 1) * sizeof(vs_word)));
        ctx->words_cap = n_words + 1;
    }
    if (n_words) CK(cudaMemcpyAsync(ctx->d_words, words, (n_words + 1) * sizeof(vs_word), cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_words = n_words;
    ctx->global_base = global_base;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->err.clear();
    return VS_OK;
}

static void make_pam(int extra_pam, PamParams &pp)
{
    // forward {GG, GA} + XY; reverse {CC, TC} + revcomp(XY)   (bidir_mapping.cpp:240-247)
    pp.n = 2;
    pp.fx[0] = 2; pp.fy[0] = 2; pp.fx[1] = 2; pp.fy[1] = 0;
    pp.fx[2] = 0; pp.fy[2] = 0;
    if (extra_pam >= 0) { pp.fx[2] = extra_pam / 4; pp.fy[2] = extra_pam % 4; pp.n = 3; }
    for (int j = 0; j < 3; ++j) { pp.rx[j] = 3 - pp.fy[j]; pp.ry[j] = 3 - pp.fx[j]; }
}

template <int K>
static void launch_score(const ScoreArgs &a, cudaStream_t st)
{
    unsigned grid = (unsigned)((a.n_blocks + SCORE_THREADS - 1) / SCORE_THREADS);
    k_score<K><<<grid, SCORE_THREADS, NPLANES * SCORE_THREADS * 4, st>>>(a);
}

static void dispatch_score(int k, const ScoreArgs &a, cudaStream_t st)
{
    switch (k) {
    case 0: launch_score<0>(a, st); break;
    case 1: launch_score<1>(a, st); break;
    case 2: launch_score<2>(a, st); break;
    case 3: launch_score<3>(a, st); break;
    case 4: launch_score<4>(a, st); break;
    case 5: launch_score<5>(a, st); break;
    case 6: launch_score<6>(a, st); break;
    case 7: launch_score<7>(a, st); break;
    default: launch_score<8>(a, st); break;
    }
}

extern "C" int vs_scan(vs_ctx *ctx, const uint8_t *guides, uint32_t n_guides, int k, int extra_pam,
                       vs_hit *out, uint64_t out_cap, uint64_t *n_hits, vs_scan_stats *stats)
{
    if (!ctx) return fail(nullptr, VS_ERR_ARG, "vs_scan: ctx is NULL");
    if (k < 0 || k > VS_MAX_MISMATCHES) return fail(ctx, VS_ERR_ARG, "vs_scan: mismatches must lie between 0 and 8");
    if (extra_pam < -1 || extra_pam > 15) return fail(ctx, VS_ERR_ARG, "vs_scan: extra_pam must be -1 or 4*x+y");
    if (n_guides && !guides) return fail(ctx, VS_ERR_ARG, "vs_scan: guides is NULL");
    if (n_guides >= (1u << 24)) return fail(ctx, VS_ERR_ARG, "vs_scan: at most 2^24-1 guides per call");
    for (uint64_t i = 0; i < (uint64_t)n_guides * VS_GLEN; ++i)
        if (guides[i] > 3) return fail(ctx, VS_ERR_ARG, "vs_scan: guide codes must be 0..3");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    vs_scan_stats S;
    memset(&S, 0, sizeof(S));
    if (n_hits) *n_hits = 0;
    ctx->last_n_hits = 0;
    if (ctx->n_words == 0 || n_guides == 0) { if (stats) *stats = S; return VS_OK; }

    PamParams pp;
    make_pam(extra_pam, pp);
    const uint32_t n_tiles = (uint32_t)((ctx->n_words + TILE_WORDS - 1) / TILE_WORDS);

    // pattern tables: byte offset of the selected mismatch plane per position, per strand pass
    const uint32_t n_chunks = (n_guides + PAT_CHUNK - 1) / PAT_CHUNK;
    std::vector<uint32_t> pat((size_t)2 * n_chunks * PAT_CHUNK * PAT_STRIDE, 0u);
    for (int s = 0; s < 2; ++s)
        for (uint32_t g = 0; g < n_guides; ++g) {
            uint32_t *dst = pat.data() + (((size_t)s * n_chunks + g / PAT_CHUNK) * PAT_CHUNK + g % PAT_CHUNK) * PAT_STRIDE;
            const uint8_t *gd = guides + (size_t)g * VS_GLEN;
            for (int i = 0; i < VS_GLEN; ++i) {
                int b = s ? 3 - gd[VS_GLEN - 1 - i] : gd[i];      // reverse pass scores revcomp(guide), bidir_mapping.cpp:293
                dst[i] = (uint32_t)(4 * i + b) * SCORE_THREADS * 4u;
            }
        }

    // candidate stores: sized for the expected PAM density, regrown (and the extraction repeated) on overflow
    auto ensure_blocks = [&](int s, uint64_t need) -> int {
        if (need <= ctx->blocks_cap[s]) return VS_OK;
        cudaFree(ctx->d_planes[s]); cudaFree(ctx->d_pos[s]);
        ctx->d_planes[s] = ctx->d_pos[s] = nullptr; ctx->blocks_cap[s] = 0;
        CK(cudaMalloc(&ctx->d_planes[s], need * BLK_WORDS * sizeof(uint32_t)));
        CK(cudaMalloc(&ctx->d_pos[s], need * 32 * sizeof(uint32_t)));
        ctx->blocks_cap[s] = need;
        return VS_OK;
    };
    {
        uint64_t est = (uint64_t)((double)ctx->n_words * pp.n / 16.0 * 1.15) + n_tiles + 1024;
        for (int s = 0; s < 2; ++s)
            if (ctx->blocks_cap[s] == 0) { int r = ensure_blocks(s, est); if (r != VS_OK) return r; }
    }
    CK(cudaEventRecord(ctx->ev[0], st));
    uint64_t nb[2] = {0, 0};
    for (int attempt = 0; attempt < 2; ++attempt) {
        CK(cudaMemsetAsync(ctx->d_cnt, 0, 8 * sizeof(unsigned long long), st));
        k_extract<<<n_tiles, TILE_THREADS, 0, st>>>(ctx->d_words, ctx->n_words, ctx->global_base, pp,
                                                    ctx->d_planes[0], ctx->d_pos[0], ctx->blocks_cap[0],
                                                    ctx->d_planes[1], ctx->d_pos[1], ctx->blocks_cap[1], ctx->d_cnt);
        S.launches += 1;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        S.n_cand_fwd = ctx->h_cnt[0]; S.n_cand_rev = ctx->h_cnt[1];
        nb[0] = S.n_blocks_fwd = ctx->h_cnt[2]; nb[1] = S.n_blocks_rev = ctx->h_cnt[3];
        if (nb[0] <= ctx->blocks_cap[0] && nb[1] <= ctx->blocks_cap[1]) break;
        if (attempt == 1) return fail(ctx, VS_ERR_CUDA, "vs_scan: candidate store overflow after regrow");
        for (int s = 0; s < 2; ++s) { int r = ensure_blocks(s, nb[s] + nb[s] / 32 + 64); if (r != VS_OK) return r; }
    }
    CK(cudaEventRecord(ctx->ev[3], st));

    if (!ctx->d_hits) {
        uint64_t cap = 1u << 20;
        CK(cudaMalloc(&ctx->d_hits, cap * sizeof(vs_hit)));
        ctx->hits_cap = cap;
    }
    uint64_t found = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        CK(cudaMemsetAsync(ctx->d_cnt + 4, 0, sizeof(unsigned long long), st));
        for (int s = 0; s < 2; ++s) {
            if (nb[s] == 0) continue;
            for (uint32_t c = 0; c < n_chunks; ++c) {
                uint32_t np = n_guides - c * PAT_CHUNK;
                if (np > (uint32_t)PAT_CHUNK) np = PAT_CHUNK;
                const uint32_t *src = pat.data() + ((size_t)s * n_chunks + c) * PAT_CHUNK * PAT_STRIDE;
                CK(cudaMemcpyToSymbolAsync(c_pat, src, (size_t)np * PAT_STRIDE * sizeof(uint32_t), 0, cudaMemcpyHostToDevice, st));
                ScoreArgs a;
                a.planes = ctx->d_planes[s]; a.pos = ctx->d_pos[s]; a.n_blocks = nb[s];
                a.n_pat = np; a.guide_base = c * PAT_CHUNK; a.strand = (uint32_t)s;
                a.hits = ctx->d_hits; a.n_hits = ctx->d_cnt + 4; a.hit_cap = ctx->hits_cap;
                dispatch_score(k, a, st);
                if (attempt == 0) { S.launches++; S.score_launches++; }
            }
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[4], st));
        CK(cudaMemcpyAsync(ctx->h_cnt + 4, ctx->d_cnt + 4, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        found = ctx->h_cnt[4];
        if (found <= ctx->hits_cap) break;
        if (attempt == 1) return fail(ctx, VS_ERR_CUDA, "vs_scan: hit buffer overflow after regrow");
        // device hit buffer too small: grow and score again (candidates are kept)
        cudaFree(ctx->d_hits); ctx->d_hits = nullptr; ctx->hits_cap = 0;
        uint64_t cap = found + found / 8 + 1024;
        CK(cudaMalloc(&ctx->d_hits, cap * sizeof(vs_hit)));
        ctx->hits_cap = cap;
    }
    ctx->last_n_hits = found;
    S.n_hits = found;
    if (n_hits) *n_hits = found;
    uint64_t ncopy = found < out_cap ? found : out_cap;
    if (out && ncopy) CK(cudaMemcpyAsync(out, ctx->d_hits, ncopy * sizeof(vs_hit), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[5], st));
    CK(cudaStreamSynchronize(st));
    S.count_ms = 0.f;     // the separate count pass is gone: k_extract claims block ranges with atomics
    CK(cudaEventElapsedTime(&S.extract_ms, ctx->ev[0], ctx->ev[3]));
    CK(cudaEventElapsedTime(&S.score_ms, ctx->ev[3], ctx->ev[4]));
    CK(cudaEventElapsedTime(&S.total_ms, ctx->ev[0], ctx->ev[5]));
    if (stats) *stats = S;
    ctx->err.clear();
    if (found > out_cap) return fail(ctx, VS_ERR_OVERFLOW, "vs_scan: caller hit buffer too small; use vs_scan_fetch");
    return VS_OK;
}

extern "C" int vs_scan_fetch(vs_ctx *ctx, vs_hit *out, uint64_t out_cap, uint64_t *n_hits)
{
    if (!ctx) return fail(nullptr, VS_ERR_ARG, "vs_scan_fetch: ctx is NULL");
    if (n_hits) *n_hits = ctx->last_n_hits;
    if (ctx->last_n_hits > out_cap) return fail(ctx, VS_ERR_OVERFLOW, "vs_scan_fetch: buffer too small");
    CK(cudaSetDevice(ctx->device));
    if (ctx->last_n_hits) {
        if (!out) return fail(ctx, VS_ERR_ARG, "vs_scan_fetch: out is NULL");
        CK(cudaMemcpyAsync(out, ctx->d_hits, ctx->last_n_hits * sizeof(vs_hit), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->err.clear();
    return VS_OK;
}

extern "C" int vs_measure_int_peaks(vs_ctx *ctx, double *lop3_per_s, double *lds_words_per_s)
{
    if (!ctx) return fail(nullptr, VS_ERR_ARG, "vs_measure_int_peaks: ctx is NULL");
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    const int grid = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    uint32_t *d_out = nullptr;
    CK(cudaMalloc(&d_out, (size_t)grid * threads * sizeof(uint32_t)));
    cudaStream_t st = ctx->stream;
    float best_lop = 1e30f, best_lds = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        float ms = 0;
        CK(cudaEventRecord(ctx->ev[0], st));
        k_peak_lop3<<<grid, threads, 0, st>>>(d_out, iters);
        CK(cudaEventRecord(ctx->ev[1], st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        if (rep && ms < best_lop) best_lop = ms;
        CK(cudaEventRecord(ctx->ev[0], st));
        k_peak_lds<<<grid, threads, 0, st>>>(d_out, iters);
        CK(cudaEventRecord(ctx->ev[1], st));
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        if (rep && ms < best_lds) best_lds = ms;
    }
    CK(cudaGetLastError());
    cudaFree(d_out);
    const double nthreads = (double)grid * threads;
    if (lop3_per_s) *lop3_per_s = nthreads * iters * 64.0 / (best_lop * 1e-3);
    if (lds_words_per_s) *lds_words_per_s = nthreads * iters * 8.0 / (best_lds * 1e-3);
    return VS_OK;
}
