// varscot_b200/csrc/vs_device.cu — device half of the C ABI in include/varscot_scan.h:
// context, packed-text upload, the scan (extract -> score) and the integer-pipe
// microbenchmarks.  Replaces the index-resident search loop of bidir_mapping.cpp:268,285-295.
// There is NO CPU fallback: every entry point fails with VS_ERR_CUDA / VS_ERR_NODEVICE when no
// sm_100 device is usable.
#include "vs_kernels.cuh"
#include "vs_internal.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace vs;

static thread_local std::string g_last_error;

constexpr uint64_t DEFAULT_CHUNK_WORDS = 8ull << 20;      // 256 Mi bases per pipeline chunk

struct vs_ctx {
    int device = -1;
    cudaStream_t stream = nullptr;       // scoring + everything ordered with it
    cudaStream_t exs = nullptr;          // extraction of the next chunk (overlaps the scoring of the current one)
    cudaStream_t copy = nullptr;         // H2D of the next chunk (copies only, so that the copy engine never waits for a kernel)
    cudaStream_t prep = nullptr;         // small kernels that turn the uploaded mask source of a chunk into its window masks
    cudaEvent_t ev_copied = nullptr, ev_zeroed = nullptr;
    cudaEvent_t ev[6] = {};
    std::vector<cudaEvent_t> ev_pool;    // per chunk: copied, extract start, extract done, score start, score done
    uint64_t chunk_words = DEFAULT_CHUNK_WORDS;
    // resident text shard: device word 0 = global word first_word
    vs_bases *d_bases = nullptr;
    vs_masks *d_masks = nullptr;
    uint64_t words_cap = 0, n_words = 0, first_word = 0;
    vs_mask_entry *d_sparse = nullptr;
    uint64_t sparse_cap = 0;
    // compact mask source path: the N and contig-end planes of the shard (+ halo word) and a staging area for their runs
    uint32_t *d_nm = nullptr, *d_em = nullptr;
    uint8_t *d_emcode = nullptr, *d_dense = nullptr;     // code bytes of the shard's contig-end plane; coded-block flags of the whole text
    uint64_t planes_cap = 0, dense_cap = 0;
    vs_plane_run *d_runs = nullptr;
    uint64_t runs_cap = 0;
    // counters: per chunk [0] cand fwd, [1] cand rev, [2] blocks fwd, [3] blocks rev; then one hit counter
    unsigned long long *d_cnt = nullptr, *h_cnt = nullptr;
    uint64_t cnt_chunks = 0;
    // candidate stores: two buffers (chunk parity) x two strands
    uint32_t *d_planes[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}, *d_pos[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    uint64_t blocks_cap = 0;
    // pattern tables
    uint32_t *d_pat = nullptr, *h_pat = nullptr;
    uint64_t pat_cap = 0;
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

void vs_warmup_device(int device)
{
    if (cudaSetDevice(device) == cudaSuccess) cudaFree(nullptr);
    (void)cudaGetLastError();
}

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

// k_score's dynamic shared memory (at most 92 planes x 128 threads x 4 B = 46 KB, for k = 8) fits the default 48 KB
// limit, so no cudaFuncSetAttribute is needed — and none of the nine k_score<K> variants is loaded before it is used.
static_assert(NPLANES * SCORE_THREADS * 4 <= 48 * 1024, "k_score needs cudaFuncAttributeMaxDynamicSharedMemorySize above 48 KB");

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
    // The copy and mask-preparation streams get the highest priority: the small kernels that build the window masks of
    // the next chunk must not queue behind the scoring grid of the current one.  The optional concurrent-extraction
    // stream gets the lowest.
    int prio_lo = 0, prio_hi = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi < prio_lo ? prio_hi + 1 : prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->copy, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->prep, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_zeroed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->exs, cudaStreamNonBlocking, prio_lo);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
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
    if (ctx->copy) cudaStreamSynchronize(ctx->copy);
    if (ctx->prep) cudaStreamSynchronize(ctx->prep);
    if (ctx->exs) cudaStreamSynchronize(ctx->exs);
    cudaFree(ctx->d_bases); cudaFree(ctx->d_masks); cudaFree(ctx->d_sparse);
    cudaFree(ctx->d_nm); cudaFree(ctx->d_em); cudaFree(ctx->d_runs); cudaFree(ctx->d_emcode); cudaFree(ctx->d_dense);
    cudaFree(ctx->d_cnt);
    if (ctx->h_cnt) cudaFreeHost(ctx->h_cnt);
    for (int b = 0; b < 2; ++b) for (int s = 0; s < 2; ++s) { cudaFree(ctx->d_planes[b][s]); cudaFree(ctx->d_pos[b][s]); }
    cudaFree(ctx->d_pat);
    if (ctx->h_pat) cudaFreeHost(ctx->h_pat);
    cudaFree(ctx->d_hits);
    for (int i = 0; i < 6; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy) cudaStreamDestroy(ctx->copy);
    if (ctx->prep) cudaStreamDestroy(ctx->prep);
    if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
    if (ctx->ev_zeroed) cudaEventDestroy(ctx->ev_zeroed);
    if (ctx->exs) cudaStreamDestroy(ctx->exs);
    delete ctx;
}

extern "C" int vs_ctx_set_chunk_words(vs_ctx *ctx, uint64_t chunk_words)
{
    if (!ctx || chunk_words == 0) return fail(ctx, VS_ERR_ARG, "vs_ctx_set_chunk_words: bad arguments");
    ctx->chunk_words = chunk_words;
    return VS_OK;
}

extern "C" void *vs_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void vs_host_free(void *p) { if (p) cudaFreeHost(p); }

static void make_pam(int extra_pam, PamParams &pp)
{
    // forward {GG, GA} + XY; reverse {CC, TC} + revcomp(XY)   (bidir_mapping.cpp:240-247)
    pp.n = 2;
    pp.fx[0] = 2; pp.fy[0] = 2; pp.fx[1] = 2; pp.fy[1] = 0;
    pp.fx[2] = 0; pp.fy[2] = 0;
    if (extra_pam >= 0) { pp.fx[2] = extra_pam / 4; pp.fy[2] = extra_pam % 4; pp.n = 3; }
    for (int j = 0; j < 3; ++j) { pp.rx[j] = 3 - pp.fy[j]; pp.ry[j] = 3 - pp.fx[j]; }
}

// grid = forward CTAs then reverse CTAs, sized by the capacity of the candidate stores (CTAs past the claimed block
// count of their strand exit at once)
template <int K>
static void launch_score(const ScoreArgs &a, cudaStream_t st)
{
    k_score<K><<<2 * a.ctas_per_strand, SCORE_THREADS, score_smem_planes(K) * SCORE_THREADS * 4, st>>>(a);
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

// Pipeline chunks of a shard of n_words words: bounds[0] = 0 < ... < bounds[n] = n_words, chunk_words apart.  For a
// streamed scan the last chunk is cut into 1/2, 1/4, 1/8, 1/8: the scan of a chunk can only start when its copy has
// arrived, so what follows the last copy — the tail of the pipeline — shrinks to the scan of an eighth of a chunk.
static std::vector<uint64_t> chunk_plan(uint64_t n_words, uint64_t chunk_words, bool taper)
{
    std::vector<uint64_t> b{0};
    for (uint64_t c0 = 0; c0 < n_words; c0 += chunk_words) b.push_back(std::min(n_words, c0 + chunk_words));
    if (taper && b.size() >= 3) {                  // at least two chunks: split the last one
        const uint64_t c0 = b[b.size() - 2], len = n_words - c0;
        if (len >= 8192) {
            b.pop_back();
            for (uint64_t cut : {len / 2, len / 2 + len / 4, len / 2 + len / 4 + len / 8}) b.push_back(c0 + (cut & ~255ull));
            b.push_back(n_words);
        }
    }
    return b;
}

// ---- text residency ------------------------------------------------------------------------------
static int ensure_text_buffers(vs_ctx *ctx, uint64_t n_words)
{
    if (n_words + 1 > ctx->words_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->copy));
        cudaFree(ctx->d_bases); cudaFree(ctx->d_masks);
        ctx->d_bases = nullptr; ctx->d_masks = nullptr; ctx->words_cap = 0;
        CK(cudaMalloc(&ctx->d_bases, (n_words + 1) * sizeof(vs_bases)));
        CK(cudaMalloc(&ctx->d_masks, (n_words + 1) * sizeof(vs_masks)));
        ctx->words_cap = n_words + 1;
    }
    return VS_OK;
}

static inline bool has_mask_source(const vs_text_view *t) { return t->em_code && t->em_dense; }

// runs of a sorted run list that overlap the word range [lo, hi): index range [first, last)
static void runs_in_range(const vs_plane_run *r, uint64_t n, uint64_t lo, uint64_t hi, uint64_t &first, uint64_t &last)
{
    first = (uint64_t)(std::partition_point(r, r + n, [&](const vs_plane_run &x) { return (uint64_t)x.word + x.count <= lo; }) - r);
    last = (uint64_t)(std::partition_point(r + first, r + n, [&](const vs_plane_run &x) { return (uint64_t)x.word < hi; }) - r);
}

constexpr uint64_t SKIP_MIN_WORDS = 16384;      // all-N runs at least this long are not copied: their bases are zero

// Enqueue the H2D of chunk [c0, c1) (device word indices) on the copy stream.
//  * view with a compact mask source: bases (minus long all-N runs, which are zero-filled on the device), the N plane as
//    runs, the contig-end plane dense or as runs per block, then k_masks_from_planes computes the chunk's masks;
//  * otherwise the masks travel dense, or — when the view carries a sparse list and it is smaller — as memset + sparse
//    entries + a scatter kernel.
// `ready` returns the stream on which the chunk's data is complete (record the chunk's "arrived" event there).
static int enqueue_chunk_copy(vs_ctx *ctx, const vs_text_view *t, uint64_t first_word, uint64_t c0, uint64_t c1,
                              uint64_t &staged, uint64_t &bytes, uint32_t &launches, cudaStream_t &ready)
{
    cudaStream_t cs = ctx->copy;
    ready = cs;
    const uint64_t n = c1 - c0, g0 = first_word + c0;
    // bases [c0, c1] incl. the halo word; word c0 of a later chunk already arrived as the previous chunk's halo
    const uint64_t skip = c0 ? 1 : 0;
    if (has_mask_source(t)) {
        const uint64_t lo = g0, hi = g0 + n + 1;                 // plane words needed: the chunk and its halo word
        uint64_t r0 = 0, r1 = 0;
        runs_in_range(t->nm_runs, t->n_nm_runs, lo, hi, r0, r1);
        // bases, skipping long all-N runs
        uint64_t w = g0 + skip;
        auto copy_bases = [&](uint64_t a, uint64_t b) -> int {   // global words [a, b)
            if (b <= a) return VS_OK;
            CK(cudaMemcpyAsync(ctx->d_bases + (a - first_word), t->bases + a, (b - a) * sizeof(vs_bases), cudaMemcpyHostToDevice, cs));
            bytes += (b - a) * sizeof(vs_bases);
            return VS_OK;
        };
        int r;
        std::vector<std::pair<uint64_t, uint64_t>> skipped;
        for (uint64_t i = r0; i < r1; ++i) {
            const vs_plane_run &x = t->nm_runs[i];
            if (x.value != ~0u || x.count < SKIP_MIN_WORDS) continue;
            const uint64_t a = std::max<uint64_t>(x.word, w), b = std::min<uint64_t>((uint64_t)x.word + x.count, hi);
            if (b <= a) continue;
            if ((r = copy_bases(w, a)) != VS_OK) return r;
            skipped.emplace_back(a, b);                      // zero-filled on the preparation stream below
            w = b;
        }
        if ((r = copy_bases(w, hi)) != VS_OK) return r;
        // N plane and contig-end plane (zeroed by begin_mask_source): the run lists go to the staging area, the code bytes of
        // coded blocks to their own buffer; the expand, fill and mask kernels run on the preparation stream
        vs_plane_run *nm_dst = nullptr, *em_dst = nullptr;
        uint64_t nm_n = 0, em_n = 0;
        auto stage = [&](const vs_plane_run *runs, uint64_t a, uint64_t b, vs_plane_run *&dst, uint64_t &cnt) -> int {
            cnt = b > a ? b - a : 0;
            if (!cnt) return VS_OK;
            if (staged + cnt > ctx->runs_cap) return fail(ctx, VS_ERR_CUDA, "run staging area too small");
            dst = ctx->d_runs + staged;
            CK(cudaMemcpyAsync(dst, runs + a, cnt * sizeof(vs_plane_run), cudaMemcpyHostToDevice, cs));
            staged += cnt;
            bytes += cnt * sizeof(vs_plane_run);
            return VS_OK;
        };
        auto longest_run = [&](const vs_plane_run *runs, uint64_t a, uint64_t b) {     // clipped to [lo, hi): sets grid.y of the fill
            uint64_t m = 1;
            for (uint64_t i = a; i < b; ++i)
                m = std::max<uint64_t>(m, std::min<uint64_t>((uint64_t)runs[i].word + runs[i].count, hi) - std::max<uint64_t>(runs[i].word, lo));
            return m;
        };
        const uint64_t nm_long = longest_run(t->nm_runs, r0, r1);
        if ((r = stage(t->nm_runs, r0, r1, nm_dst, nm_n)) != VS_OK) return r;
        uint64_t span0 = 0, span1 = 0;       // pending span [span0, span1) of coded blocks, in global words
        bool any_coded = false;
        auto flush_span = [&]() -> int {
            if (span1 > span0) {
                CK(cudaMemcpyAsync(ctx->d_emcode + (span0 - first_word), t->em_code + span0, span1 - span0, cudaMemcpyHostToDevice, cs));
                bytes += span1 - span0;
                any_coded = true;
            }
            span0 = span1 = 0;
            return VS_OK;
        };
        for (uint64_t b = lo / VS_EM_BLOCK; b * VS_EM_BLOCK < hi; ++b) {
            if (!t->em_dense[b]) { if ((r = flush_span()) != VS_OK) return r; continue; }
            const uint64_t a = std::max<uint64_t>(b * VS_EM_BLOCK, lo), e = std::min<uint64_t>((b + 1) * VS_EM_BLOCK, hi);
            if (span1 == a && span1 > span0) span1 = e; else { if ((r = flush_span()) != VS_OK) return r; span0 = a; span1 = e; }
        }
        if ((r = flush_span()) != VS_OK) return r;
        runs_in_range(t->em_runs, t->n_em_runs, lo, hi, r0, r1);
        const uint64_t em_long = longest_run(t->em_runs, r0, r1);
        if ((r = stage(t->em_runs, r0, r1, em_dst, em_n)) != VS_OK) return r;
        // everything of this chunk is on its way: hand over to the preparation stream
        cudaStream_t ps = ctx->prep;
        CK(cudaEventRecord(ctx->ev_copied, cs));
        CK(cudaStreamWaitEvent(ps, ctx->ev_copied, 0));
        for (const auto &sk : skipped) CK(cudaMemsetAsync(ctx->d_bases + (sk.first - first_word), 0, (sk.second - sk.first) * sizeof(vs_bases), ps));
        if (nm_n) {
            k_fill_runs<<<dim3((unsigned)((nm_n + 7) / 8), (unsigned)((nm_long + FILL_SEG - 1) / FILL_SEG)), 256, 0, ps>>>(nm_dst, nm_n, lo, hi, first_word, ctx->d_nm);
            launches++;
        }
        if (any_coded) {
            k_expand_em_code<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ps>>>(ctx->d_emcode + c0, ctx->d_dense, lo, n + 1, ctx->d_em + c0);
            launches++;
        }
        if (em_n) {
            k_fill_runs<<<dim3((unsigned)((em_n + 7) / 8), (unsigned)((em_long + FILL_SEG - 1) / FILL_SEG)), 256, 0, ps>>>(em_dst, em_n, lo, hi, first_word, ctx->d_em);
            launches++;
        }
        k_masks_from_planes<<<(unsigned)((n + 255) / 256), 256, 0, ps>>>(ctx->d_nm + c0, ctx->d_em + c0, n, ctx->d_masks + c0);
        launches++;
        ready = ps;
        return VS_OK;
    }
    CK(cudaMemcpyAsync(ctx->d_bases + c0 + skip, t->bases + g0 + skip, (n + 1 - skip) * sizeof(vs_bases), cudaMemcpyHostToDevice, cs));
    bytes += (n + 1 - skip) * sizeof(vs_bases);
    bool sparse = false;
    uint64_t e0 = 0, e1 = 0;
    if (t->sparse) {
        const vs_mask_entry *sb = t->sparse, *se = t->sparse + t->n_sparse;
        auto lb = [&](uint64_t w) { return (uint64_t)(std::lower_bound(sb, se, w, [](const vs_mask_entry &x, uint64_t v) { return x.word < v; }) - sb); };
        e0 = lb(g0); e1 = lb(g0 + n);
        sparse = (e1 - e0) * sizeof(vs_mask_entry) < n * sizeof(vs_masks) * 3 / 4 && staged + (e1 - e0) <= ctx->sparse_cap;
    }
    if (sparse) {
        CK(cudaMemsetAsync(ctx->d_masks + c0, 0, n * sizeof(vs_masks), cs));
        if (e1 > e0) {
            vs_mask_entry *dst = ctx->d_sparse + staged;
            CK(cudaMemcpyAsync(dst, t->sparse + e0, (e1 - e0) * sizeof(vs_mask_entry), cudaMemcpyHostToDevice, cs));
            k_scatter_masks<<<(unsigned)((e1 - e0 + 255) / 256), 256, 0, cs>>>(dst, e1 - e0, g0, ctx->d_masks + c0);
            launches++;
            staged += e1 - e0;
            bytes += (e1 - e0) * sizeof(vs_mask_entry);
        }
    } else {
        CK(cudaMemcpyAsync(ctx->d_masks + c0, t->masks + g0, n * sizeof(vs_masks), cudaMemcpyHostToDevice, cs));
        bytes += n * sizeof(vs_masks);
    }
    return VS_OK;
}

// device staging for the sparse mask entries, or — for a view with a compact mask source — the plane buffers and the
// staging area of their runs (a run that straddles a chunk border is staged once per chunk it touches)
static int ensure_sparse_staging(vs_ctx *ctx, const vs_text_view *t, uint64_t first_word, uint64_t n_words)
{
    if (has_mask_source(t)) {
        if (n_words + 1 > ctx->planes_cap) {
            CK(cudaStreamSynchronize(ctx->copy));
            cudaFree(ctx->d_nm); cudaFree(ctx->d_em); cudaFree(ctx->d_emcode);
            ctx->d_nm = ctx->d_em = nullptr; ctx->d_emcode = nullptr; ctx->planes_cap = 0;
            CK(cudaMalloc(&ctx->d_nm, (n_words + 1) * sizeof(uint32_t)));
            CK(cudaMalloc(&ctx->d_em, (n_words + 1) * sizeof(uint32_t)));
            CK(cudaMalloc(&ctx->d_emcode, n_words + 1));
            ctx->planes_cap = n_words + 1;
        }
        const uint64_t n_blocks = (t->n_words + VS_EM_BLOCK) / VS_EM_BLOCK;
        if (n_blocks > ctx->dense_cap) {
            CK(cudaStreamSynchronize(ctx->copy));
            CK(cudaStreamSynchronize(ctx->prep));
            cudaFree(ctx->d_dense); ctx->d_dense = nullptr; ctx->dense_cap = 0;
            CK(cudaMalloc(&ctx->d_dense, n_blocks));
            ctx->dense_cap = n_blocks;
        }
        uint64_t a0, a1, b0, b1;
        runs_in_range(t->nm_runs, t->n_nm_runs, first_word, first_word + n_words + 1, a0, a1);
        runs_in_range(t->em_runs, t->n_em_runs, first_word, first_word + n_words + 1, b0, b1);
        const uint64_t n_chunks = (n_words + ctx->chunk_words - 1) / ctx->chunk_words + 3;     // + the tapered tail of chunk_plan()
        const uint64_t need = (a1 - a0) + (b1 - b0) + 4 * n_chunks + 16;
        if (need > ctx->runs_cap) {
            CK(cudaStreamSynchronize(ctx->copy));
            cudaFree(ctx->d_runs); ctx->d_runs = nullptr; ctx->runs_cap = 0;
            CK(cudaMalloc(&ctx->d_runs, need * sizeof(vs_plane_run)));
            ctx->runs_cap = need;
        }
        return VS_OK;
    }
    if (!t->sparse || t->n_sparse == 0) return VS_OK;
    const vs_mask_entry *sb = t->sparse, *se = t->sparse + t->n_sparse;
    auto lb = [&](uint64_t w) { return (uint64_t)(std::lower_bound(sb, se, w, [](const vs_mask_entry &x, uint64_t v) { return x.word < v; }) - sb); };
    uint64_t need = lb(first_word + n_words) - lb(first_word);
    if (need > ctx->sparse_cap) {
        CK(cudaStreamSynchronize(ctx->copy));
        cudaFree(ctx->d_sparse); ctx->d_sparse = nullptr; ctx->sparse_cap = 0;
        CK(cudaMalloc(&ctx->d_sparse, need * sizeof(vs_mask_entry)));
        ctx->sparse_cap = need;
    }
    return VS_OK;
}

// Start of an upload of a view with a compact mask source: zero both planes of the shard on the preparation stream
// (after `after`, an event of the caller's stream, if given) and make the copy stream wait for it — the dense blocks of
// the coded-block flags of the contig-end plane go to the device once per upload.
static int begin_mask_source(vs_ctx *ctx, const vs_text_view *t, uint64_t n_words, cudaEvent_t after)
{
    if (!has_mask_source(t)) return VS_OK;
    if (after) CK(cudaStreamWaitEvent(ctx->prep, after, 0));
    CK(cudaMemsetAsync(ctx->d_nm, 0, (n_words + 1) * sizeof(uint32_t), ctx->prep));
    CK(cudaMemsetAsync(ctx->d_em, 0, (n_words + 1) * sizeof(uint32_t), ctx->prep));
    CK(cudaMemcpyAsync(ctx->d_dense, t->em_dense, (t->n_words + VS_EM_BLOCK) / VS_EM_BLOCK, cudaMemcpyHostToDevice, ctx->prep));
    CK(cudaEventRecord(ctx->ev_zeroed, ctx->prep));
    CK(cudaStreamWaitEvent(ctx->copy, ctx->ev_zeroed, 0));
    return VS_OK;
}

static int check_view(vs_ctx *ctx, const vs_text_view *t, uint64_t first_word, uint64_t n_words)
{
    if (!t || !t->bases || (!t->masks && t->n_words)) return fail(ctx, VS_ERR_ARG, "text view is incomplete");
    if (first_word + n_words > t->n_words) return fail(ctx, VS_ERR_ARG, "shard lies outside the text");
    if (t->n_words * 32 > (1ull << 32)) return fail(ctx, VS_ERR_ARG, "text exceeds 4 Gbases (32-bit positions, as common.h:9-19)");
    if ((t->em_code != nullptr) != (t->em_dense != nullptr) || (has_mask_source(t) && ((t->n_nm_runs && !t->nm_runs) || (t->n_em_runs && !t->em_runs))))
        return fail(ctx, VS_ERR_ARG, "text view carries an incomplete compact mask source");
    return VS_OK;
}

extern "C" int vs_text_upload(vs_ctx *ctx, const vs_text_view *t, uint64_t first_word, uint64_t n_words)
{
    if (!ctx) return fail(nullptr, VS_ERR_ARG, "vs_text_upload: ctx is NULL");
    int r = check_view(ctx, t, first_word, n_words);
    if (r != VS_OK) return r;
    CK(cudaSetDevice(ctx->device));
    if ((r = ensure_text_buffers(ctx, n_words)) != VS_OK) return r;
    if ((r = ensure_sparse_staging(ctx, t, first_word, n_words)) != VS_OK) return r;
    uint64_t sparse_used = 0, bytes = 0;
    uint32_t launches = 0;
    cudaStream_t ready;
    if ((r = begin_mask_source(ctx, t, n_words, nullptr)) != VS_OK) return r;
    for (uint64_t c0 = 0; c0 < n_words; c0 += ctx->chunk_words) {
        uint64_t c1 = std::min(n_words, c0 + ctx->chunk_words);
        if ((r = enqueue_chunk_copy(ctx, t, first_word, c0, c1, sparse_used, bytes, launches, ready)) != VS_OK) return r;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->copy));
    CK(cudaStreamSynchronize(ctx->prep));
    ctx->n_words = n_words;
    ctx->first_word = first_word;
    ctx->err.clear();
    return VS_OK;
}

// ---- the scan -----------------------------------------------------------------------------------
// One pass = for every chunk of the resident shard: [H2D on the copy stream when `src` is given] -> k_extract ->
// k_score per strand and guide chunk, all enqueued without host synchronisation; counters are read back once
// at the end.  Chunks whose candidate store overflowed are redone afterwards; a hit-buffer overflow repeats the
// pass (from the now resident text) with a larger buffer.
static int scan_core(vs_ctx *ctx, const vs_text_view *src, uint64_t first_word, uint64_t n_words,
                     const uint8_t *guides, uint32_t n_guides, int k, int extra_pam,
                     vs_hit *out, uint64_t out_cap, uint64_t *n_hits, vs_scan_stats *stats)
{
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
    int r;
    if (src) {
        if ((r = ensure_text_buffers(ctx, n_words)) != VS_OK) return r;
        if ((r = ensure_sparse_staging(ctx, src, first_word, n_words)) != VS_OK) return r;
        ctx->n_words = n_words;
        ctx->first_word = first_word;
    }
    n_words = ctx->n_words;
    if (n_words == 0 || n_guides == 0) {
        if (src && n_words) { r = vs_text_upload(ctx, src, first_word, n_words); if (r != VS_OK) return r; }
        if (stats) *stats = S;
        return VS_OK;
    }
    PamParams pp;
    make_pam(extra_pam, pp);
    const uint64_t chunk_words = ctx->chunk_words;
    const std::vector<uint64_t> plan = chunk_plan(n_words, chunk_words, src != nullptr);
    const uint32_t n_chunks = (uint32_t)(plan.size() - 1);
    const uint32_t g_chunks = (n_guides + PAT_CHUNK - 1) / PAT_CHUNK;
    // tile size: about 60 blocks (both strands) per 64-thread CTA at the expected PAM density pp.n / 16 per strand
    uint32_t tile_words = (uint32_t)(60.0 * 16.0 / (2.0 * pp.n)) & ~7u;
    if (const char *e = getenv("VARSCOT_TILE_WORDS")) tile_words = (uint32_t)atoi(e);      // tuning knob
    if (tile_words > (uint32_t)EX_MAX_WORDS) tile_words = EX_MAX_WORDS;
    if (tile_words < 8) tile_words = 8;
    S.n_chunks = n_chunks;

    // pattern tables per guide chunk: [strand][PAT_CHUNK][PAT_STRIDE] (layout: see c_pat in vs_kernels.cuh), staged in pinned memory
    const uint64_t pat_chunk_words = (uint64_t)PAT_TABLE_WORDS;
    const uint64_t pat_words = (uint64_t)g_chunks * pat_chunk_words;
    if (pat_words > ctx->pat_cap) {
        CK(cudaStreamSynchronize(st));
        cudaFree(ctx->d_pat); if (ctx->h_pat) cudaFreeHost(ctx->h_pat);
        ctx->d_pat = ctx->h_pat = nullptr; ctx->pat_cap = 0;
        CK(cudaMalloc(&ctx->d_pat, pat_words * sizeof(uint32_t)));
        CK(cudaMallocHost(&ctx->h_pat, pat_words * sizeof(uint32_t)));
        ctx->pat_cap = pat_words;
    }
    memset(ctx->h_pat, 0, pat_words * sizeof(uint32_t));
    for (int s = 0; s < 2; ++s)
        for (uint32_t g = 0; g < n_guides; ++g) {
            uint32_t *dst = ctx->h_pat + (size_t)(g / PAT_CHUNK) * pat_chunk_words + ((size_t)s * PAT_CHUNK + g % PAT_CHUNK) * PAT_STRIDE;
            const uint8_t *gd = guides + (size_t)g * VS_GLEN;
            // slot order: informative positions first, the PAM dinucleotide last (forward 0..22; reverse 2..22, 0, 1)
            for (int j = 0; j < VS_GLEN; ++j) {
                const int i = slot_position(s, j);
                const int b = s ? 3 - gd[VS_GLEN - 1 - i] : gd[i];      // reverse pass scores revcomp(guide), bidir_mapping.cpp:293
                dst[j] = pat_slot(k, s, j, b);
            }
        }
    // counters
    if (n_chunks > ctx->cnt_chunks) {
        CK(cudaStreamSynchronize(st));
        cudaFree(ctx->d_cnt); if (ctx->h_cnt) cudaFreeHost(ctx->h_cnt);
        ctx->d_cnt = ctx->h_cnt = nullptr; ctx->cnt_chunks = 0;
        uint64_t nc = n_chunks + 8;
        CK(cudaMalloc(&ctx->d_cnt, (nc * 4 + 4) * sizeof(unsigned long long)));
        CK(cudaMallocHost(&ctx->h_cnt, (nc * 4 + 4) * sizeof(unsigned long long)));
        ctx->cnt_chunks = nc;
    }
    unsigned long long *d_hitcnt = ctx->d_cnt + ctx->cnt_chunks * 4, *h_hitcnt = ctx->h_cnt + ctx->cnt_chunks * 4;
    // candidate stores sized for one chunk at the expected density (+15 %), regrown when a chunk overflows
    auto ensure_blocks = [&](uint64_t need) -> int {
        if (need <= ctx->blocks_cap) return VS_OK;
        CK(cudaStreamSynchronize(st));
        CK(cudaStreamSynchronize(ctx->exs));
        for (int b = 0; b < 2; ++b)
            for (int s = 0; s < 2; ++s) {
                cudaFree(ctx->d_planes[b][s]); cudaFree(ctx->d_pos[b][s]);
                ctx->d_planes[b][s] = ctx->d_pos[b][s] = nullptr;
            }
        ctx->blocks_cap = 0;
        for (int b = 0; b < 2; ++b)
            for (int s = 0; s < 2; ++s) {
                CK(cudaMalloc(&ctx->d_planes[b][s], need * BLK_WORDS * sizeof(uint32_t)));
                CK(cudaMalloc(&ctx->d_pos[b][s], need * 32 * sizeof(uint32_t)));
            }
        ctx->blocks_cap = need;
        return VS_OK;
    };
    {
        const uint64_t cw = std::min(chunk_words, n_words);
        const uint64_t tiles = (cw + tile_words - 1) / tile_words;
        uint64_t est = (uint64_t)((double)cw * pp.n / 16.0 * 1.15) + tiles + 256;
        est = (est + SCORE_THREADS - 1) / SCORE_THREADS * SCORE_THREADS;
        if ((r = ensure_blocks(est)) != VS_OK) return r;
    }
    if (!ctx->d_hits) {
        uint64_t cap = 4u << 20;
        CK(cudaMalloc(&ctx->d_hits, cap * sizeof(vs_hit)));
        ctx->hits_cap = cap;
    }
    constexpr int EVC = 5;      // events per chunk
    while (ctx->ev_pool.size() < (size_t)EVC * n_chunks) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        ctx->ev_pool.push_back(e);
    }

    // extraction of chunk c into candidate buffer `buf` on stream `es`, scoring on the main stream
    auto launch_extract = [&](uint32_t c, int buf, cudaStream_t es) -> int {
        const uint64_t c0 = plan[c], c1 = plan[c + 1];
        const unsigned tiles = (unsigned)((c1 - c0 + tile_words - 1) / tile_words);
        k_extract<<<tiles, EX_THREADS, 0, es>>>(ctx->d_bases, ctx->d_masks, c0, c1, tile_words, ctx->first_word * 32, pp,
                                                 ctx->d_planes[buf][0], ctx->d_pos[buf][0], ctx->d_planes[buf][1], ctx->d_pos[buf][1],
                                                 ctx->blocks_cap, ctx->d_cnt + (uint64_t)c * 4);
        S.launches++;
        return VS_OK;
    };
    auto launch_score = [&](uint32_t c, int buf) -> int {
        unsigned long long *cnt = ctx->d_cnt + (uint64_t)c * 4;
        for (uint32_t gc = 0; gc < g_chunks; ++gc) {
            const uint32_t np = std::min<uint32_t>(PAT_CHUNK, n_guides - gc * PAT_CHUNK);
            CK(cudaMemcpyToSymbolAsync(c_pat, ctx->d_pat + (size_t)gc * pat_chunk_words, pat_chunk_words * sizeof(uint32_t), 0,
                                       cudaMemcpyDeviceToDevice, st));
            ScoreArgs a;
            for (int s = 0; s < 2; ++s) { a.planes[s] = ctx->d_planes[buf][s]; a.pos[s] = ctx->d_pos[buf][s]; }
            a.n_blocks_ptr = cnt + 2; a.cap = ctx->blocks_cap;
            a.ctas_per_strand = (uint32_t)((ctx->blocks_cap + SCORE_THREADS - 1) / SCORE_THREADS);
            a.n_pat = np; a.guide_base = gc * PAT_CHUNK;
            a.pat_global = ctx->d_pat + (size_t)gc * pat_chunk_words;
            a.hits = ctx->d_hits; a.n_hits = d_hitcnt; a.hit_cap = ctx->hits_cap;
            dispatch_score(k, a, st);
            S.launches++; S.score_launches++;
        }
        return VS_OK;
    };

    // Optional: run the extraction of chunk c+1 on its own (low-priority) stream, concurrently with the scoring of chunk c.
    // Measured on B200 (config 3) with 3, 4 and 5 resident k_score CTAs per SM and room left for k_extract CTAs: no gain
    // (the step takes the sum of the two kernels either way), so it is off by default, which also keeps the per-phase
    // event times additive.
    const bool overlap_extract = getenv("VARSCOT_OVERLAP_EXTRACT") && atoi(getenv("VARSCOT_OVERLAP_EXTRACT")) != 0;
    cudaStream_t es = overlap_extract ? ctx->exs : st;
    uint64_t found = 0;
    const vs_text_view *source = src;
    for (int attempt = 0; attempt < 3; ++attempt) {
        S.launches = 0; S.score_launches = 0; S.h2d_bytes = 0; S.redo_chunks = 0;
        CK(cudaEventRecord(ctx->ev[0], st));
        CK(cudaMemcpyAsync(ctx->d_pat, ctx->h_pat, pat_words * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        S.h2d_bytes += pat_words * sizeof(uint32_t);
        CK(cudaMemsetAsync(ctx->d_cnt, 0, (ctx->cnt_chunks * 4 + 4) * sizeof(unsigned long long), st));
        CK(cudaEventRecord(ctx->ev[2], st));
        CK(cudaStreamWaitEvent(es, ctx->ev[2], 0));                  // counters are zeroed before any extraction
        if (source) {
            CK(cudaStreamWaitEvent(ctx->copy, ctx->ev[0], 0));
            if ((r = begin_mask_source(ctx, source, n_words, ctx->ev[0])) != VS_OK) return r;
        }
        uint64_t sparse_used = 0;
        for (uint32_t c = 0; c < n_chunks; ++c) {
            cudaEvent_t *E = &ctx->ev_pool[(size_t)EVC * c];
            const int buf = (int)(c & 1);
            if (source) {
                const uint64_t c0 = plan[c], c1 = plan[c + 1];
                cudaStream_t ready;
                if ((r = enqueue_chunk_copy(ctx, source, first_word, c0, c1, sparse_used, S.h2d_bytes, S.launches, ready)) != VS_OK) return r;
                CK(cudaEventRecord(E[0], ready));
                CK(cudaStreamWaitEvent(es, E[0], 0));
            } else {
                CK(cudaEventRecord(E[0], es));
            }
            if (c >= 2) CK(cudaStreamWaitEvent(es, ctx->ev_pool[(size_t)EVC * (c - 2) + 4], 0));   // its buffer was scored
            CK(cudaEventRecord(E[1], es));
            if ((r = launch_extract(c, buf, es)) != VS_OK) return r;
            CK(cudaEventRecord(E[2], es));
            CK(cudaStreamWaitEvent(st, E[2], 0));
            CK(cudaEventRecord(E[3], st));
            if ((r = launch_score(c, buf)) != VS_OK) return r;
            CK(cudaEventRecord(E[4], st));
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[1], st));
        CK(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, (ctx->cnt_chunks * 4 + 4) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        source = nullptr;                                   // the text is resident from here on
        // chunks whose candidate store overflowed were skipped by k_score: redo them with a larger store
        S.n_cand_fwd = S.n_cand_rev = S.n_blocks_fwd = S.n_blocks_rev = 0;
        const uint64_t cap_in_pass = ctx->blocks_cap;        // what k_score compared against during the pass
        for (uint32_t c = 0; c < n_chunks; ++c) {
            unsigned long long *hc = ctx->h_cnt + (uint64_t)c * 4;
            if (hc[2] > cap_in_pass || hc[3] > cap_in_pass) {
                uint64_t need = std::max(hc[2], hc[3]);
                need = (need + need / 32 + SCORE_THREADS) / SCORE_THREADS * SCORE_THREADS;
                if ((r = ensure_blocks(need)) != VS_OK) return r;
                CK(cudaMemsetAsync(ctx->d_cnt + (uint64_t)c * 4, 0, 4 * sizeof(unsigned long long), st));
                if ((r = launch_extract(c, 0, st)) != VS_OK) return r;
                if ((r = launch_score(c, 0)) != VS_OK) return r;
                CK(cudaMemcpyAsync(hc, ctx->d_cnt + (uint64_t)c * 4, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(h_hitcnt, d_hitcnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                if (hc[2] > ctx->blocks_cap || hc[3] > ctx->blocks_cap) return fail(ctx, VS_ERR_CUDA, "vs_scan: candidate store overflow after regrow");
                S.redo_chunks++;
            }
            S.n_cand_fwd += hc[0]; S.n_cand_rev += hc[1]; S.n_blocks_fwd += hc[2]; S.n_blocks_rev += hc[3];
        }
        found = *h_hitcnt;
        if (found <= ctx->hits_cap) break;
        if (attempt == 2) return fail(ctx, VS_ERR_CUDA, "vs_scan: hit buffer overflow after regrow");
        // device hit buffer too small: grow and repeat the pass
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
    S.d2h_bytes = ncopy * sizeof(vs_hit) + (ctx->cnt_chunks * 4 + 4) * sizeof(unsigned long long);
    // per-phase device times: sums over the chunks of the last pass (same stream, so the intervals do not overlap)
    // (extraction of chunk c+1 runs concurrently with the scoring of chunk c, so the two sums overlap in time)
    for (uint32_t c = 0; c < n_chunks; ++c) {
        float a = 0.f, b = 0.f;
        CK(cudaEventElapsedTime(&a, ctx->ev_pool[(size_t)EVC * c + 1], ctx->ev_pool[(size_t)EVC * c + 2]));
        CK(cudaEventElapsedTime(&b, ctx->ev_pool[(size_t)EVC * c + 3], ctx->ev_pool[(size_t)EVC * c + 4]));
        S.extract_ms += a; S.score_ms += b;
    }
    CK(cudaEventElapsedTime(&S.total_ms, ctx->ev[0], ctx->ev[5]));
    if (src) CK(cudaEventElapsedTime(&S.upload_ms, ctx->ev[0], ctx->ev_pool[(size_t)EVC * (n_chunks - 1)]));
    if (stats) *stats = S;
    ctx->err.clear();
    if (found > out_cap) return fail(ctx, VS_ERR_OVERFLOW, "vs_scan: caller hit buffer too small; use vs_scan_fetch");
    return VS_OK;
}

extern "C" int vs_scan(vs_ctx *ctx, const uint8_t *guides, uint32_t n_guides, int k, int extra_pam,
                       vs_hit *out, uint64_t out_cap, uint64_t *n_hits, vs_scan_stats *stats)
{
    if (!ctx) return fail(nullptr, VS_ERR_ARG, "vs_scan: ctx is NULL");
    return scan_core(ctx, nullptr, ctx->first_word, ctx->n_words, guides, n_guides, k, extra_pam, out, out_cap, n_hits, stats);
}

extern "C" int vs_scan_text(vs_ctx *ctx, const vs_text_view *text, uint64_t first_word, uint64_t n_words,
                            const uint8_t *guides, uint32_t n_guides, int k, int extra_pam,
                            vs_hit *out, uint64_t out_cap, uint64_t *n_hits, vs_scan_stats *stats)
{
    if (!ctx) return fail(nullptr, VS_ERR_ARG, "vs_scan_text: ctx is NULL");
    int r = check_view(ctx, text, first_word, n_words);
    if (r != VS_OK) return r;
    return scan_core(ctx, text, first_word, n_words, guides, n_guides, k, extra_pam, out, out_cap, n_hits, stats);
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
