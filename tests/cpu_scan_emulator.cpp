// tests/cpu_scan_emulator.cpp — TEST INFRASTRUCTURE: the device code of the scan (k_extract, k_score<K>, compiled from
// varscot_b200/csrc/vs_kernels.cuh under tests/cuda_on_host.h) run on the HOST over a packed text, chunk by chunk as
// scan_core (vs_device.cu) drives it on the GPU.  tests/test_host.py feeds it texts packed by the library and compares the
// hits, resolved by the library's host side, with the oracle: the kernels' logic end to end, without a GPU.  What it
// cannot show — ptxas, launch geometry, streams, the upload path — is what the tests marked gpu are for.
// Never linked into the product; far too slow to be a fallback (one OS thread per CUDA thread).
//
// usage: cpu_scan_emulator IN OUT
//   IN : u64 n_words, u32 n_guides, i32 k, i32 extra_pam, u32 tile_words (0 = scan_core's choice), u64 chunk_words,
//        bases[(n_words + 1)], masks[n_words], guides[n_guides][23]
//   OUT: u64 n_hits, hits[n_hits]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cuda_on_host.h"
#include "../varscot_b200/csrc/vs_kernels.cuh"

namespace vs { uint32_t sm[SC_SMEM_BYTES / 4] __attribute__((aligned(16))); }
using namespace vs;

template <int K>
static void run_score_bk(const BkScoreArgs &a, unsigned ctas) { launch_cta(ctas, SC_THREADS, [&] { k_score_bucketed<K>(a); }); }

static void dispatch_score_bk(int k, const BkScoreArgs &a, unsigned ctas)
{
    switch (k) {
    case 0: run_score_bk<0>(a, ctas); break; case 1: run_score_bk<1>(a, ctas); break; case 2: run_score_bk<2>(a, ctas); break;
    case 3: run_score_bk<3>(a, ctas); break; case 4: run_score_bk<4>(a, ctas); break; case 5: run_score_bk<5>(a, ctas); break;
    case 6: run_score_bk<6>(a, ctas); break; case 7: run_score_bk<7>(a, ctas); break; default: run_score_bk<8>(a, ctas); break;
    }
}

template <int K>
static void run_score(const ScoreArgs &a, unsigned ctas, unsigned threads) { launch_cta(ctas, threads, [&] { k_score<K>(a); }); }

static void dispatch_score(int k, const ScoreArgs &a, unsigned ctas, unsigned threads)
{
    switch (k) {
    case 0: run_score<0>(a, ctas, threads); break; case 1: run_score<1>(a, ctas, threads); break; case 2: run_score<2>(a, ctas, threads); break;
    case 3: run_score<3>(a, ctas, threads); break; case 4: run_score<4>(a, ctas, threads); break; case 5: run_score<5>(a, ctas, threads); break;
    case 6: run_score<6>(a, ctas, threads); break; case 7: run_score<7>(a, ctas, threads); break; default: run_score<8>(a, ctas, threads); break;
    }
}

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: cpu_scan_emulator IN OUT\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    uint64_t n_words = 0, chunk_words = 0;
    uint32_t n_guides = 0, tile_words = 0;
    int32_t k = 0, extra_pam = -1;
    bool ok = fread(&n_words, 8, 1, f) == 1 && fread(&n_guides, 4, 1, f) == 1 && fread(&k, 4, 1, f) == 1 && fread(&extra_pam, 4, 1, f) == 1 &&
              fread(&tile_words, 4, 1, f) == 1 && fread(&chunk_words, 8, 1, f) == 1;
    std::vector<vs_bases> bases(n_words + 1);
    std::vector<vs_masks> masks(n_words ? n_words : 1);
    std::vector<uint8_t> guides((size_t)n_guides * VS_GLEN + 1);
    ok = ok && fread(bases.data(), sizeof(vs_bases), n_words + 1, f) == n_words + 1 && fread(masks.data(), sizeof(vs_masks), n_words, f) == n_words &&
         fread(guides.data(), 1, (size_t)n_guides * VS_GLEN, f) == (size_t)n_guides * VS_GLEN;
    fclose(f);
    if (!ok || k < 0 || k > VS_MAX_MISMATCHES || chunk_words == 0) { fprintf(stderr, "bad input\n"); return 2; }

    // as scan_engine (vs_device.cu): PAM lists, tile size, the pattern table, one whole-shard candidate store; per pipeline
    // chunk k_extract -> k_extract_mark -> k_score of the chunk's block range; then ONE k_score over the whole store (what
    // a scan of the resident index runs), which must find the same hits
    PamParams pp;
    pp.n = 2;
    pp.fx[0] = 2; pp.fy[0] = 2; pp.fx[1] = 2; pp.fy[1] = 0; pp.fx[2] = 0; pp.fy[2] = 0;
    if (extra_pam >= 0) { pp.fx[2] = extra_pam / 4; pp.fy[2] = extra_pam % 4; pp.n = 3; }
    for (int j = 0; j < 3; ++j) { pp.rx[j] = 3 - pp.fy[j]; pp.ry[j] = 3 - pp.fx[j]; }
    if (tile_words == 0) tile_words = (uint32_t)(60.0 * 16.0 / (2.0 * pp.n)) & ~7u;
    tile_words = std::min<uint32_t>(std::max<uint32_t>(tile_words, 8), EX_MAX_WORDS);
    std::vector<uint16_t> pat((size_t)2 * std::max<uint32_t>(n_guides, 1) * PAT_STRIDE + 8, 0);
    uint16_t *pat16 = (uint16_t *)(((uintptr_t)pat.data() + 15) & ~(uintptr_t)15);         // rows are read with 16-byte loads
    for (int s = 0; s < 2; ++s)
        for (uint32_t g = 0; g < n_guides; ++g) {
            uint16_t *dst = pat16 + ((size_t)s * n_guides + g) * PAT_STRIDE;
            const uint8_t *gd = guides.data() + (size_t)g * VS_GLEN;
            for (int j = 0; j < VS_GLEN; ++j) {
                const int i = slot_position(s, j);
                dst[j] = pat_slot(s, j, s ? 3 - gd[VS_GLEN - 1 - i] : gd[i]);
            }
        }
    const uint64_t n_chunks = (n_words + chunk_words - 1) / chunk_words;
    uint64_t cap = n_words + (n_words + tile_words - 1) / tile_words + 64 * n_chunks + 256;   // generous: at most 32 candidates per word and strand
    cap = (cap + BLK_GROUP - 1) / BLK_GROUP * BLK_GROUP;
    std::vector<uint32_t> planes[2], pos[2];
    for (int s = 0; s < 2; ++s) { planes[s].assign(cap * BLK_WORDS, 0xDEADBEEFu); pos[s].assign(cap * 32, 0xDEADBEEFu); }
    std::vector<unsigned long long> cnt(4 + 4 * (n_chunks + 1) + 4, 0);
    unsigned long long *rng = cnt.data() + 4, *all = cnt.data() + 4 + 4 * (n_chunks + 1);
    std::vector<vs_hit> hits(1 << 16), hits2;
    unsigned long long n_hits = 0;
    // knobs of the emulation (environment): guides per pass of the resident scans (scan_engine's guide super-chunks: guide_base > 0,
    // a pattern table wider than the launch), persistent CTAs, rotation period of the warp roles
    const uint32_t guide_pass = getenv("VS_EMU_GUIDE_PASS") ? (uint32_t)atoi(getenv("VS_EMU_GUIDE_PASS")) : 0u;
    const unsigned ctas = getenv("VS_EMU_CTAS") ? (unsigned)std::max(1, atoi(getenv("VS_EMU_CTAS"))) : 3u;
    const uint32_t rot = getenv("VS_EMU_ROT") ? (uint32_t)atoi(getenv("VS_EMU_ROT")) : 1u;
    // CTA size as scan_engine's score_cta: one warp per 32 guides, a tail of <= 8 guides gets no warp of its own
    auto cta_threads = [](uint32_t ng) {
        if (ng >= (uint32_t)SC_THREADS) return (unsigned)SC_THREADS;
        const unsigned full = ng / 32, tail = ng % 32;
        return 32u * std::max(1u, full + ((tail > 8u || full == 0) ? 1u : 0u));
    };
    auto score = [&](const unsigned long long *r, std::vector<vs_hit> &out, unsigned long long &n, uint32_t g0 = 0, uint32_t ng = ~0u) {
        if (ng == ~0u) ng = n_guides;
        for (;;) {
            ScoreArgs a;
            for (int s = 0; s < 2; ++s) { a.planes[s] = planes[s].data(); a.pos[s] = pos[s].data(); }
            a.rng = r; a.cap = cap; a.n_guides = ng; a.guide_base = g0; a.pat_guides = n_guides; a.pat = pat16; a.rot_shift = rot;
            const unsigned long long before = n;
            a.hits = out.data(); a.n_hits = &n; a.hit_cap = out.size();
            dispatch_score(k, a, ctas, cta_threads(ng));       // persistent CTAs stride over the batches
            if (n <= out.size()) break;
            out.resize(n + n / 4);                             // hit buffer overflow: grow and redo this launch
            n = before;
        }
    };
    for (uint64_t c = 0; c < n_chunks && n_guides; ++c) {
        const uint64_t c0 = c * chunk_words, c1 = std::min(n_words, c0 + chunk_words);
        const unsigned tiles = (unsigned)((c1 - c0 + tile_words - 1) / tile_words);
        launch_cta(tiles, EX_THREADS, [&] {
            k_extract(bases.data(), masks.data(), c0, c1, tile_words, 0, pp, planes[0].data(), pos[0].data(), planes[1].data(), pos[1].data(), cap, cnt.data());
        });
        launch(1, 1, 32, [&] { k_extract_mark(cnt.data(), rng + 4 * c, all, planes[0].data(), planes[1].data(), cap); });
        if (cnt[2] > cap || cnt[3] > cap) { fprintf(stderr, "candidate store overflow\n"); return 3; }
        score(rng + 4 * c, hits, n_hits);
    }
    if (n_guides && n_words) {
        hits2.resize(hits.size());
        unsigned long long n2 = 0;
        if (guide_pass == 0) score(all, hits2, n2);
        else for (uint32_t g0 = 0; g0 < n_guides; g0 += guide_pass) score(all, hits2, n2, g0, std::min(guide_pass, n_guides - g0));
        auto key = [](const vs_hit &x) { return ((uint64_t)x.pos << 32) | x.info; };
        std::vector<uint64_t> a, b;
        for (unsigned long long i = 0; i < n_hits; ++i) a.push_back(key(hits[i]));
        for (unsigned long long i = 0; i < n2; ++i) b.push_back(key(hits2[i]));
        std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
        if (a != b) { fprintf(stderr, "the scan of the whole store (%llu hits) differs from the chunk-by-chunk scan (%llu hits)\n", n2, n_hits); return 4; }
        // ---- the bucketed index, built from the plain store and the text as build_bucket_index (vs_device.cu) does, then
        // scored by k_score_bucketed: the same hits once more
        std::vector<unsigned long long> ctl(2 * BK_N + 2 * BK_N + 2 * (BK_N + 1), 0);
        unsigned long long *hist = ctl.data(), *cursor = hist + 2 * BK_N, *start = cursor + 2 * BK_N;
        const uint64_t max_blocks = std::max(all[2], all[3]);
        const unsigned g = (unsigned)((max_blocks + BKB_THREADS - 1) / BKB_THREADS);
        launch_cta(g, BKB_THREADS, [&] { k_bucket_hist(planes[0].data(), planes[1].data(), all, pp, hist); }, 2);
        launch_cta(2, 1024, [&] { k_bucket_scan(hist, start, cursor); });
        std::vector<uint32_t> bpl[2], bps[2];
        uint64_t nbk[2];
        for (int s = 0; s < 2; ++s) {
            nbk[s] = start[s * (BK_N + 1) + BK_N];
            if (nbk[s] % SC_NB) { fprintf(stderr, "bucketed store is not a whole number of batches\n"); return 5; }
            bpl[s].assign((nbk[s] + SC_NB) * BLK_WORDS, 0xDEADBEEFu);
            bps[s].assign((nbk[s] + 1) * 32, BK_NOPOS);
        }
        launch_cta(g, BKB_THREADS, [&] {
            k_bucket_scatter(planes[0].data(), planes[1].data(), pos[0].data(), pos[1].data(), all, pp, cursor, bps[0].data(), bps[1].data(), nbk[0] * 32, nbk[1] * 32);
        }, 2);
        for (int s = 0; s < 2; ++s)
            if (nbk[s]) launch((unsigned)((nbk[s] + 63) / 64), 1, 64, [&] { k_bucket_gather(bases.data(), masks.data(), 0, bps[s].data(), nbk[s], bpl[s].data()); });
        // every candidate of the plain store sits in exactly one bucket slot
        for (int s = 0; s < 2; ++s) {
            uint64_t placed = 0;
            for (uint64_t i = 0; i < nbk[s] * 32; ++i) placed += bps[s][i] != BK_NOPOS;
            if (placed != cnt[s]) { fprintf(stderr, "strand %d: %llu candidates, %llu bucket slots filled\n", s, cnt[s], (unsigned long long)placed); return 5; }
        }
        std::vector<uint16_t> gkey_all((size_t)2 * n_guides);
        for (int s = 0; s < 2; ++s)
            for (uint32_t gi = 0; gi < n_guides; ++gi) {
                uint8_t pc[VS_GLEN];
                for (int i = 0; i < VS_GLEN; ++i) pc[i] = s ? (uint8_t)(3 - guides[(size_t)gi * VS_GLEN + VS_GLEN - 1 - i]) : guides[(size_t)gi * VS_GLEN + i];
                gkey_all[(size_t)s * n_guides + gi] = (uint16_t)key_of_codes(s, pc);
            }
        std::vector<vs_hit> hits3(hits.size());
        unsigned long long n3 = 0;
        const uint32_t pass = guide_pass ? guide_pass : n_guides;
        for (uint32_t g0 = 0; g0 < n_guides; g0 += pass) {
            // per guide pass, as scan_engine: the keys of the pass's guides [2][ng], their order per bucket, one launch
            const uint32_t ng = std::min(pass, n_guides - g0);
            std::vector<uint16_t> gkey((size_t)2 * ng), perm((size_t)2 * BK_N * ng, 0xFFFF);
            std::vector<uint32_t> cls((size_t)2 * BK_N * BK_CLS, 0);
            for (int s = 0; s < 2; ++s)
                for (uint32_t gi = 0; gi < ng; ++gi) gkey[(size_t)s * ng + gi] = gkey_all[(size_t)s * n_guides + g0 + gi];
            launch_cta(BK_N, 32, [&] { k_guide_classes(gkey.data(), ng, pp, perm.data(), cls.data()); }, 2);
            const unsigned long long before = n3;
            for (;;) {
                BkScoreArgs q;
                for (int s = 0; s < 2; ++s) { q.planes[s] = bpl[s].data(); q.pos[s] = bps[s].data(); }
                q.start = start; q.n_guides = ng; q.guide_base = g0; q.pat_guides = n_guides; q.pat = pat16;
                q.perm = perm.data(); q.cls = cls.data();
                n3 = before;
                q.hits = hits3.data(); q.n_hits = &n3; q.hit_cap = hits3.size();
                dispatch_score_bk(k, q, ctas);
                if (n3 <= hits3.size()) break;
                hits3.resize(n3 + n3 / 4);
            }
        }
        std::vector<uint64_t> c3;
        for (unsigned long long i = 0; i < n3; ++i) c3.push_back(key(hits3[i]));
        std::sort(c3.begin(), c3.end());
        if (a != c3) { fprintf(stderr, "the scan of the bucketed index (%llu hits) differs from the plain scan (%llu hits)\n", n3, n_hits); return 6; }
    }
    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 2; }
    const uint64_t n = n_hits;
    ok = fwrite(&n, 8, 1, f) == 1 && (n == 0 || fwrite(hits.data(), sizeof(vs_hit), n, f) == n);
    ok = fclose(f) == 0 && ok;
    return ok ? 0 : 2;
}
