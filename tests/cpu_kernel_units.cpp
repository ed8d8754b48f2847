// tests/cpu_kernel_units.cpp — TEST INFRASTRUCTURE: varscot_b200/csrc/vs_kernels.cuh compiled for the HOST with g++.
//  * the kernels whose threads never cooperate (k_fill_runs, k_expand_em_code, k_masks_from_planes, k_scatter_masks,
//    k_extract_mark, k_resolve_hits, k_pack_loc_hits) run thread by thread from their real source, against naive restatements;
//    k_score<K> and the contig-start kernels (barriers) run with one OS thread per CUDA thread;
//  * k_extract (warp shuffles, cp.async) is guarded out; its phase 2 runs through the very text the kernel includes
//    (vs_extract_block.inc), its phase 1 through cand_masks;
//  * the helpers (register transposes, bit-sliced adders and thresholds, pattern-table encoding, plane layout) have
//    their own checks.
// Nothing in the product uses this file.  Built and run by tests/test_host.py::test_device_code_on_the_host.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <algorithm>
#include <vector>

#include "cuda_on_host.h"
#include "../varscot_b200/csrc/vs_kernels.cuh"

namespace vs { uint32_t sm[SC_SMEM_BYTES / 4] __attribute__((aligned(16))); }      // k_score's dynamic shared memory (extern __shared__ in the kernel)
using namespace vs;

static int failures = 0;
#define CHECK(cond)                                                                     \
    do {                                                                                \
        if (!(cond)) { ++failures; fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

static std::mt19937_64 rng(12345);
static uint32_t r32() { return (uint32_t)rng(); }

static void test_plane_index()
{
    // word w of block b: a bijection onto [0, 48 * n_blocks) that puts the same word of 32 consecutive blocks side by side
    const uint64_t nb = 4096;
    std::vector<char> seen(nb * BLK_WORDS, 0);
    for (uint64_t b = 0; b < nb; ++b)
        for (int w = 0; w < BLK_WORDS; ++w) {
            const uint64_t i = plane_index(b, w);
            CHECK(i < nb * BLK_WORDS && !seen[i]);
            seen[i] = 1;
            if (b % 32 != 31) CHECK(plane_index(b + 1, w) == i + 1);
        }
}

static void test_cand_masks()
{
    // R2 on the genome: forward W[21..22] in the PAM list, reverse W[0..1] in its reverse complement; not on invalid starts
    for (int extra = -1; extra < 16; ++extra) {
        PamParams pp;
        pp.n = 2;
        pp.fx[0] = 2; pp.fy[0] = 2; pp.fx[1] = 2; pp.fy[1] = 0; pp.fx[2] = 0; pp.fy[2] = 0;
        if (extra >= 0) { pp.fx[2] = extra / 4; pp.fy[2] = extra % 4; pp.n = 3; }
        for (int j = 0; j < 3; ++j) { pp.rx[j] = 3 - pp.fy[j]; pp.ry[j] = 3 - pp.fx[j]; }
        for (int rep = 0; rep < 200; ++rep) {
            const vs_bases a{r32(), r32()}, b{r32(), r32()};
            const vs_masks m{r32() & r32(), 0};
            uint32_t fwd, rev;
            cand_masks(a, b, m, pp, fwd, rev);
            const uint64_t H = ((uint64_t)b.hi << 32) | a.hi, L = ((uint64_t)b.lo << 32) | a.lo;
            auto code = [&](int p) { return (int)(((H >> p) & 1) * 2 + ((L >> p) & 1)); };
            for (int s = 0; s < 32; ++s) {
                bool f = false, r = false;
                for (int j = 0; j < pp.n; ++j) {
                    f |= code(s + 21) == pp.fx[j] && code(s + 22) == pp.fy[j];
                    r |= code(s) == pp.rx[j] && code(s + 1) == pp.ry[j];
                }
                const bool valid = !((m.iv >> s) & 1);
                CHECK((((fwd >> s) & 1) != 0) == (f && valid));
                CHECK((((rev >> s) & 1) != 0) == (r && valid));
            }
        }
    }
}

static void test_transposes()
{
    for (int rep = 0; rep < 200; ++rep) {
        uint32_t in[32], a[32];
        for (int i = 0; i < 32; ++i) a[i] = in[i] = r32();
        transpose32(a);
        for (int i = 0; i < 32; ++i)
            for (int c = 0; c < 32; ++c) CHECK(((a[i] >> c) & 1) == ((in[c] >> i) & 1));
    }
}

template <int N>
static void check_popcount()
{
    for (int rep = 0; rep < 300; ++rep) {
        uint32_t in[N], init[5] = {0, 0, 0, 0, 0}, bit[5], bit2[5];
        // sparse and dense planes, so that all counts 0..N occur
        const uint32_t bias = rep % 3 == 0 ? r32() & r32() & r32() : rep % 3 == 1 ? r32() | r32() : r32();
        for (int i = 0; i < N; ++i) in[i] = rep % 3 == 2 ? r32() : (rep % 3 == 0 ? r32() & bias : r32() | bias);
        popcount_planes<N, false>(in, init, bit);
        // a second batch added onto the first count (what stage B does)
        uint32_t more[VS_GLEN - N > 0 ? VS_GLEN - N : 1];
        for (int i = 0; i < VS_GLEN - N; ++i) more[i] = r32() & r32();
        if constexpr (N < VS_GLEN) popcount_planes<VS_GLEN - N, true>(more, bit, bit2);
        for (int c = 0; c < 32; ++c) {
            int n = 0, n2 = 0;
            for (int i = 0; i < N; ++i) n += (in[i] >> c) & 1;
            for (int i = 0; i < VS_GLEN - N; ++i) n2 += (more[i] >> c) & 1;
            int got = 0, got2 = 0;
            for (int w = 0; w < 5; ++w) { got |= ((bit[w] >> c) & 1) << w; if (N < VS_GLEN) got2 |= ((bit2[w] >> c) & 1) << w; }
            CHECK(got == n);
            if (N < VS_GLEN) CHECK(got2 == n + n2);
        }
    }
}

template <int K>
static void check_le_k()
{
    // counts 0..23 in every lane position
    for (int n = 0; n <= VS_GLEN; ++n) {
        uint32_t b[5];
        for (int w = 0; w < 5; ++w) b[w] = ((n >> w) & 1) ? ~0u : 0u;
        CHECK(le_k<K>(b) == (n <= K ? ~0u : 0u));
    }
    check_popcount<stage_a_slots(K)>();
}

static void test_pattern_table()
{
    for (int k = 0; k <= VS_MAX_MISMATCHES; ++k) {
        const int pa = stage_a_slots(k);
        CHECK(pa > k && pa <= VS_GLEN);                     // an invalid lane (mismatch in all stage-A planes) can never pass
    }
    CHECK(SC_SMEM_BYTES <= 48 * 1024);
    for (int s = 0; s < 2; ++s) {
        std::set<int> pos;
        std::set<uint32_t> planes;
        for (int j = 0; j < VS_GLEN; ++j) {
            const int i = slot_position(s, j);
            CHECK(i >= 0 && i < VS_GLEN);
            pos.insert(i);
            for (int b = 0; b < 4; ++b) {
                const uint32_t e = pat_slot(s, j, b);
                int di, db;
                pat_decode(s, j, e, di, db);
                CHECK(di == i && db == b);
                CHECK(e % 4 == 0 && e / 4 == (uint32_t)(4 * i + b) && e / 4 < (uint32_t)SC_STRIDE);
                CHECK(planes.insert(e).second);
            }
        }
        CHECK((int)pos.size() == VS_GLEN);                 // the slot order is a permutation of the 23 positions
        CHECK((int)planes.size() == 4 * VS_GLEN);
    }
    // the PAM dinucleotide is scored last on both strands
    CHECK(slot_position(0, 21) == 21 && slot_position(0, 22) == 22 && slot_position(1, 21) == 0 && slot_position(1, 22) == 1);
    // shared-memory banks: the 4-word groups of up to 8 consecutive block rows fall into distinct banks
    for (int i = 0; i < VS_GLEN; ++i) {
        std::set<int> banks;
        for (int blk = 0; blk < 8; ++blk)
            for (int b = 0; b < 4; ++b) CHECK(banks.insert((blk * SC_STRIDE + 4 * i + b) % 32).second);
    }
}

// ---- phase 2 of k_extract on the host: the very text the kernel includes (vs_extract_block.inc, and the experimental
// runs over a random tile and is compared with a naive gather of the tile's candidates.
struct Tile {
    uint32_t nw;
    uint2 s_hl[EX_MAX_WORDS + 2], s_mk[EX_MAX_WORDS + 2];
    uint32_t s_m[2][EX_MAX_WORDS + 1], s_p[2][EX_MAX_WORDS + 1];
    uint32_t nf, nr, nbf, nbr;
};

static void make_tile(Tile &t, uint32_t nw, int density, const PamParams &pp)
{
    t.nw = nw;
    for (uint32_t i = 0; i < EX_MAX_WORDS + 2; ++i) {
        t.s_hl[i] = uint2{r32(), r32()};
        // invalid starts: none / sparse / long runs; last-window bits: sparse
        const uint32_t iv = density == 0 ? 0u : density == 1 ? (r32() & r32() & r32()) : ((i / 7) % 3 == 0 ? ~0u : 0u);
        t.s_mk[i] = uint2{iv, r32() & r32() & r32()};
    }
    uint32_t pf = 0, pr = 0;
    for (uint32_t i = 0; i < nw; ++i) {
        uint32_t f, r;
        cand_masks(vs_bases{t.s_hl[i].x, t.s_hl[i].y}, vs_bases{t.s_hl[i + 1].x, t.s_hl[i + 1].y}, vs_masks{t.s_mk[i].x, t.s_mk[i].y}, pp, f, r);
        t.s_m[0][i] = f; t.s_m[1][i] = r;
        t.s_p[0][i] = pf; t.s_p[1][i] = pr;
        pf += __popc(f); pr += __popc(r);
    }
    t.s_p[0][nw] = t.nf = pf; t.s_p[1][nw] = t.nr = pr;
    t.s_m[0][nw] = ~0u; t.s_m[1][nw] = ~0u;                 // the sentinel (k_extract, after phase 1)
    t.nbf = (pf + 31) >> 5; t.nbr = (pr + 31) >> 5;
}

static void run_phase2(const Tile &t, bool half, const unsigned long long base[2], uint64_t cap, uint32_t gbase,
                       uint32_t *planes_f, uint32_t *pos_f, uint32_t *planes_r, uint32_t *pos_r)
{
    const uint32_t nw = t.nw, nf = t.nf, nr = t.nr, nbf = t.nbf, nbr = t.nbr;
    const uint2 *s_hl = t.s_hl, *s_mk = t.s_mk;
    const uint32_t (*s_m)[EX_MAX_WORDS + 1] = t.s_m, (*s_p)[EX_MAX_WORDS + 1] = t.s_p;
    (void)half;
    for (uint32_t j = 0; j < nbf + nbr; ++j) {
#include "../varscot_b200/csrc/vs_extract_block.inc"
    }
}

static void test_extract_phase2()
{
    PamParams pp;
    pp.n = 3;
    pp.fx[0] = 2; pp.fy[0] = 2; pp.fx[1] = 2; pp.fy[1] = 0; pp.fx[2] = 0; pp.fy[2] = 2;
    for (int j = 0; j < 3; ++j) { pp.rx[j] = 3 - pp.fy[j]; pp.ry[j] = 3 - pp.fx[j]; }
    const unsigned long long base[2] = {37, 5};            // where the tile's blocks land in the two candidate stores
    const uint64_t n_store = 37 + 5 + 200;                  // blocks per store, a multiple of BLK_GROUP after rounding
    const uint64_t store_words = ((n_store + BLK_GROUP - 1) / BLK_GROUP) * BLK_GROUP * BLK_WORDS;
    static Tile t;
    for (int rep = 0; rep < 60; ++rep) {
        const uint32_t nw = rep % 5 == 0 ? 1 + r32() % 8 : 8 + r32() % (EX_MAX_WORDS - 8);
        make_tile(t, nw, rep % 3, pp);
        const uint32_t gbase = r32() & 0x0FFFFFFFu;
        for (int half = 0; half < 1; ++half) {
            std::vector<uint32_t> pl[2], ps[2];
            for (int s = 0; s < 2; ++s) { pl[s].assign(store_words, 0xDEADBEEFu); ps[s].assign(n_store * 32 + 32, 0xDEADBEEFu); }
            run_phase2(t, half != 0, base, n_store, gbase, pl[0].data(), ps[0].data(), pl[1].data(), ps[1].data());
            for (int s = 0; s < 2; ++s) {
                // naive gather: the candidates of the strand in text order
                uint32_t rank = 0;
                const uint32_t n = s ? t.nr : t.nf;
                for (uint32_t w = 0; w < nw; ++w)
                    for (int b = 0; b < 32; ++b) {
                        if (!((t.s_m[s][w] >> b) & 1)) continue;
                        const uint64_t blk = base[s] + rank / 32;
                        const int c = rank % 32;
                        const uint64_t H = ((uint64_t)t.s_hl[w + 1].x << 32) | t.s_hl[w].x, L = ((uint64_t)t.s_hl[w + 1].y << 32) | t.s_hl[w].y;
                        for (int i = 0; i < VS_GLEN; ++i) {
                            CHECK(((pl[s][plane_index(blk, i)] >> c) & 1) == ((H >> (b + i)) & 1));
                            CHECK(((pl[s][plane_index(blk, VS_GLEN + i)] >> c) & 1) == ((L >> (b + i)) & 1));
                        }
                        CHECK(((pl[s][plane_index(blk, BLK_LAST)] >> c) & 1) == ((t.s_mk[w].y >> b) & 1));
                        CHECK(((pl[s][plane_index(blk, BLK_VALID)] >> c) & 1) == 1);
                        CHECK(ps[s][blk * 32 + c] == gbase + w * 32 + (uint32_t)b);
                        ++rank;
                    }
                CHECK(rank == n);
                // the valid mask of the last block ends with the last candidate; blocks beyond stay untouched
                const uint32_t nb = (n + 31) / 32;
                if (n % 32) CHECK(pl[s][plane_index(base[s] + nb - 1, BLK_VALID)] == (1u << (n % 32)) - 1u);
                CHECK(pl[s][plane_index(base[s] + nb, BLK_VALID)] == 0xDEADBEEFu);
            }
        }
    }
}

// ---- the mask kernels (k_fill_runs, k_expand_em_code, k_masks_from_planes, k_scatter_masks) thread by thread on the host
static void test_mask_kernels()
{
    for (int rep = 0; rep < 20; ++rep) {
        const uint64_t n = 3000 + r32() % 9000;            // words of the range; planes hold n + 1 (halo)
        const uint64_t word_base = 4096ull * (r32() % 3) + r32() % 100;        // the shard starts somewhere inside the text
        std::vector<uint32_t> nm(n + 1, 0), em(n + 1, 0);
        // N plane: a few runs; contig-end plane: single bits (coded blocks), multi-bit words (runs) and empty stretches
        std::vector<vs_plane_run> nm_runs, em_runs;
        for (uint64_t w = 0; w <= n;) {
            if (r32() % 50 == 0) {
                const uint32_t cnt = 1 + r32() % 700, val = r32() % 3 ? ~0u : r32() | 1u;
                const uint64_t e = w + cnt < n + 1 ? w + cnt : n + 1;
                nm_runs.push_back(vs_plane_run{(uint32_t)(word_base + w), (uint32_t)(e - w), val});
                for (uint64_t i = w; i < e; ++i) nm[i] = val;
                w = e + 1;
            } else ++w;
        }
        const uint64_t blocks_at = word_base / VS_EM_BLOCK, n_blocks = (word_base + n) / VS_EM_BLOCK + 1;
        std::vector<uint8_t> dense(n_blocks, 0), code(n + 1, (uint8_t)VS_EM_NONE);
        for (uint64_t b = blocks_at; b < n_blocks; ++b) dense[b] = r32() % 2;
        for (uint64_t w = 0; w <= n; ++w) {
            const bool coded = dense[(word_base + w) / VS_EM_BLOCK];
            const uint32_t kind = r32() % 8;
            if (kind < 3) continue;
            uint32_t v = 1u << (r32() % 32);
            if (kind == 7) v |= 1u << (r32() % 32);
            em[w] = v;
            if (coded && !(v & (v - 1))) code[w] = (uint8_t)__builtin_ctz(v);
            else em_runs.push_back(vs_plane_run{(uint32_t)(word_base + w), 1u, v});
        }
        // device buffers of a shard whose word 0 is global word `word_base`
        std::vector<uint32_t> d_nm(n + 1, 0), d_em(n + 1, 0);
        const uint64_t lo = word_base, hi = word_base + n + 1;
        uint64_t longest = 1;
        for (const auto &x : nm_runs) longest = longest > x.count ? longest : x.count;
        launch((unsigned)((nm_runs.size() + 7) / 8), (unsigned)((longest + FILL_SEG - 1) / FILL_SEG), 256,
               [&] { k_fill_runs(nm_runs.data(), nm_runs.size(), lo, hi, word_base, d_nm.data()); });
        launch((unsigned)((n + 1 + 255) / 256), 1, 256, [&] { k_expand_em_code(code.data(), dense.data(), lo, n + 1, d_em.data()); });
        launch((unsigned)((em_runs.size() + 7) / 8), 1, 256, [&] { k_fill_runs(em_runs.data(), em_runs.size(), lo, hi, word_base, d_em.data()); });
        for (uint64_t w = 0; w <= n; ++w) { CHECK(d_nm[w] == nm[w]); CHECK(d_em[w] == em[w]); }
        std::vector<vs_masks> got(n), sparse_got(n, vs_masks{0, 0});
        launch((unsigned)((n + 255) / 256), 1, 256, [&] { k_masks_from_planes(d_nm.data(), d_em.data(), n, got.data()); });
        std::vector<vs_mask_entry> entries;
        for (uint64_t w = 0; w < n; ++w) {
            // the definition (include/varscot_scan.h): iv = N in [p, p+23) or contig end in [p, p+22); lw = end at p+22, valid
            const uint64_t N = ((uint64_t)nm[w + 1] << 32) | nm[w], E = ((uint64_t)em[w + 1] << 32) | em[w];
            uint32_t iv = 0, lw = 0;
            for (int p = 0; p < 32; ++p) {
                const bool bad = ((N >> p) & 0x7FFFFF) != 0 || ((E >> p) & 0x3FFFFF) != 0;
                if (bad) iv |= 1u << p;
                else if ((E >> (p + 22)) & 1) lw |= 1u << p;
            }
            CHECK(got[w].iv == iv && got[w].lw == lw);
            if (iv | lw) entries.push_back(vs_mask_entry{(uint32_t)(word_base + w), iv, lw});
        }
        launch((unsigned)((entries.size() + 255) / 256), 1, 256, [&] { k_scatter_masks(entries.data(), entries.size(), word_base, sparse_got.data()); });
        for (uint64_t w = 0; w < n; ++w) CHECK(sparse_got[w].iv == got[w].iv && sparse_got[w].lw == got[w].lw);
    }
}

// ---- k_score<K> on the host (one OS thread per CUDA thread): random candidate blocks (with planted near matches,
// last-window flags, a partial last block and padding blocks with an empty valid mask) against random guides, both
// strands, a guide count that exercises the 32 / 16 / 8 / 4-guide segments; the hits must be exactly those of a naive count.
template <int K>
static void check_k_score(uint32_t n_guides, unsigned threads = 0)
{
    const uint32_t nb[2] = {SC_NB + 7, 2 * SC_NB + 5};                         // blocks per strand (not multiples of the batch)
    const uint32_t lo[2] = {SC_NB, 0};                                         // the forward range starts inside the store
    const uint64_t cap = 4 * SC_NB;
    std::vector<uint8_t> guides(n_guides * VS_GLEN);
    for (auto &g : guides) g = r32() % 4;
    // candidates: [strand][block][lane] -> 23 codes, last-window flag, position; some lanes are near copies of a pattern
    std::vector<uint32_t> planes[2], pos[2];
    std::vector<uint8_t> cand[2];
    std::vector<uint32_t> valid_n[2];
    for (int s = 0; s < 2; ++s) {
        planes[s].assign(cap * BLK_WORDS, 0xA5A5A5A5u);
        pos[s].assign(cap * 32, 0);
        cand[s].assign((size_t)nb[s] * 32 * VS_GLEN, 0);
        valid_n[s].assign(nb[s], 32);
        valid_n[s][nb[s] - 1] = 1 + r32() % 31;                                // partial last block
        valid_n[s][nb[s] / 2] = 0;                                             // a padding block (k_extract_mark): garbage planes, empty valid mask
        for (uint32_t b = 0; b < nb[s]; ++b) {
            uint32_t w[BLK_WORDS] = {0};
            for (int c = 0; c < 32; ++c) {
                uint8_t *x = &cand[s][((size_t)b * 32 + c) * VS_GLEN];
                for (int i = 0; i < VS_GLEN; ++i) x[i] = r32() % 4;
                if (r32() % 4 == 0) {                                          // near copy: 0..K+2 mismatches
                    const uint8_t *g = &guides[(r32() % n_guides) * VS_GLEN];
                    for (int i = 0; i < VS_GLEN; ++i) x[i] = s ? 3 - g[VS_GLEN - 1 - i] : g[i];
                    for (int m = r32() % (K + 3); m > 0; --m) { const int i = r32() % VS_GLEN; x[i] = (x[i] + 1 + r32() % 3) % 4; }
                }
                for (int i = 0; i < VS_GLEN; ++i) {
                    w[i] |= (uint32_t)((x[i] >> 1) & 1) << c;
                    w[VS_GLEN + i] |= (uint32_t)(x[i] & 1) << c;
                }
                if (r32() % 3 == 0) w[BLK_LAST] |= 1u << c;
                if ((uint32_t)c < valid_n[s][b]) w[BLK_VALID] |= 1u << c;
                pos[s][(size_t)(lo[s] + b) * 32 + c] = (uint32_t)(s * 1000000 + b * 32 + c);
            }
            for (int i = 0; i < BLK_WORDS; ++i) planes[s][plane_index(lo[s] + b, i)] = w[i];
        }
    }
    // pattern table as scan_engine builds it; the launch scores rows [3, 3 + n_guides) of a larger table
    const uint32_t rows = n_guides + 5, g_base = 3;
    std::vector<uint16_t> store((size_t)2 * rows * PAT_STRIDE + 8, 0);
    uint16_t *pat = (uint16_t *)(((uintptr_t)store.data() + 15) & ~(uintptr_t)15);
    for (int s = 0; s < 2; ++s)
        for (uint32_t g = 0; g < n_guides; ++g)
            for (int j = 0; j < VS_GLEN; ++j) {
                const int i = slot_position(s, j);
                const int b = s ? 3 - guides[g * VS_GLEN + VS_GLEN - 1 - i] : guides[g * VS_GLEN + i];
                pat[((size_t)s * rows + g_base + g) * PAT_STRIDE + j] = pat_slot(s, j, b);
            }
    unsigned long long rng[4] = {lo[0], lo[1], lo[0] + nb[0], lo[1] + nb[1]}, n_hits = 0;
    std::vector<vs_hit> hits(1 << 20);
    ScoreArgs a;
    for (int s = 0; s < 2; ++s) { a.planes[s] = planes[s].data(); a.pos[s] = pos[s].data(); }
    a.rng = rng; a.cap = cap;
    a.n_guides = n_guides; a.guide_base = g_base; a.pat_guides = rows; a.pat = pat; a.rot_shift = n_guides % 2 ? 0 : 40;
    a.hits = hits.data(); a.n_hits = &n_hits; a.hit_cap = hits.size();
    launch_cta(2, threads ? threads : std::min<unsigned>(SC_THREADS, (n_guides + 31) / 32 * 32), [&] { k_score<K>(a); });
    std::set<std::pair<uint32_t, uint32_t>> got, want;
    CHECK(n_hits <= hits.size());
    for (unsigned long long i = 0; i < n_hits; ++i) CHECK(got.insert({hits[i].pos, hits[i].info}).second);
    for (int s = 0; s < 2; ++s)
        for (uint32_t b = 0; b < nb[s]; ++b)
            for (uint32_t c = 0; c < valid_n[s][b]; ++c)
                for (uint32_t g = 0; g < n_guides; ++g) {
                    const uint8_t *x = &cand[s][((size_t)b * 32 + c) * VS_GLEN], *gd = &guides[g * VS_GLEN];
                    int mm = 0, h2 = 0;
                    for (int i = 0; i < VS_GLEN; ++i) {
                        const int p = s ? 3 - gd[VS_GLEN - 1 - i] : gd[i];
                        mm += x[i] != p;
                        if (i >= 11) h2 += x[i] != p;
                    }
                    const bool last = (planes[s][plane_index(lo[s] + b, BLK_LAST)] >> c) & 1;
                    if (mm <= K && (!last || h2 <= K / 2))
                        want.insert({pos[s][(size_t)(lo[s] + b) * 32 + c], ((g_base + g) << 8) | ((uint32_t)s << 7) | (uint32_t)mm});
                }
    CHECK(got == want);
    CHECK(want.size() > 20);
}

// ---- k_extract_mark: closes a chunk's block range, pads the claims to a layout group, empties the padding blocks
static void test_extract_mark()
{
    const uint64_t cap = 8 * BLK_GROUP;
    std::vector<uint32_t> pl[2];
    for (int s = 0; s < 2; ++s) pl[s].assign(cap * BLK_WORDS, 0xDEADBEEFu);
    unsigned long long cnt[4] = {1000, 2000, 37, 64}, rng[8] = {0, 32, 0, 0, 0, 0, 0, 0}, all[4] = {9, 9, 9, 9};
    launch(1, 1, 32, [&] { k_extract_mark(cnt, rng, all, pl[0].data(), pl[1].data(), cap); });
    CHECK(rng[0] == 0 && rng[1] == 32 && rng[2] == 37 && rng[3] == 64);       // this chunk's range
    CHECK(rng[4] == 64 && rng[5] == 64 && cnt[2] == 64 && cnt[3] == 64);      // the next chunk starts on a group boundary
    CHECK(all[0] == 0 && all[1] == 0 && all[2] == 64 && all[3] == 64);
    for (uint64_t b = 0; b < cap; ++b) {
        const bool pad = b >= 37 && b < 64;
        CHECK(pl[0][plane_index(b, BLK_VALID)] == (pad ? 0u : 0xDEADBEEFu));
        CHECK(pl[1][plane_index(b, BLK_VALID)] == 0xDEADBEEFu);
    }
    // claims past the capacity: nothing is written out of bounds, the counters still advance
    unsigned long long cnt2[4] = {0, 0, cap + 5, 3}, rng2[8] = {0}, all2[4] = {0};
    launch(1, 1, 32, [&] { k_extract_mark(cnt2, rng2, all2, pl[0].data(), pl[1].data(), cap); });
    CHECK(cnt2[2] == cap + BLK_GROUP && rng2[2] == cap + 5);
}

// ---- contig starts from the contig-end plane, and the hit resolution that uses them
static void test_contig_starts_and_resolve()
{
    for (int rep = 0; rep < 6; ++rep) {
        const uint64_t n_words = rep == 0 ? 1 : 100 + r32() % (3 * CS_TILE);
        const uint64_t first_word = r32() % 1000;
        std::vector<uint32_t> em(n_words + 1, 0);
        std::vector<uint32_t> want;
        const uint32_t first_start = (uint32_t)(first_word * 32 - (rep % 2 ? 17 : 0));
        want.push_back(first_start);
        for (uint64_t w = 0; w < n_words; ++w) {
            const uint32_t kind = r32() % 6;
            em[w] = kind == 0 ? (1u << (r32() % 32)) | (1u << (r32() % 32)) : kind == 1 ? 1u << (r32() % 32) : kind == 2 && rep == 3 ? ~0u : 0u;
            for (int b = 0; b < 32; ++b) if ((em[w] >> b) & 1) want.push_back((uint32_t)((first_word + w) * 32 + b + 1));
        }
        em[n_words] = ~0u;                                   // the halo word does not count
        const unsigned tiles = (unsigned)((n_words + CS_TILE - 1) / CS_TILE);
        std::vector<uint32_t> tile(tiles + 1, 0xFFu), starts(want.size() + 4, 0xDEADBEEFu);
        uint32_t total = 0;
        launch_cta(tiles, 256, [&] { k_contig_starts_count(em.data(), n_words, tile.data()); });
        launch_cta(1, 1024, [&] { k_contig_starts_scan(tile.data(), tiles, &total); });
        launch_cta(tiles, 256, [&] { k_contig_starts_scatter(em.data(), n_words, tile.data(), first_word * 32, first_start, starts.data()); });
        CHECK(total + 1 == want.size());
        for (size_t i = 0; i < want.size(); ++i) CHECK(starts[i] == want[i]);
        CHECK(starts[want.size()] == 0xDEADBEEFu);
        // resolution: random hits inside the shard -> (contig, pos) by a linear search
        const uint64_t n = 500;
        std::vector<vs_hit> hits(n);
        std::vector<unsigned long long> keys(n), vals(n);
        const uint32_t first_contig = 70000 + r32() % 1000, guide_lo = 256;
        for (auto &h : hits) {
            h.pos = (uint32_t)(first_word * 32 + r32() % (n_words * 32));
            h.info = ((guide_lo + r32() % 300) << 8) | ((r32() & 1) << 7) | (r32() % 9);
        }
        launch((unsigned)((n + 255) / 256), 1, 256, [&] {
            k_resolve_hits(hits.data(), n, starts.data(), (uint32_t)want.size(), first_contig, guide_lo, keys.data(), vals.data());
        });
        std::vector<vs_loc_hit> loc(n);
        launch((unsigned)((n + 255) / 256), 1, 256, [&] { k_pack_loc_hits(keys.data(), vals.data(), n, loc.data()); });
        for (uint64_t i = 0; i < n; ++i) {
            size_t c = 0;
            while (c + 1 < want.size() && want[c + 1] <= hits[i].pos) ++c;
            const uint32_t contig = first_contig + (uint32_t)c, pos = hits[i].pos - want[c];
            CHECK(loc[i].contig == contig && loc[i].info == hits[i].info);
            CHECK((uint32_t)loc[i].key == pos && ((loc[i].key >> 32) & 0xFFFF) == (contig & 0xFFFF));
            CHECK(((loc[i].key >> 48) & 1) == ((hits[i].info >> 7) & 1) && (loc[i].key >> 49) == (hits[i].info >> 8) - guide_lo);
        }
    }
}

// CTA size as scan_engine's score_cta (vs_device.cu): one warp per 32 guides, a tail of <= 8 guides gets no warp of its own
static unsigned score_cta_threads(uint32_t ng)
{
    if (ng >= (uint32_t)SC_THREADS) return (unsigned)SC_THREADS;
    const unsigned full = ng / 32, tail = ng % 32;
    return 32u * std::max(1u, full + ((tail > 8u || full == 0) ? 1u : 0u));
}

// every guide count of a launch, with the CTA size the library would pick: no guide may go unscored (round 2's first k_score split
// a warp slice's tail into 16 + 8 + 4 guides and scored nothing for slices of 29..31)
static void sweep_guide_counts()
{
    for (uint32_t n = 1; n <= 200; ++n) check_k_score<2>(n, score_cta_threads(n));
    for (uint32_t n : {253u, 254u, 255u, 256u, 257u, 285u, 286u, 287u, 288u, 383u, 384u, 413u, 415u, 511u, 543u}) check_k_score<1>(n, score_cta_threads(n));
}

int main()
{
    test_plane_index();
    test_cand_masks();
    test_transposes();
    check_le_k<0>(); check_le_k<1>(); check_le_k<2>(); check_le_k<3>(); check_le_k<4>();
    check_le_k<5>(); check_le_k<6>(); check_le_k<7>(); check_le_k<8>();
    test_pattern_table();
    test_extract_phase2();
    test_mask_kernels();
    test_extract_mark();
    test_contig_starts_and_resolve();
    check_k_score<0>(9); check_k_score<1>(33); check_k_score<2>(5); check_k_score<3>(60); check_k_score<4>(128 + 28);
    check_k_score<5>(20); check_k_score<6>(100); check_k_score<7>(12); check_k_score<8>(4);
    check_k_score<6>(100, 96); check_k_score<4>(40, 32); check_k_score<3>(300, 64);      // CTAs with fewer warps than guide slices (a short tail gets no warp of its own)
    sweep_guide_counts();
    if (failures) { fprintf(stderr, "%d check(s) failed\n", failures); return 1; }
    printf("kernel helper units ok\n");
    return 0;
}
