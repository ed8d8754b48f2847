// tests/cpu_kernel_units.cpp — TEST INFRASTRUCTURE: the pure helper functions of varscot_b200/csrc/vs_kernels.cuh
// (candidate masks, register transposes, bit-sliced adders and thresholds, pattern-table encoding, plane layout) and the
// body of k_extract's phase 2 (vs_extract_block.inc, textually included by the kernel, and the experimental half-block
// variant) compiled for the HOST with g++ and checked against naive per-bit restatements.  The kernels themselves are not
// compiled here (VS_HOST_UNIT_TEST guards them out) and nothing in the product uses this file.
// Built and run by tests/test_host.py::test_kernel_helpers_on_the_host.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>

#define VS_HOST_UNIT_TEST
#define __host__
#define __device__
#define __forceinline__ inline
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
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }

#include "../varscot_b200/csrc/vs_kernels.cuh"

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
        uint32_t in16[16], h[16];
        for (int i = 0; i < 16; ++i) h[i] = in16[i] = r32();
        transpose16(h);
        for (int r = 0; r < 16; ++r)
            for (int c = 0; c < 16; ++c) {
                CHECK(((h[r] >> c) & 1) == ((in16[c] >> r) & 1));
                CHECK(((h[r] >> (16 + c)) & 1) == ((in16[c] >> (16 + r)) & 1));
            }
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
        CHECK(score_smem_planes(k) * SCORE_THREADS * 4 <= 48 * 1024);
        CHECK((score_smem_planes(k) * SCORE_THREADS * 4 + 1024) * score_min_blocks(k) <= 227 * 1024 + 1024 * score_min_blocks(k));
        for (int s = 0; s < 2; ++s) {
            std::set<int> pos;
            std::set<uint32_t> planes;
            for (int j = 0; j < VS_GLEN; ++j) {
                const int i = slot_position(s, j);
                CHECK(i >= 0 && i < VS_GLEN);
                pos.insert(i);
                if (j < pa) CHECK(i == score_pos_base(k, s) + j || (k >= 8));      // stage A covers a contiguous position range
                for (int b = 0; b < 4; ++b) {
                    const uint32_t e = pat_slot(k, s, j, b);
                    int di, db;
                    pat_decode(k, s, j, e, di, db);
                    CHECK(di == i && db == b);
                    const uint32_t off = j < pa ? e : (e & 0xFFFFu);
                    CHECK(off % (SCORE_THREADS * 4u) == 0);
                    const uint32_t plane = off / (SCORE_THREADS * 4u);
                    if (j < pa) {
                        CHECK(plane < (uint32_t)(4 * pa) && planes.insert(plane).second);
                    } else {
                        CHECK(plane >= (uint32_t)(4 * pa) && plane + 1 < (uint32_t)score_smem_planes(k) && (plane - 4 * pa) % 2 == 0);
                        CHECK((e >> 16) == (uint32_t)b);
                    }
                }
            }
            CHECK((int)pos.size() == VS_GLEN);             // the slot order is a permutation of the 23 positions
            CHECK((int)planes.size() == 4 * pa);
        }
    }
    // the PAM dinucleotide is scored last on both strands
    CHECK(slot_position(0, 21) == 21 && slot_position(0, 22) == 22 && slot_position(1, 21) == 0 && slot_position(1, 22) == 1);
}

static void test_mismatch_plane()
{
    for (int rep = 0; rep < 100; ++rep) {
        const uint32_t h = r32(), l = r32();
        for (uint32_t b = 0; b < 4; ++b) {
            const uint32_t m = mismatch_plane(h, l, b);
            for (int c = 0; c < 32; ++c) CHECK(((m >> c) & 1) == ((((h >> c) & 1) * 2 + ((l >> c) & 1)) != b));
        }
    }
}

// ---- phase 2 of k_extract on the host: the very text the kernel includes (vs_extract_block.inc, and the experimental
// vs_extract_half_block.inc) runs over a random tile and is compared with a naive gather of the tile's candidates.
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
    if (half) {
        for (uint32_t j2 = 0; j2 < 2 * (nbf + nbr); ++j2) {
#include "../varscot_b200/csrc/vs_extract_half_block.inc"
        }
    } else {
        for (uint32_t j = 0; j < nbf + nbr; ++j) {
#include "../varscot_b200/csrc/vs_extract_block.inc"
        }
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
        for (int half = 0; half < 2; ++half) {
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

int main()
{
    test_plane_index();
    test_cand_masks();
    test_transposes();
    check_le_k<0>(); check_le_k<1>(); check_le_k<2>(); check_le_k<3>(); check_le_k<4>();
    check_le_k<5>(); check_le_k<6>(); check_le_k<7>(); check_le_k<8>();
    test_pattern_table();
    test_mismatch_plane();
    test_extract_phase2();
    if (failures) { fprintf(stderr, "%d check(s) failed\n", failures); return 1; }
    printf("kernel helper units ok\n");
    return 0;
}
