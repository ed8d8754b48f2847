// varscot_b200/csrc/vs_kernels.cuh — hand-written sm_100a kernels of the off-target scan.
//
// What they replace: the inner loops of VARSCOT_pipeline/read_mapping/bidir_mapping.cpp —
// SeqAn's find<0,K>(delegate, index, half, HammingDistance()) (:129-146) plus the verify
// delegate (:34-127) — restated as a dense, PAM-first Hamming scan (rules R1-R4 of SURVEY.md 8a).
//
// Pipeline per scan (all on one stream):
//   k_extract  : per tile of 8192 window starts, find the PAM-valid / N-free / in-contig windows of each strand
//                (bit-parallel), compact them into blocks of 32 candidates and store each block bit-sliced
//                ACROSS candidates (per-thread 32x32 register transposes): 48 words per block =
//                hi_0..hi_22, lo_0..lo_22, last-window mask, valid mask; + 32 positions
//   k_score<K> : one thread per candidate block; expands the 46 planes into 92 "mismatch if guide base is b"
//                planes in shared memory, then for every guide: 23 LDS (plane selected by the guide base,
//                offset warp-uniform from constant memory) + 36 LOP3 carry-save adder + 2 LOP3 threshold.
//                Hits (rare) take a slow path: exact count, R4 check, atomic append.
// Integer pipe + shared-memory bound; no tensor cores (nothing here is a dense contraction worth a GEMM:
// the bit-sliced form costs ~1.2 ALU ops per (window, guide) pair, below one op per output element).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/varscot_scan.h"

namespace vs {

constexpr int TILE_WORDS   = 256;                 // words (of 32 window starts) per tile
constexpr int TILE_THREADS = 256;
constexpr int TILE_STARTS  = TILE_WORDS * 32;     // 8192
constexpr int BLK_WORDS    = 48;                  // words per candidate block
constexpr int BLK_LAST     = 46;                  // word index of the last-window mask
constexpr int BLK_VALID    = 47;                  // word index of the valid mask
constexpr int SCORE_THREADS = 256;
constexpr int NPLANES      = 92;                  // 23 positions x 4 guide bases
constexpr int PAT_STRIDE   = 24;                  // uint32 per pattern in constant memory (23 offsets + pad)
constexpr int PAT_CHUNK    = 512;                 // patterns per k_score launch (48 KB of constant memory)

struct PamParams {
    int n;            // number of forward dinucleotides (2 or 3)
    int fx[3], fy[3]; // forward: W[21] == fx && W[22] == fy
    int rx[3], ry[3]; // reverse: W[0]  == rx && W[1]  == ry   (reverse complement of the forward list)
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t eq_plane(uint32_t h, uint32_t l, int code)
{
    uint32_t mh = (code & 2) ? 0u : ~0u;   // XNOR masks: ~(h ^ H) = h ^ ~H
    uint32_t ml = (code & 1) ? 0u : ~0u;
    return (h ^ mh) & (l ^ ml);
}

// Candidate masks for the 32 window starts of one word (a = its bases, b = next word's bases, m = its masks).
//   m.iv (precomputed by the packer): start is invalid — the window holds an N, runs over a contig end or past the
//         text (R1, R3);  m.lw: the window ends exactly at a contig end (R4, bidir_mapping.cpp:51)
//   R2: PAM on the genome: forward W[21..22], reverse W[0..1]   (bidir_mapping.cpp:70-76, :240-247)
__device__ __forceinline__ void cand_masks(const vs_bases &a, const vs_bases &b, const vs_masks &m, const PamParams &pp,
                                           uint32_t &fwd, uint32_t &rev)
{
    const uint32_t h21 = __funnelshift_r(a.hi, b.hi, 21), h22 = __funnelshift_r(a.hi, b.hi, 22);
    const uint32_t l21 = __funnelshift_r(a.lo, b.lo, 21), l22 = __funnelshift_r(a.lo, b.lo, 22);
    const uint32_t h0 = a.hi, l0 = a.lo, h1 = __funnelshift_r(a.hi, b.hi, 1), l1 = __funnelshift_r(a.lo, b.lo, 1);
    uint32_t f = 0, r = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (j < pp.n) {
            f |= eq_plane(h21, l21, pp.fx[j]) & eq_plane(h22, l22, pp.fy[j]);
            r |= eq_plane(h0, l0, pp.rx[j]) & eq_plane(h1, l1, pp.ry[j]);
        }
    }
    fwd = f & ~m.iv;
    rev = r & ~m.iv;
}

// ------------------------------------------------------------------------------------------------
// 32 x 32 bit-matrix transpose in registers (LSB-first): out[i] bit c = in[c] bit i.
// Stages 16 and 8 move whole bytes (2 PRMT per pair), stages 4, 2, 1 are masked swaps (5 ops per pair);
// ptxas drops the swaps that only feed unused outputs.
__device__ __forceinline__ void transpose32(uint32_t (&a)[32])
{
#pragma unroll
    for (int k = 0; k < 16; ++k) {                       // 16-bit halves: a[k].hi <-> a[k+16].lo
        const uint32_t x = a[k], y = a[k + 16];
        a[k] = __byte_perm(x, y, 0x5410);
        a[k + 16] = __byte_perm(x, y, 0x7632);
    }
#pragma unroll
    for (int k = 0; k < 32; k = ((k | 8) + 1) & ~8) {    // bytes: odd bytes of a[k] <-> even bytes of a[k+8]
        const uint32_t x = a[k], y = a[k | 8];
        a[k] = __byte_perm(x, y, 0x6240);
        a[k | 8] = __byte_perm(x, y, 0x7351);
    }
    uint32_t m = 0x0F0F0F0Fu;
#pragma unroll
    for (int j = 4; j; j >>= 1, m ^= m << j) {
#pragma unroll
        for (int k = 0; k < 32; k = ((k | j) + 1) & ~j) {
            const uint32_t t = ((a[k] >> j) ^ a[k | j]) & m;
            a[k | j] ^= t;
            a[k] ^= t << j;
        }
    }
}

constexpr int EX_THREADS   = 64;                  // threads per extraction CTA
constexpr int EX_MAX_WORDS = 256;                 // words per tile (<=); the host picks the tile so that it holds ~60 blocks

// k_extract: one 64-thread CTA per tile of `tile_words` words.
//   phase 1 (bit-parallel, 32 starts per word): candidate masks per strand -> shared memory + exclusive rank prefix;
//   phase 2 (one THREAD per 32-candidate block): walk the masks from the block's first candidate, gather the 23-base
//            windows (funnel shifts on the staged planes), transpose 32 x 23 bits twice in registers, write
//            48 plane words + 32 positions.
// Block ranges are claimed with one atomicAdd per strand per tile on cnt[2], cnt[3] (layout order is not
// deterministic; hit resolution sorts).  If a claim runs past the capacity nothing is written for that tile; k_score
// then skips the whole chunk and the host, which reads the counters back, regrows the stores and redoes the chunk.
// Block layout (48 words): hi_0..hi_22, lo_0..lo_22, last-window mask, valid mask.
__global__ void __launch_bounds__(EX_THREADS)
k_extract(const vs_bases *__restrict__ B, const vs_masks *__restrict__ M, uint64_t w_begin, uint64_t w_end, uint32_t tile_words,
          uint64_t global_base, PamParams pp,
          uint32_t *__restrict__ planes_f, uint32_t *__restrict__ pos_f,
          uint32_t *__restrict__ planes_r, uint32_t *__restrict__ pos_r, uint64_t cap,
          unsigned long long *__restrict__ cnt)
{
    __shared__ uint2 s_hl[EX_MAX_WORDS + 2];
    __shared__ uint32_t s_m[2][EX_MAX_WORDS + 1];      // candidate masks per strand
    __shared__ uint32_t s_p[2][EX_MAX_WORDS + 1];      // exclusive rank prefix per word (+ total at [nw])
    __shared__ uint32_t s_lw[EX_MAX_WORDS + 1];
    __shared__ uint32_t wsum[2][EX_THREADS / 32];
    __shared__ unsigned long long base[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t w0 = w_begin + (uint64_t)blockIdx.x * tile_words;
    const uint32_t nw = (uint32_t)min((uint64_t)tile_words, w_end - w0);      // words of this tile
    uint32_t run_f = 0, run_r = 0;                     // candidates before the current group of 64 words
    for (uint32_t i0 = 0; i0 < nw; i0 += EX_THREADS) {
        const uint32_t i = i0 + tid;
        uint32_t fwd = 0, rev = 0;
        if (i < nw) {
            const vs_bases a = B[w0 + i], b = B[w0 + i + 1];    // B[w_end] is the halo / pad word, always readable
            const vs_masks m = M[w0 + i];
            s_hl[i] = make_uint2(a.hi, a.lo);
            if (i == nw - 1) s_hl[i + 1] = make_uint2(b.hi, b.lo);
            s_lw[i] = m.lw;
            cand_masks(a, b, m, pp, fwd, rev);
            s_m[0][i] = fwd; s_m[1][i] = rev;
        }
        const uint32_t cf = __popc(fwd), cr = __popc(rev);
        uint32_t xf = cf, xr = cr;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t p = __shfl_up_sync(0xffffffffu, xf, o), q = __shfl_up_sync(0xffffffffu, xr, o);
            if (lane >= o) { xf += p; xr += q; }
        }
        if (lane == 31) { wsum[0][wid] = xf; wsum[1][wid] = xr; }
        __syncthreads();
        const uint32_t w0f = wsum[0][0], w0r = wsum[1][0], w1f = wsum[0][1], w1r = wsum[1][1];
        if (i < nw) {
            s_p[0][i] = run_f + (wid ? w0f : 0u) + xf - cf;
            s_p[1][i] = run_r + (wid ? w0r : 0u) + xr - cr;
        }
        run_f += w0f + w1f; run_r += w0r + w1r;
        __syncthreads();
    }
    const uint32_t nf = run_f, nr = run_r;
    const uint32_t nbf = (nf + 31) >> 5, nbr = (nr + 31) >> 5;
    if (tid == 0) {
        s_p[0][nw] = nf; s_p[1][nw] = nr;
        s_m[0][nw] = 0; s_m[1][nw] = 0;
        base[0] = atomicAdd(&cnt[2], (unsigned long long)nbf);
        base[1] = atomicAdd(&cnt[3], (unsigned long long)nbr);
        if (nf) atomicAdd(&cnt[0], (unsigned long long)nf);
        if (nr) atomicAdd(&cnt[1], (unsigned long long)nr);
    }
    __syncthreads();
    const uint32_t gbase = (uint32_t)(global_base + w0 * 32);      // device word 0 sits at global_base
    for (uint32_t j = tid; j < nbf + nbr; j += EX_THREADS) {
        const int s = j >= nbf;
        const uint32_t jj = s ? j - nbf : j;
        const uint32_t n = s ? nr : nf;
        const uint64_t blk = base[s] + jj;
        if (blk >= cap) continue;                          // overflow: the chunk is redone by the host
        const uint32_t cntc = min(32u, n - jj * 32);
        const uint32_t *sm = s_m[s], *sp = s_p[s];
        // word holding candidate rank r0 = 32*jj: last word with prefix <= r0
        const uint32_t r0 = jj * 32;
        uint32_t lo_w = 0, hi_w = nw;                      // invariant: sp[lo_w] <= r0 < sp[hi_w] (sp[nw] = n > r0)
        while (hi_w - lo_w > 1) {
            const uint32_t mid = (lo_w + hi_w) >> 1;
            if (sp[mid] <= r0) lo_w = mid; else hi_w = mid;
        }
        uint32_t wcur = lo_w;
        uint32_t m = sm[wcur];
        for (uint32_t skip = r0 - sp[wcur]; skip; --skip) m &= m - 1;     // drop the candidates of earlier blocks
        uint2 ha = s_hl[wcur], hb = s_hl[wcur + 1];
        uint32_t lwm = s_lw[wcur];
        uint32_t ah[32], al[32];
        uint32_t lastw = 0;
        uint4 *ps_out = reinterpret_cast<uint4 *>((s ? pos_r : pos_f) + blk * 32);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            uint32_t pp4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c4 * 4 + u;
                uint32_t hw = 0, lw = 0, ps = 0xFFFFFFFFu;
                if ((uint32_t)c < cntc) {
                    while (m == 0) {                       // next word with candidates (the tile has >= cntc left)
                        ++wcur;
                        m = sm[wcur];
                        ha = hb; hb = s_hl[wcur + 1];
                        lwm = s_lw[wcur];
                    }
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    hw = __funnelshift_r(ha.x, hb.x, b) & 0x7FFFFFu;
                    lw = __funnelshift_r(ha.y, hb.y, b) & 0x7FFFFFu;
                    lastw |= ((lwm >> b) & 1u) << c;
                    ps = gbase + wcur * 32 + b;
                }
                ah[c] = hw; al[c] = lw; pp4[u] = ps;
            }
            ps_out[c4] = make_uint4(pp4[0], pp4[1], pp4[2], pp4[3]);
        }
        uint4 *pl_out = reinterpret_cast<uint4 *>((s ? planes_r : planes_f) + blk * BLK_WORDS);
        transpose32(ah);
#pragma unroll
        for (int i = 0; i < 5; ++i) pl_out[i] = make_uint4(ah[4 * i], ah[4 * i + 1], ah[4 * i + 2], ah[4 * i + 3]);
        transpose32(al);
        // words 20..47: hi_20, hi_21, hi_22, lo_0 .. lo_22, last, valid
        pl_out[5] = make_uint4(ah[20], ah[21], ah[22], al[0]);
#pragma unroll
        for (int i = 0; i < 5; ++i) pl_out[6 + i] = make_uint4(al[4 * i + 1], al[4 * i + 2], al[4 * i + 3], al[4 * i + 4]);
        const uint32_t validw = cntc >= 32 ? ~0u : ((1u << cntc) - 1u);
        pl_out[11] = make_uint4(al[21], al[22], lastw, validw);
    }
}

// k_scatter_masks: expand the sparse form of the window masks (only words with a non-zero mask travel over PCIe).
__global__ void __launch_bounds__(256)
k_scatter_masks(const vs_mask_entry *__restrict__ e, uint64_t n, uint64_t word_base, vs_masks *__restrict__ M)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        vs_mask_entry x = e[i];
        M[x.word - word_base] = vs_masks{x.iv, x.lw};
    }
}

// ------------------------------------------------------------------------------------------------
// Scoring.
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}
#define VS_FA(a, b, c, s, cy) uint32_t s = lop3<0x96>(a, b, c), cy = lop3<0xE8>(a, b, c)

// byte offsets (plane index * SCORE_THREADS * 4) of the plane selected by each pattern base
__constant__ uint32_t c_pat[PAT_CHUNK * PAT_STRIDE];

struct ScoreArgs {
    const uint32_t *planes;     // [n_blocks][48]
    const uint32_t *pos;        // [n_blocks][32]
    const unsigned long long *n_blocks_ptr;   // blocks claimed by k_extract for this chunk and strand (device counter)
    uint64_t cap;               // capacity of the candidate store; a chunk that overflowed it is skipped (host redoes it)
    uint32_t n_pat;             // patterns in this launch (<= PAT_CHUNK)
    uint32_t guide_base;        // index of pattern 0 of this launch in the guide list
    uint32_t strand;            // 0 forward pass, 1 reverse pass
    vs_hit *hits;
    unsigned long long *n_hits;
    uint64_t hit_cap;
};

template <int K>
__global__ void __launch_bounds__(SCORE_THREADS, 2)
k_score(ScoreArgs a)
{
    extern __shared__ uint32_t sm[];     // [NPLANES][SCORE_THREADS]
    const int tid = threadIdx.x;
    const uint64_t n_blocks = *a.n_blocks_ptr;
    if (n_blocks > a.cap || (uint64_t)blockIdx.x * SCORE_THREADS >= n_blocks) return;     // grid is sized by capacity
    const uint64_t blk = (uint64_t)blockIdx.x * SCORE_THREADS + tid;
    uint32_t *my = sm + tid;
    uint32_t lastm = 0;
    if (blk < n_blocks) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.planes + blk * BLK_WORDS);
        uint32_t v[BLK_WORDS];
#pragma unroll
        for (int i = 0; i < BLK_WORDS / 4; ++i) {
            uint4 t = __ldg(src + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
        lastm = v[BLK_LAST];
        const uint32_t inv = ~v[BLK_VALID];
#pragma unroll
        for (int i = 0; i < VS_GLEN; ++i) {
            const uint32_t h = v[i], l = v[VS_GLEN + i];
            const uint32_t x = (i < 9) ? inv : 0u;     // 9 forced mismatches keep invalid lanes above any k <= 8
            my[(4 * i + 0) * SCORE_THREADS] = (h | l) | x;      // mismatch if guide base is A (00)
            my[(4 * i + 1) * SCORE_THREADS] = (h | ~l) | x;     // C (01)
            my[(4 * i + 2) * SCORE_THREADS] = (~h | l) | x;     // G (10)
            my[(4 * i + 3) * SCORE_THREADS] = (~h | ~l) | x;    // T (11)
        }
    } else {
#pragma unroll
        for (int i = 0; i < NPLANES; ++i) my[i * SCORE_THREADS] = ~0u;
    }
    // each thread reads back only what it wrote: no barrier needed
    const char *myb = reinterpret_cast<const char *>(my);
    constexpr int LUT_F = (K < 8) ? ((1 << (K + 1)) - 1) : 0x01;   // f(b2,b1,b0): low3 <= K  (K = 8: low3 == 0)
    constexpr int LUT_G = (K < 8) ? 0x02 : 0x2B;                   // g(x,y,f): ~x & ~y & f   (K = 8: (~x&~y) | ((x^y)&f))

#pragma unroll 2
    for (uint32_t g = 0; g < a.n_pat; ++g) {
        const uint32_t *po = c_pat + g * PAT_STRIDE;
        uint32_t m[VS_GLEN];
#pragma unroll
        for (int i = 0; i < VS_GLEN; ++i) m[i] = *reinterpret_cast<const uint32_t *>(myb + po[i]);
        // carry-save adder tree: 23 one-bit planes -> b0, b1, b2 and two weight-8 planes x, y  (18 full adders)
        VS_FA(m[0], m[1], m[2], s0, c0);
        VS_FA(m[3], m[4], m[5], s1, c1);
        VS_FA(m[6], m[7], m[8], s2, c2);
        VS_FA(m[9], m[10], m[11], s3, c3);
        VS_FA(m[12], m[13], m[14], s4, c4);
        VS_FA(m[15], m[16], m[17], s5, c5);
        VS_FA(m[18], m[19], m[20], s6, c6);
        VS_FA(s0, s1, s2, t0, d0);
        VS_FA(s3, s4, s5, t1, d1);
        VS_FA(s6, m[21], m[22], t2, d2);
        VS_FA(t0, t1, t2, b0, d3);
        VS_FA(c0, c1, c2, u0, e0);
        VS_FA(c3, c4, c5, u1, e1);
        VS_FA(c6, d0, d1, u2, e2);
        VS_FA(u0, u1, u2, v0, e3);
        VS_FA(v0, d2, d3, b1, e4);
        VS_FA(e0, e1, e2, p0, x);
        VS_FA(p0, e3, e4, b2, y);
        const uint32_t f = lop3<LUT_F>(b2, b1, b0);
        uint32_t le = lop3<LUT_G>(x, y, f);
        if (le) {
            // slow path (rare): exact count, R4 for last windows, append
            do {
                const int c = __ffs(le) - 1;
                le &= le - 1;
                uint32_t mm = 0, h2 = 0;
#pragma unroll
                for (int i = 0; i < VS_GLEN; ++i) {
                    uint32_t bit = (*reinterpret_cast<const uint32_t *>(myb + po[i]) >> c) & 1u;
                    mm += bit;
                    if (i >= 11) h2 += bit;
                }
                if (((lastm >> c) & 1u) && h2 > (uint32_t)(K / 2)) continue;     // R4
                unsigned long long idx = atomicAdd(a.n_hits, 1ull);
                if (idx < a.hit_cap) {
                    vs_hit hrec;
                    hrec.pos = a.pos[blk * 32 + c];
                    hrec.info = ((a.guide_base + g) << 8) | (a.strand << 7) | mm;
                    a.hits[idx] = hrec;
                }
            } while (le);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Microbenchmarks for the roofline denominators (alu-pipe LOP3 issue rate, shared-memory LDS rate).
__global__ void __launch_bounds__(256)
k_peak_lop3(uint32_t *out, int iters)
{
    uint32_t a0 = threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3, a4 = a0 * 11 + 4, a5 = a0 * 13 + 5, a6 = a0 * 17 + 6, a7 = a0 * 19 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = lop3<0x96>(a0, a1, a2); a1 = lop3<0xE8>(a1, a2, a3); a2 = lop3<0x96>(a2, a3, a4); a3 = lop3<0xE8>(a3, a4, a5);
            a4 = lop3<0x96>(a4, a5, a6); a5 = lop3<0xE8>(a5, a6, a7); a6 = lop3<0x96>(a6, a7, a0); a7 = lop3<0xE8>(a7, a0, a1);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

__global__ void __launch_bounds__(256)
k_peak_lds(uint32_t *out, int iters)
{
    __shared__ uint32_t s[256 * 8];
    for (int i = threadIdx.x; i < 256 * 8; i += 256) s[i] = i * 2654435761u;
    __syncthreads();
    uint32_t acc = 0;
    const volatile uint32_t *p = s + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc ^= p[u * 256];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace vs
