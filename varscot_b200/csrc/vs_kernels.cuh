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

// Candidate masks for the 32 window starts of word a (b = next word).
//   R1: the window must not run over a contig end (em) -> no end bit in [p, p+22)
//   R3: no N in [p, p+23)
//   R2: PAM on the genome: forward W[21..22], reverse W[0..1]   (bidir_mapping.cpp:70-76, :240-247)
//   last = window ends exactly at a contig end (R4 needs H2 <= K there, bidir_mapping.cpp:51)
__device__ __forceinline__ void cand_masks(const vs_word &a, const vs_word &b, const PamParams &pp,
                                           uint32_t &fwd, uint32_t &rev, uint32_t &last)
{
    uint64_t N = ((uint64_t)b.nm << 32) | a.nm;
    uint64_t E = ((uint64_t)b.em << 32) | a.em;
    uint64_t H = ((uint64_t)b.hi << 32) | a.hi;
    uint64_t L = ((uint64_t)b.lo << 32) | a.lo;
    uint64_t t = N | (N >> 1); t |= t >> 2; t |= t >> 4; t |= t >> 8;   // OR over 16 consecutive
    uint64_t n23 = t | (t >> 7);                                          // OR over 23
    uint64_t e = E | (E >> 1); e |= e >> 2; e |= e >> 4; e |= e >> 8;
    uint64_t e22 = e | (e >> 6);                                          // OR over 22
    uint32_t inv = (uint32_t)(n23 | e22);
    last = (uint32_t)(E >> 22);
    uint32_t h21 = (uint32_t)(H >> 21), h22 = (uint32_t)(H >> 22), l21 = (uint32_t)(L >> 21), l22 = (uint32_t)(L >> 22);
    uint32_t h0 = a.hi, l0 = a.lo, h1 = (uint32_t)(H >> 1), l1 = (uint32_t)(L >> 1);
    uint32_t f = 0, r = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (j < pp.n) {
            f |= eq_plane(h21, l21, pp.fx[j]) & eq_plane(h22, l22, pp.fy[j]);
            r |= eq_plane(h0, l0, pp.rx[j]) & eq_plane(h1, l1, pp.ry[j]);
        }
    }
    fwd = f & ~inv;
    rev = r & ~inv;
}

// ------------------------------------------------------------------------------------------------
// 32 x 32 bit-matrix transpose in registers (LSB-first): out[i] bit c = in[c] bit i.
// 5 butterfly stages x 16 swaps; ptxas drops the swaps that only feed unused outputs.
__host__ __device__ __forceinline__ void transpose32(uint32_t (&a)[32])
{
    uint32_t m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j; j >>= 1, m ^= m << j) {
#pragma unroll
        for (int k = 0; k < 32; k = ((k | j) + 1) & ~j) {
            uint32_t t = ((a[k] >> j) ^ a[k | j]) & m;
            a[k | j] ^= t;
            a[k] ^= t << j;
        }
    }
}

constexpr int Q_STRIDE = 33;                                   // halfwords per queued block: odd stride -> conflict-free reads
constexpr int Q_CAP    = (TILE_STARTS / 32) * Q_STRIDE;

// k_extract: one CTA per tile of 8192 window starts.
//   phase 1 (all threads, bit-parallel over 32 starts each): candidate masks, CTA-wide prefix sum, queue the
//            candidates of each strand in shared memory as (local start | last << 15);
//   phase 2 (one THREAD per 32-candidate block): gather the 23-base windows from the staged planes, transpose
//            32 x 23 bits twice in registers, write 48 plane words + 32 positions.
// Block ranges are claimed with one atomicAdd per strand per tile on cnt[2], cnt[3] (layout order is not
// deterministic; hit resolution sorts).  If a claim runs past the capacity nothing is written and the host,
// which reads the counters back, grows the stores and launches again.
// Block layout (48 words): hi_0..hi_22, lo_0..lo_22, last-window mask, valid mask.
__global__ void __launch_bounds__(TILE_THREADS)
k_extract(const vs_word *__restrict__ W, uint64_t n_words, uint64_t global_base, PamParams pp,
          uint32_t *__restrict__ planes_f, uint32_t *__restrict__ pos_f, uint64_t cap_f,
          uint32_t *__restrict__ planes_r, uint32_t *__restrict__ pos_r, uint64_t cap_r,
          unsigned long long *__restrict__ cnt)
{
    __shared__ uint2 s_hl[TILE_WORDS + 1];
    __shared__ uint16_t q[2][Q_CAP];
    __shared__ uint32_t wsum[2][TILE_THREADS / 32];
    __shared__ uint32_t tot[2];
    __shared__ unsigned long long base[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t w0 = (uint64_t)blockIdx.x * TILE_WORDS;
    const uint64_t w = w0 + tid;
    uint32_t fwd = 0, rev = 0, last = 0;
    {
        vs_word a = {0u, 0u, ~0u, 0u}, b = {0u, 0u, ~0u, 0u};
        if (w <= n_words) a = W[w];          // w == n_words is the halo / pad word, always readable
        if (w < n_words) b = W[w + 1];
        s_hl[tid] = make_uint2(a.hi, a.lo);
        if (tid == TILE_THREADS - 1) s_hl[TILE_WORDS] = make_uint2(b.hi, b.lo);
        if (w < n_words) cand_masks(a, b, pp, fwd, rev, last);
    }
    // CTA-wide exclusive scan of the per-thread candidate counts
    const uint32_t cf = __popc(fwd), cr = __popc(rev);
    uint32_t xf = cf, xr = cr;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t a = __shfl_up_sync(0xffffffffu, xf, o), b = __shfl_up_sync(0xffffffffu, xr, o);
        if (lane >= o) { xf += a; xr += b; }
    }
    if (lane == 31) { wsum[0][wid] = xf; wsum[1][wid] = xr; }
    __syncthreads();
    uint32_t bf = 0, br = 0;
#pragma unroll
    for (int i = 0; i < TILE_THREADS / 32; ++i) {
        uint32_t a = wsum[0][i], b = wsum[1][i];
        if (i < wid) { bf += a; br += b; }
    }
    if (tid == TILE_THREADS - 1) {
        const uint32_t nf_ = bf + xf, nr_ = br + xr;
        tot[0] = nf_; tot[1] = nr_;
        // claim block ranges; candidate totals are for the statistics only
        base[0] = atomicAdd(&cnt[2], (unsigned long long)((nf_ + 31) >> 5));
        base[1] = atomicAdd(&cnt[3], (unsigned long long)((nr_ + 31) >> 5));
        if (nf_) atomicAdd(&cnt[0], (unsigned long long)nf_);
        if (nr_) atomicAdd(&cnt[1], (unsigned long long)nr_);
    }
    {
        uint32_t of = bf + xf - cf, orr = br + xr - cr;
        uint32_t m = fwd;
        while (m) {
            int b = __ffs(m) - 1; m &= m - 1;
            q[0][(of >> 5) * Q_STRIDE + (of & 31)] = (uint16_t)((tid << 5) | b | (((last >> b) & 1u) << 15));
            ++of;
        }
        m = rev;
        while (m) {
            int b = __ffs(m) - 1; m &= m - 1;
            q[1][(orr >> 5) * Q_STRIDE + (orr & 31)] = (uint16_t)((tid << 5) | b | (((last >> b) & 1u) << 15));
            ++orr;
        }
    }
    __syncthreads();
    const uint32_t nf = tot[0], nr = tot[1];
    const uint32_t nbf = (nf + 31) >> 5, nbr = (nr + 31) >> 5;
    const uint32_t gbase = (uint32_t)(global_base + w0 * 32);
    for (uint32_t j = tid; j < nbf + nbr; j += TILE_THREADS) {
        const int s = j >= nbf;
        const uint32_t jj = s ? j - nbf : j;
        const uint32_t n = s ? nr : nf;
        const uint64_t blk = base[s] + jj;
        if (blk >= (s ? cap_r : cap_f)) continue;          // overflow: host regrows and relaunches
        const uint32_t cntc = min(32u, n - jj * 32);
        const uint16_t *qq = &q[s][jj * Q_STRIDE];
        uint32_t *pl_out = (s ? planes_r : planes_f) + blk * BLK_WORDS;
        uint4 *ps_out = reinterpret_cast<uint4 *>((s ? pos_r : pos_f) + blk * 32);
        uint32_t a[32];
        uint32_t lastw = 0;
        // hi planes (+ positions, last mask)
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            uint32_t pp4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c4 * 4 + u;
                uint32_t hw = 0, ps = 0xFFFFFFFFu;
                if ((uint32_t)c < cntc) {
                    const uint32_t e = qq[c];
                    const uint32_t lp = e & 0x1FFFu, wi = lp >> 5, o = lp & 31;
                    hw = __funnelshift_r(s_hl[wi].x, s_hl[wi + 1].x, o) & 0x7FFFFFu;
                    lastw |= (e >> 15) << c;
                    ps = gbase + lp;
                }
                a[c] = hw; pp4[u] = ps;
            }
            ps_out[c4] = make_uint4(pp4[0], pp4[1], pp4[2], pp4[3]);
        }
        transpose32(a);
        uint32_t carry = a[20], carry1 = a[21], carry2 = a[22];   // words 20..22 are stored with the first lo planes
#pragma unroll
        for (int i = 0; i < 5; ++i)
            reinterpret_cast<uint4 *>(pl_out)[i] = make_uint4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
        // lo planes
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            uint32_t lw = 0;
            if ((uint32_t)c < cntc) {
                const uint32_t lp = qq[c] & 0x1FFFu, wi = lp >> 5, o = lp & 31;
                lw = __funnelshift_r(s_hl[wi].y, s_hl[wi + 1].y, o) & 0x7FFFFFu;
            }
            a[c] = lw;
        }
        transpose32(a);
        // words 20..47: hi_20, hi_21, hi_22, lo_0 .. lo_22, last, valid
        reinterpret_cast<uint4 *>(pl_out)[5] = make_uint4(carry, carry1, carry2, a[0]);
#pragma unroll
        for (int i = 0; i < 5; ++i)
            reinterpret_cast<uint4 *>(pl_out)[6 + i] = make_uint4(a[4 * i + 1], a[4 * i + 2], a[4 * i + 3], a[4 * i + 4]);
        const uint32_t validw = cntc >= 32 ? ~0u : ((1u << cntc) - 1u);
        reinterpret_cast<uint4 *>(pl_out)[11] = make_uint4(a[21], a[22], lastw, validw);
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
    uint64_t n_blocks;
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
    const uint64_t blk = (uint64_t)blockIdx.x * SCORE_THREADS + tid;
    uint32_t *my = sm + tid;
    uint32_t lastm = 0;
    if (blk < a.n_blocks) {
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
