// varscot_b200/csrc/vs_kernels.cuh — hand-written sm_100a kernels of the off-target scan.
//
// What they replace: the inner loops of VARSCOT_pipeline/read_mapping/bidir_mapping.cpp —
// SeqAn's find<0,K>(delegate, index, half, HammingDistance()) (:129-146) plus the verify
// delegate (:34-127) — restated as a dense, PAM-first Hamming scan (rules R1-R4 of SURVEY.md 8a).
//
// Pipeline per scan, per pipeline chunk of the text (8 Mi words), all on one stream:
//   k_extract  : per tile of ~240 words (7680 window starts), find the PAM-valid / scannable windows of each strand
//                (bit-parallel on the packed planes and the precomputed window masks), compact them into blocks of 32
//                candidates and store each block bit-sliced ACROSS candidates (per-thread 32x32 register
//                transposes): 48 words per block = hi_0..hi_22, lo_0..lo_22, last-window mask, valid mask; + 32 positions
//   k_score<K> : one thread per candidate block, both strands in one launch; the 46 planes are expanded into 92
//                "mismatch if the guide base at position i is b" planes in shared memory; per guide the select is one
//                LDS per position at a warp-uniform offset from constant memory, the count a bit-sliced carry-save
//                adder tree (2 LOP3 per full adder), the threshold 2 LOP3; two stages with a warp-uniform early out.
//                Hits (rare): count read from the bit-sliced counter, R4 check on last windows, atomic append.
// Integer pipe + shared-memory bound; no tensor cores (nothing here is a dense contraction worth a GEMM:
// the bit-sliced form costs ~1.1 ALU ops per (window, guide) pair, below one op per output element).
#pragma once
#include <cstdint>
#include <type_traits>
#ifndef VS_HOST_UNIT_TEST          // tests/cpu_kernel_units.cpp and tests/cpu_scan_emulator.cpp compile this header with g++ and run the
                                   // kernels on the host (k_extract with one OS thread per CUDA thread, the others thread by thread)
#include <cuda_runtime.h>
#endif
#include "../../include/varscot_scan.h"

namespace vs {

constexpr int BLK_WORDS    = 48;                  // words per candidate block
constexpr int BLK_GROUP    = 32;                  // blocks per layout group: word w of block b sits at plane_index(b, w), so that
                                                  // the 32 lanes of a warp (32 consecutive blocks) touch 128 contiguous bytes per word
constexpr int BLK_LAST     = 46;                  // word index of the last-window mask
constexpr int BLK_VALID    = 47;                  // word index of the valid mask
#ifndef VS_SCORE_THREADS
#define VS_SCORE_THREADS 128          // measured on B200 (score ms, cfg3 x 0.25 / cfg4 x 0.1): 96 thr 2.11 / 12.7, 128 thr 2.15 / 11.0, 256 thr 2.20 / 11.2
#endif
constexpr int SCORE_THREADS = VS_SCORE_THREADS;
constexpr int NPLANES      = 4 * VS_GLEN;         // 92 "mismatch if the guide base at position i is b" planes per block (k = 8: all in shared memory)
constexpr int PAT_STRIDE   = 24;                  // uint32 per pattern in constant memory (23 slot offsets + pad, 16-byte aligned)
#ifndef VS_PAT_CHUNK
#define VS_PAT_CHUNK 256
#endif
constexpr int PAT_CHUNK    = VS_PAT_CHUNK;        // guides per k_score launch: 2 strands x 256 x 96 B = 48 KB of constant memory
constexpr int PAT_TABLE_WORDS = 2 * PAT_CHUNK * PAT_STRIDE;             // words of c_pat = words per guide chunk

__host__ __device__ __forceinline__ uint64_t plane_index(uint64_t blk, int w)
{
    return (blk / BLK_GROUP) * (uint64_t)(BLK_WORDS * BLK_GROUP) + (uint64_t)w * BLK_GROUP + (blk % BLK_GROUP);
}

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

// 16 x 32 bit-matrix, both 16 x 16 halves transposed in place: out[r] bits 0..15 = bit r of in[0..15], bits 16..31 =
// bit 16 + r of in[0..15] (transpose32 without its first stage, on 16 rows).
__device__ __forceinline__ void transpose16(uint32_t (&a)[16])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {                        // bytes: odd bytes of a[k] <-> even bytes of a[k+8]
        const uint32_t x = a[k], y = a[k | 8];
        a[k] = __byte_perm(x, y, 0x6240);
        a[k | 8] = __byte_perm(x, y, 0x7351);
    }
    uint32_t m = 0x0F0F0F0Fu;
#pragma unroll
    for (int j = 4; j; j >>= 1, m ^= m << j) {
#pragma unroll
        for (int k = 0; k < 16; k = ((k | j) + 1) & ~j) {
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
// Block layout (48 words): hi_0..hi_22, lo_0..lo_22, last-window mask, valid mask; word w of block b at plane_index(b, w).
#ifndef VS_EX_HALF
#define VS_EX_HALF 0                  // 1: EXPERIMENTAL, not yet run on a GPU — one thread per half block (see phase 2 below)
#endif
#ifndef VS_EX_MINBLOCKS
#define VS_EX_MINBLOCKS (VS_EX_HALF ? 16 : 10)
#endif
__global__ void __launch_bounds__(EX_THREADS, VS_EX_MINBLOCKS)
k_extract(const vs_bases *__restrict__ B, const vs_masks *__restrict__ M, uint64_t w_begin, uint64_t w_end, uint32_t tile_words,
          uint64_t global_base, PamParams pp,
          uint32_t *__restrict__ planes_f, uint32_t *__restrict__ pos_f,
          uint32_t *__restrict__ planes_r, uint32_t *__restrict__ pos_r, uint64_t cap,
          unsigned long long *__restrict__ cnt)
{
    __shared__ __align__(16) uint2 s_hl[EX_MAX_WORDS + 2];   // bases {hi, lo} per word (+ halo, + one word read under the sentinel)
    __shared__ uint32_t s_m[2][EX_MAX_WORDS + 1];      // candidate masks per strand
    __shared__ uint32_t s_p[2][EX_MAX_WORDS + 1];      // exclusive rank prefix per word (+ total at [nw])
    __shared__ __align__(16) uint2 s_mk[EX_MAX_WORDS + 2];   // window masks {iv, lw} per word (+ one word read under the sentinel)
    __shared__ uint32_t wsum[2][EX_MAX_WORDS / EX_THREADS][EX_THREADS / 32];
    __shared__ unsigned long long base[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t w0 = w_begin + (uint64_t)blockIdx.x * tile_words;
    const uint32_t nw = (uint32_t)min((uint64_t)tile_words, w_end - w0);      // words of this tile
    // phase 1a: the tile's bases (+ halo word) and window masks are staged into shared memory with cp.async (LDGSTS):
    // coalesced 128-bit copies when the tile starts on a 16-byte boundary (always, except for odd test chunk sizes),
    // 64-bit copies otherwise; nothing passes through registers and every copy of the tile is in flight at once.
#ifndef VS_HOST_UNIT_TEST
    {
        const uint32_t sb = (uint32_t)__cvta_generic_to_shared(s_hl), smk = (uint32_t)__cvta_generic_to_shared(s_mk);
        const char *gb = reinterpret_cast<const char *>(B + w0), *gm = reinterpret_cast<const char *>(M + w0);
        const uint32_t nb_bytes = (nw + 1) * 8u, nm_bytes = nw * 8u;        // B[w_end] is the halo / pad word, always readable
        if (((w0 & 1) == 0)) {
            for (uint32_t o = tid * 16u; o + 16u <= nb_bytes; o += EX_THREADS * 16u)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb + o), "l"(gb + o));
            for (uint32_t o = tid * 16u; o + 16u <= nm_bytes; o += EX_THREADS * 16u)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smk + o), "l"(gm + o));
            if (tid == 0 && (nb_bytes & 8u)) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sb + nb_bytes - 8u), "l"(gb + nb_bytes - 8u));
            if (tid == 1 && (nm_bytes & 8u)) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smk + nm_bytes - 8u), "l"(gm + nm_bytes - 8u));
        } else {
            for (uint32_t o = tid * 8u; o < nb_bytes; o += EX_THREADS * 8u)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sb + o), "l"(gb + o));
            for (uint32_t o = tid * 8u; o < nm_bytes; o += EX_THREADS * 8u)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smk + o), "l"(gm + o));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
#else       // host emulation (tests/cpu_kernel_units.cpp): plain copies
    for (uint32_t i = tid; i < nw + 1; i += EX_THREADS) s_hl[i] = uint2{B[w0 + i].hi, B[w0 + i].lo};
    for (uint32_t i = tid; i < nw; i += EX_THREADS) s_mk[i] = uint2{M[w0 + i].iv, M[w0 + i].lw};
#endif
    __syncthreads();
    constexpr int EX_ITERS = EX_MAX_WORDS / EX_THREADS;
    // phase 1b: candidate masks and their rank prefix (word order: thread t owns words t, t+64, t+128, t+192)
    uint32_t cf[EX_ITERS], cr[EX_ITERS], xf[EX_ITERS], xr[EX_ITERS];
#pragma unroll
    for (int it = 0; it < EX_ITERS; ++it) {
        const uint32_t i = it * EX_THREADS + tid;
        uint32_t fwd = 0, rev = 0;
        if (i < nw) {
            const uint2 cur = s_hl[i], nx = s_hl[i + 1], mk = s_mk[i];
            cand_masks(vs_bases{cur.x, cur.y}, vs_bases{nx.x, nx.y}, vs_masks{mk.x, mk.y}, pp, fwd, rev);
            s_m[0][i] = fwd; s_m[1][i] = rev;
        }
        cf[it] = __popc(fwd); cr[it] = __popc(rev);
        uint32_t a = cf[it], b = cr[it];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t p = __shfl_up_sync(0xffffffffu, a, o), q = __shfl_up_sync(0xffffffffu, b, o);
            if (lane >= o) { a += p; b += q; }
        }
        xf[it] = a; xr[it] = b;
        if (lane == 31) { wsum[0][it][wid] = a; wsum[1][it][wid] = b; }
    }
    __syncthreads();
    uint32_t run_f = 0, run_r = 0;                     // candidates before the current group of 64 words
#pragma unroll
    for (int it = 0; it < EX_ITERS; ++it) {
        const uint32_t i = it * EX_THREADS + tid;
        const uint32_t w0f = wsum[0][it][0], w0r = wsum[1][it][0], w1f = wsum[0][it][1], w1r = wsum[1][it][1];
        if (i < nw) {
            s_p[0][i] = run_f + (wid ? w0f : 0u) + xf[it] - cf[it];
            s_p[1][i] = run_r + (wid ? w0r : 0u) + xr[it] - cr[it];
        }
        run_f += w0f + w1f; run_r += w0r + w1r;
    }
    const uint32_t nf = run_f, nr = run_r;
    const uint32_t nbf = (nf + 31) >> 5, nbr = (nr + 31) >> 5;
    if (tid == 0) {
        s_p[0][nw] = nf; s_p[1][nw] = nr;
        s_m[0][nw] = ~0u; s_m[1][nw] = ~0u;              // sentinel: ends the mask walk of a partial block (see phase 2)
        base[0] = atomicAdd(&cnt[2], (unsigned long long)nbf);
        base[1] = atomicAdd(&cnt[3], (unsigned long long)nbr);
        if (nf) atomicAdd(&cnt[0], (unsigned long long)nf);
        if (nr) atomicAdd(&cnt[1], (unsigned long long)nr);
    }
    __syncthreads();
    const uint32_t gbase = (uint32_t)(global_base + w0 * 32);      // device word 0 sits at global_base
#if VS_EX_HALF
    // EXPERIMENTAL (compiled only with -DVS_EX_HALF=1, never run on a GPU so far): one thread per HALF block of 16
    // candidates.  Half the registers (two 16-row transposes instead of two 32-row ones) let twice as many warps reside,
    // which is what this latency-bound kernel lacks; the plane words leave as 16-bit halves (lane pairs write one word).
    for (uint32_t j2 = tid; j2 < 2 * (nbf + nbr); j2 += EX_THREADS) {
#include "vs_extract_half_block.inc"
    }
#else
    for (uint32_t j = tid; j < nbf + nbr; j += EX_THREADS) {
#include "vs_extract_block.inc"
    }
#endif
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
// Window masks on the device, from the compact mask source of a view (include/varscot_scan.h).
// k_fill_runs: plane[word - word_base + i] = value for the part of every run {word, count, value} that lies inside
// [lo, hi) (global word indices).  Warp (blockIdx.x * 8 + warp) owns a run, blockIdx.y selects a segment of FILL_SEG
// words of it, so that a long run (an N stretch of a chromosome is 10^5 words) is filled by many warps.
constexpr uint32_t FILL_SEG = 1024;
__global__ void __launch_bounds__(256)
k_fill_runs(const vs_plane_run *__restrict__ runs, uint64_t n_runs, uint64_t lo, uint64_t hi, uint64_t word_base, uint32_t *__restrict__ plane)
{
    const uint64_t r = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n_runs) return;
    const vs_plane_run x = runs[r];
    const uint64_t a = max((uint64_t)x.word, lo) + (uint64_t)blockIdx.y * FILL_SEG;
    const uint64_t b = min(min((uint64_t)x.word + x.count, hi), a + FILL_SEG);
    for (uint64_t w = a + (threadIdx.x & 31); w < b; w += 32) plane[w - word_base] = x.value;
}

// k_expand_em_code: contig-end plane words from their one-byte codes, for the words of blocks that travel coded
// (code = index of the only set bit, VS_EM_NONE = no single bit; words with several bits arrive as runs afterwards).
// code / out point at the first word of the range (global word index g0), dense at block 0 of the text.
__global__ void __launch_bounds__(256)
k_expand_em_code(const uint8_t *__restrict__ code, const uint8_t *__restrict__ dense, uint64_t g0, uint64_t n, uint32_t *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !dense[(g0 + i) / VS_EM_BLOCK]) return;
    const uint32_t c = code[i];
    out[i] = c < 32u ? 1u << c : 0u;
}

// k_masks_from_planes: the device twin of masks_of() in vs_host.cpp.  nm / em point at the first word of the range and
// hold n + 1 words (the last one is the halo);
//   iv: any N in [p, p+23)  or  any contig end in [p, p+22)   (R1, R3)
//   lw: the window is valid and its last base (p+22) is a contig end   (R4)
__global__ void __launch_bounds__(256)
k_masks_from_planes(const uint32_t *__restrict__ nm, const uint32_t *__restrict__ em, uint64_t n, vs_masks *__restrict__ out)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const uint64_t N = ((uint64_t)nm[w + 1] << 32) | nm[w], E = ((uint64_t)em[w + 1] << 32) | em[w];
    uint64_t t = N | (N >> 1); t |= t >> 2; t |= t >> 4; t |= t >> 8;      // OR over 16 consecutive
    const uint64_t n23 = t | (t >> 7);                                       // OR over 23
    uint64_t e = E | (E >> 1); e |= e >> 2; e |= e >> 4; e |= e >> 8;
    const uint64_t e22 = e | (e >> 6);                                       // OR over 22
    vs_masks m;
    m.iv = (uint32_t)(n23 | e22);
    m.lw = (uint32_t)(E >> 22) & ~m.iv;
    out[w] = m;
}

// ------------------------------------------------------------------------------------------------
// Scoring.
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
#ifndef VS_HOST_UNIT_TEST
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
#else       // the instruction's definition: bit (4a + 2b + c) of the table, per bit position
    uint32_t r = 0;
    for (int i = 0; i < 8; ++i)
        if ((LUT >> i) & 1) r |= ((i & 4) ? a : ~a) & ((i & 2) ? b : ~b) & ((i & 1) ? c : ~c);
    return r;
#endif
}

// Bit-sliced population count of N one-bit planes (plus optional planes `init[w]` already carrying weight 2^w):
// column compression with full adders (2 LOP3 each: XOR3 0x96, MAJ3 0xE8) taken from the front of a queue so the
// tree stays balanced.  Written as template recursion so that every index is a compile-time constant and all
// planes live in registers.
template <int HEAD, int TAIL, int NN>
struct CsaColumn {
    static constexpr int LEFT = TAIL - HEAD;
    using Next = CsaColumn<(LEFT >= 3 ? HEAD + 3 : HEAD + 2), TAIL + 1, NN + 1>;
    static __device__ __forceinline__ void run(uint32_t *col, uint32_t *nxt)
    {
        if constexpr (LEFT >= 3) {
            const uint32_t x = col[HEAD], y = col[HEAD + 1], z = col[HEAD + 2];
            col[TAIL] = lop3<0x96>(x, y, z);
            nxt[NN] = lop3<0xE8>(x, y, z);
            Next::run(col, nxt);
        } else if constexpr (LEFT == 2) {
            const uint32_t x = col[HEAD], y = col[HEAD + 1];
            col[TAIL] = x ^ y;
            nxt[NN] = x & y;
            Next::run(col, nxt);
        }
    }
    static __host__ __device__ constexpr int final_head() { if constexpr (LEFT >= 2) return Next::final_head(); else return HEAD; }
    static __host__ __device__ constexpr int final_left() { if constexpr (LEFT >= 2) return Next::final_left(); else return LEFT; }
    static __host__ __device__ constexpr int final_nn() { if constexpr (LEFT >= 2) return Next::final_nn(); else return NN; }
};

template <int W, int N_IN, bool HAS_INIT>
struct CsaWeights {
    static __device__ __forceinline__ void run(uint32_t *col, const uint32_t *init, uint32_t *bit)
    {
        constexpr int T0 = N_IN + (HAS_INIT ? 1 : 0);
        if constexpr (HAS_INIT) col[N_IN] = init[W];
        using C = CsaColumn<0, T0, 0>;
        uint32_t nxt[T0 + 2];
        C::run(col, nxt);
        if constexpr (C::final_left() == 1) bit[W] = col[C::final_head()]; else bit[W] = 0u;
        if constexpr (W < 4) {
            constexpr int NN = C::final_nn();
            uint32_t col2[2 * NN + 4];
#pragma unroll
            for (int i = 0; i < NN; ++i) col2[i] = nxt[i];
            CsaWeights<W + 1, NN, HAS_INIT>::run(col2, init, bit);
        }
    }
};

template <int N, bool HAS_INIT>
__device__ __forceinline__ void popcount_planes(const uint32_t (&in)[N], const uint32_t (&init)[5], uint32_t (&bit)[5])
{
    uint32_t col[2 * N + 4];
#pragma unroll
    for (int i = 0; i < N; ++i) col[i] = in[i];
    CsaWeights<0, N, HAS_INIT>::run(col, init, bit);
}

// count <= K on a 5-bit bit-sliced count (K <= 8): two LOP3 with compile-time LUTs
template <int K>
__device__ __forceinline__ uint32_t le_k(const uint32_t (&b)[5])
{
    constexpr int LUT_F = (K < 8) ? ((1 << (K + 1)) - 1) : 0x01;   // f(b2,b1,b0): low3 <= K      (K = 8: low3 == 0)
    constexpr int LUT_G = (K < 8) ? 0x02 : 0x0B;                   // g(b4,b3,f): ~b4 & ~b3 & f   (K = 8: ~b4 & (~b3 | f))
    const uint32_t f = lop3<LUT_F>(b[2], b[1], b[0]);
    return lop3<LUT_G>(b[4], b[3], f);
}

// Number of pattern slots scored before the early-out test: with uniform-random text, after PA(K) informative
// positions fewer than ~12 % of the warps (1024 candidates) still hold a window with <= K mismatches, so the
// remaining slots (the PAM positions come last in the slot order) are loaded only for those.
__host__ __device__ constexpr int stage_a_slots(int k) { return k >= 8 ? VS_GLEN : 7 + 2 * k; }
// Shared-memory planes per block: the 4 expanded "mismatch if the pattern base is b" planes of every stage-A position
// ([base, base + PA) with base = 0 forward, 2 reverse, whose slot order is 2..22, 0, 1), then the two RAW planes {hi, lo}
// of every stage-B slot — the few warps that reach stage B pay 2 LDS + 2 LOP3 per slot instead of 1 LDS, and the block
// needs 4 PA + 2 PB planes instead of 92 (k = 6: 84), which lets one more CTA reside per SM.  k = 8 is single-stage.
__host__ __device__ constexpr int score_smem_planes(int k) { return 4 * stage_a_slots(k) + 2 * (VS_GLEN - stage_a_slots(k)); }
__host__ __device__ constexpr int score_pos_base(int k, int strand) { return (k < 8 && strand) ? 2 : 0; }
__host__ __device__ constexpr int score_min_blocks(int k)
{
    // resident CTAs per SM that fit 227 KB of shared memory (1 KB reserved per CTA) and 64 K registers at <= 72 per thread
    const int by_smem = (227 * 1024) / (score_smem_planes(k) * SCORE_THREADS * 4 + 1024);
    const int by_regs = 65536 / (72 * SCORE_THREADS);
#ifdef VS_SCORE_MINBLOCKS
    return VS_SCORE_MINBLOCKS;          // tuning override: only steers the register allocation
#endif
    const int m = by_smem < by_regs ? by_smem : by_regs;
    return m < 1 ? 1 : (m > 8 ? 8 : m);
}

// Pattern table in constant memory, per strand and guide: PAT_STRIDE words.  The host orders the slots
// informative-first: forward positions 0..22, reverse 2..22,0,1.  A stage-A slot j < PA(k) holds the byte offset
// (plane index * SCORE_THREADS * 4, plane index = 4 * (position - base) + pattern base) of the shared-memory plane it
// selects; a stage-B slot holds the byte offset of its raw hi plane (lo follows at + SCORE_THREADS * 4) | pattern base << 16
// (see pat_slot()).
__constant__ uint32_t c_pat[PAT_TABLE_WORDS];

// position scored by slot j of a strand's slot order
__host__ __device__ constexpr int slot_position(int strand, int j) { return strand ? (j < VS_GLEN - 2 ? j + 2 : j - (VS_GLEN - 2)) : j; }
// table entry of slot j for pattern base b (0..3)
__host__ __device__ constexpr uint32_t pat_slot(int k, int strand, int j, int b)
{
    const int i = slot_position(strand, j);
    const int pa = stage_a_slots(k);
    return j < pa ? (uint32_t)(4 * (i - score_pos_base(k, strand)) + b) * SCORE_THREADS * 4u
                  : ((uint32_t)(4 * pa + 2 * (j - pa)) * SCORE_THREADS * 4u) | ((uint32_t)b << 16);
}
// inverse of pat_slot: position and pattern base of slot j
__device__ __forceinline__ void pat_decode(int k, int strand, int j, uint32_t e, int &i, int &b)
{
    i = slot_position(strand, j);
    b = j < stage_a_slots(k) ? (int)((e / (SCORE_THREADS * 4u)) & 3u) : (int)(e >> 16);
}
// "mismatch against pattern base b" plane from the two planes of a position (b warp-uniform)
__device__ __forceinline__ uint32_t mismatch_plane(uint32_t h, uint32_t l, uint32_t b)
{
    return (h ^ ((b & 2u) ? ~0u : 0u)) | (l ^ ((b & 1u) ? ~0u : 0u));
}

struct ScoreArgs {
    const uint32_t *planes[2];  // per strand: [n_blocks][48]
    const uint32_t *pos[2];     // per strand: [n_blocks][32]
    const unsigned long long *n_blocks_ptr;   // [2]: blocks claimed by k_extract for this chunk, per strand (device counters)
    uint64_t cap;               // capacity of each candidate store; a chunk that overflowed it is skipped (host redoes it)
    uint32_t ctas_per_strand;   // grid = 2 * ctas_per_strand, sized by capacity
    uint32_t n_pat;             // guides in this launch (<= PAT_CHUNK)
    uint32_t guide_base;        // index of guide 0 of this launch in the guide list
    const uint32_t *pat_global; // the same table as c_pat in global memory (read by the slow path only)
    vs_hit *hits;
    unsigned long long *n_hits;
    uint64_t hit_cap;
};

// (One atomicAdd per hit on purpose: a warp-aggregated append — prefix sum of the per-lane hit counts, one atomic per
// warp — measured SLOWER on B200 in every config, including the dense one (config 5 at 0.02 scale, 34.7 M hits per
// scan: 28.4 vs 26.6 ms), because the whole warp then walks the slow path; 1.3 G atomics/s on one counter is no limit.)
// Slow path (rare): for every lane that passed the threshold read its exact count out of the bit-sliced counter,
// apply R4 to last-window candidates (H over positions 11..22 must be <= floor(K/2), bidir_mapping.cpp:48-53) and
// append the hit.  Only last-window lanes re-read planes (they need the second-half count) — from the block's global
// copy, with the pattern decoded from the GLOBAL copy of the table so that nothing of the hot loop's uniform-register
// state stays live.
template <int K>
__device__ __forceinline__ void score_hits(const uint32_t *gsrc, const uint32_t *po, int strand, uint32_t le, const uint32_t (&cnt)[5],
                                           uint32_t lastm, const uint32_t *pos, uint32_t info, vs_hit *hits,
                                           unsigned long long *n_hits, uint64_t hit_cap)
{
    while (le != 0) {
        const int c = __ffs(le) - 1;
        le &= le - 1;
        const uint32_t mm = ((cnt[0] >> c) & 1u) | (((cnt[1] >> c) & 1u) << 1) | (((cnt[2] >> c) & 1u) << 2) |
                            (((cnt[3] >> c) & 1u) << 3) | (((cnt[4] >> c) & 1u) << 4);
        if ((lastm >> c) & 1u) {                                             // R4: last window of its contig
            uint32_t h2 = 0;
#pragma unroll 1
            for (int j = 0; j < VS_GLEN; ++j) {
                int i, b;
                pat_decode(K, strand, j, __ldg(po + j), i, b);
                if (i >= 11) {
                    const uint32_t code = (((__ldg(gsrc + i * BLK_GROUP) >> c) & 1u) << 1) | ((__ldg(gsrc + (VS_GLEN + i) * BLK_GROUP) >> c) & 1u);
                    h2 += code != (uint32_t)b;
                }
            }
            if (h2 > (uint32_t)(K / 2)) continue;
        }
        const unsigned long long idx = atomicAdd(n_hits, 1ull);
        if (idx < hit_cap) {
            vs_hit hrec;
            hrec.pos = pos[c];
            hrec.info = info | mm;
            hits[idx] = hrec;
        }
    }
}

template <int K>
__global__ void __launch_bounds__(SCORE_THREADS, score_min_blocks(K))
k_score(ScoreArgs a)
{
#ifndef VS_HOST_UNIT_TEST       // (the host emulation declares vs::sm itself)
    extern __shared__ uint32_t sm[];     // [score_smem_planes(K)][SCORE_THREADS]
#endif
    constexpr int PA = stage_a_slots(K), PB = VS_GLEN - PA;
    const int tid = threadIdx.x;
    const uint32_t strand = blockIdx.x >= a.ctas_per_strand;            // forward CTAs first, then reverse
    const uint32_t cta = blockIdx.x - strand * a.ctas_per_strand;
    const uint64_t n_blocks = strand ? a.n_blocks_ptr[1] : a.n_blocks_ptr[0];
    if (max(a.n_blocks_ptr[0], a.n_blocks_ptr[1]) > a.cap || (uint64_t)cta * SCORE_THREADS >= n_blocks) return;
    const uint64_t blk = (uint64_t)cta * SCORE_THREADS + tid;
    const bool live = blk < n_blocks;
    // threads past the last block score block 0 with every lane invalid
    const uint32_t *gsrc = (strand ? a.planes[1] : a.planes[0]) + plane_index(live ? blk : 0, 0);      // word w at gsrc[w * BLK_GROUP]
    uint32_t *my = sm + tid;
    uint32_t lastm = 0;
    {
        uint32_t v[BLK_WORDS];
#pragma unroll
        for (int i = 0; i < BLK_WORDS; ++i) v[i] = __ldg(gsrc + i * BLK_GROUP);       // coalesced: 128 B per warp and word
        lastm = live ? v[BLK_LAST] : 0u;
        const uint32_t inv = live ? ~v[BLK_VALID] : ~0u;       // lanes past the end of a partial block mismatch everywhere
        auto expand = [&](auto base_c) {
            constexpr int BASE = decltype(base_c)::value;
#pragma unroll
            for (int i = 0; i < PA; ++i) {
                const uint32_t h = v[BASE + i], l = v[VS_GLEN + BASE + i];
                my[(4 * i + 0) * SCORE_THREADS] = (h | l) | inv;      // mismatch if the pattern base is A (00)
                my[(4 * i + 1) * SCORE_THREADS] = (h | ~l) | inv;     // C (01)
                my[(4 * i + 2) * SCORE_THREADS] = (~h | l) | inv;     // G (10)
                my[(4 * i + 3) * SCORE_THREADS] = (~h | ~l) | inv;    // T (11)
            }
            // raw planes of the stage-B slots (an invalid lane already mismatches in all PA > K stage-A slots)
#pragma unroll
            for (int i = 0; i < PB; ++i) {
                constexpr int S = BASE ? 1 : 0;
                const int p = slot_position(S, PA + i);
                my[(4 * PA + 2 * i) * SCORE_THREADS] = v[p];
                my[(4 * PA + 2 * i + 1) * SCORE_THREADS] = v[VS_GLEN + p];
            }
        };
        if constexpr (score_pos_base(K, 1) != 0) {
            if (strand) expand(std::integral_constant<int, score_pos_base(K, 1)>{});
            else expand(std::integral_constant<int, 0>{});
        } else {
            expand(std::integral_constant<int, 0>{});
        }
    }
    // each thread reads back only what it wrote: no barrier needed
    const char *myb = reinterpret_cast<const char *>(my);
    const uint32_t *pat0 = c_pat + strand * (PAT_CHUNK * PAT_STRIDE);
    const uint32_t zero5[5] = {0u, 0u, 0u, 0u, 0u};

    // Software pipeline over the guides: the stage-A planes of guide g+1 are loaded (LDS) before guide g is counted,
    // so the shared-memory latency and the adder tree of consecutive guides overlap inside one warp.
    // (Tried and measured slower on B200: sorting the guides by their bases at slots 0/1 and keeping those two planes in
    // registers per group — as a nested loop ptxas 12.9 drops the uniform-register slot offsets, as a flagged reload the
    // extra uniform instructions cost more than the two shared-memory wavefronts they save; three guides in flight
    // instead of two: +1 %.)
    auto load_a = [&](uint32_t g, uint32_t (&m)[PA]) {
        const uint32_t *po = pat0 + g * PAT_STRIDE;
#pragma unroll
        for (int i = 0; i < PA; ++i) m[i] = *reinterpret_cast<const uint32_t *>(myb + po[i]);
    };
    auto finish = [&](uint32_t g, const uint32_t (&ma)[PA]) {
        const uint32_t *po = pat0 + g * PAT_STRIDE;
        uint32_t ca[5];
        popcount_planes<PA, false>(ma, zero5, ca);
        uint32_t le = le_k<K>(ca);
        // warp-uniform early out: if no lane of the warp can still be within K, the remaining slots are never scored.
        // (A divergent `if (le)` here makes ptxas 12.9 crash on the uniform-register loads of stage B.)
        if (PB == 0 ? (le != 0) : __any_sync(0xffffffffu, le != 0)) {
            if constexpr (PB > 0) {
                // stage B: the remaining slots, from their raw planes, added onto stage A's count
                uint32_t mb[PB], cb[5];
#pragma unroll
                for (int i = 0; i < PB; ++i) {
                    const uint32_t e = po[PA + i];
                    const char *q = myb + (e & 0xFFFFu);
                    mb[i] = mismatch_plane(*reinterpret_cast<const uint32_t *>(q), *reinterpret_cast<const uint32_t *>(q + SCORE_THREADS * 4), e >> 16);
                }
                popcount_planes<PB, true>(mb, ca, cb);
                le = le_k<K>(cb);
#pragma unroll
                for (int w = 0; w < 5; ++w) ca[w] = cb[w];
            }
            if (le != 0)
                score_hits<K>(gsrc, a.pat_global + (strand * PAT_CHUNK + g) * PAT_STRIDE, (int)strand, le, ca, lastm,
                              (strand ? a.pos[1] : a.pos[0]) + blk * 32, ((a.guide_base + g) << 8) | (strand << 7),
                              a.hits, a.n_hits, a.hit_cap);
        }
    };
    const uint32_t n_pat = a.n_pat;
    if (n_pat == 0) return;
    uint32_t m0[PA], m1[PA];
    load_a(0, m0);
    uint32_t g = 0;
    for (; g + 1 < n_pat; g += 2) {
        load_a(g + 1, m1);
        finish(g, m0);
        if (g + 2 < n_pat) load_a(g + 2, m0);
        finish(g + 1, m1);
    }
    if (g < n_pat) finish(g, m0);
}

// ------------------------------------------------------------------------------------------------
// Microbenchmarks for the roofline denominators (alu-pipe LOP3 issue rate, shared-memory LDS rate).
#ifndef VS_HOST_UNIT_TEST
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

#endif  // VS_HOST_UNIT_TEST

}  // namespace vs
