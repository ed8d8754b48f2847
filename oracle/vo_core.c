/*
 * oracle/vo_core.c — CPU ORACLE (test infrastructure, NOT product code).  See vo_core.h.
 *
 * Restates VARSCOT_pipeline/read_mapping/bidir_mapping.cpp.  Every function cites the
 * reference lines it follows.  Two search modes produce the same records:
 *
 *   VO_MODE_LITERAL  keeps the reference's control flow: per (guide, pass) search the first
 *                    half (11-mer) and the second half (12-mer) with <= K = floor(k/2)
 *                    substitutions, and run the verify delegate on every seed occurrence
 *                    (bounds, dedup, PAM, count, record).  SeqAn's find<0,K>(.., HammingDistance())
 *                    over the bidirectional FM index is replaced by an exhaustive enumeration
 *                    of all seed occurrences with <= K substitutions inside one contig
 *                    (assumptions A1, A2 of SURVEY.md section 8c): same occurrence SET,
 *                    discovery ORDER is ours (matters only for uint16 key collisions).
 *   VO_MODE_SCAN     evaluates the distilled rules R1-R4 on every window (2-bit packed
 *                    rolling window, XOR + popcount, OpenMP).  Used for big parity tests and
 *                    as the timed CPU baseline.
 *
 * PARITY UNPINNED (no SeqAn, no golden SAM in the checkout) — see vo_core.h.
 */
#include "vo_core.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ alphabets */

/* SeqAn Dna5 conversion used for the text (common.h / bidir_index.cpp:12: Dna5String):
 * A,C,G,T case-insensitive, U as T, everything else N. [R6] */
uint8_t vo_text_code(char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;
    }
}

/* SeqAn Dna conversion used for the guides (bidir_mapping.cpp:256 StringSet<DnaString>;
 * help text :194 "everything else than ACGT will be converted to A"). [R5] */
uint8_t vo_guide_code(char c)
{
    uint8_t t = vo_text_code(c);
    return t > 3 ? 0 : t;
}

int vo_num_procs(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ small helpers */

typedef struct {
    uint32_t guide;
    uint32_t contig;
    uint32_t pos;
    uint8_t  strand;   /* 0 forward pass, 1 reverse pass */
    uint8_t  mm;
    uint32_t disc;     /* discovery index inside the pass (literal mode) */
} hit_t;

typedef struct { hit_t *v; uint64_t n, cap; } hitvec;

static int hv_push(hitvec *h, hit_t x)
{
    if (h->n == h->cap) {
        uint64_t nc = h->cap ? h->cap * 2 : 1024;
        hit_t *nv = (hit_t *)realloc(h->v, nc * sizeof(hit_t));
        if (!nv) return -1;
        h->v = nv; h->cap = nc;
    }
    h->v[h->n++] = x;
    return 0;
}

/* isValidPAM (bidir_mapping.cpp:21-29) with the lists of :240-247.
 * pam_ok[x*5+y] over Dna5 codes; forward list {GG, GA} + XY, reverse list {CC, TC} + revcomp(XY). */
static void build_pam_tables(int extra_pam, uint8_t fwd[25], uint8_t rev[25])
{
    memset(fwd, 0, 25); memset(rev, 0, 25);
    fwd[2 * 5 + 2] = 1;            /* GG */
    fwd[2 * 5 + 0] = 1;            /* GA */
    rev[1 * 5 + 1] = 1;            /* CC */
    rev[3 * 5 + 1] = 1;            /* TC */
    if (extra_pam >= 0) {
        int x = extra_pam / 4, y = extra_pam % 4;
        fwd[x * 5 + y] = 1;
        /* reverseComplement(additionalPAM) (:245): (3-y, 3-x) */
        rev[(3 - y) * 5 + (3 - x)] = 1;
    }
}

static void make_pattern(const uint8_t *guide, int reverse, uint8_t *P)
{
    if (!reverse) { memcpy(P, guide, VO_GLEN); return; }
    for (int i = 0; i < VO_GLEN; ++i) P[i] = (uint8_t)(3 - guide[VO_GLEN - 1 - i]);  /* reverseComplement(read), :293 */
}

/* ------------------------------------------------------------------ literal mode */

/* open-addressing set of 64-bit keys for records.find(TOccType(..)) (bidir_mapping.cpp:64) */
typedef struct { uint64_t *k; uint64_t cap, n; } keyset;
#define KS_EMPTY 0xFFFFFFFFFFFFFFFFull
static int ks_init(keyset *s) { s->cap = 1024; s->n = 0; s->k = (uint64_t *)malloc(s->cap * 8); if (!s->k) return -1; memset(s->k, 0xFF, s->cap * 8); return 0; }
static uint64_t ks_hash(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; return x; }
static int ks_has(const keyset *s, uint64_t key)
{
    uint64_t i = ks_hash(key) & (s->cap - 1);
    while (s->k[i] != KS_EMPTY) { if (s->k[i] == key) return 1; i = (i + 1) & (s->cap - 1); }
    return 0;
}
static int ks_add(keyset *s, uint64_t key)
{
    if ((s->n + 1) * 2 > s->cap) {
        uint64_t oc = s->cap; uint64_t *ok = s->k;
        s->cap *= 2; s->k = (uint64_t *)malloc(s->cap * 8); if (!s->k) return -1;
        memset(s->k, 0xFF, s->cap * 8); s->n = 0;
        for (uint64_t j = 0; j < oc; ++j) if (ok[j] != KS_EMPTY) ks_add(s, ok[j]);
        free(ok);
    }
    uint64_t i = ks_hash(key) & (s->cap - 1);
    while (s->k[i] != KS_EMPTY) i = (i + 1) & (s->cap - 1);
    s->k[i] = key; s->n++;
    return 0;
}

static uint64_t make_key(uint32_t contig, uint32_t pos, int key_mode)
{
    /* TOccType(chromosomeId, posInChromosome) narrows the id to uint16_t (bidir_mapping.cpp:13,64,125) */
    if (key_mode == VO_KEY_REF16) return ((uint64_t)(contig & 0xFFFFu) << 32) | pos;
    /* wide: unique per (contig,pos); we only need set membership here */
    return ((uint64_t)contig << 32) | pos;
}

/* The verify delegate, bidir_mapping.cpp:34-127, for one seed occurrence (contig c, position q). */
static int delegate(const uint8_t *text, const uint64_t *off, uint32_t c, uint64_t q,
                    const uint8_t *P, int reverse, int first_half, int k,
                    const uint8_t *fwd_pam, const uint8_t *rev_pam, int key_mode,
                    keyset *present, keyset *seen_wide, uint64_t *collisions,
                    uint32_t guide, hitvec *out)
{
    const uint8_t *chrom = text + off[c];
    uint64_t L = off[c + 1] - off[c];
    int64_t pos = (int64_t)q;
    if (first_half) {
        /* :48-53  if (length(chromosome) <= posInChromosome + length(fullRead)) continue;  (note <=) */
        if (L <= (uint64_t)pos + VO_GLEN) return 0;
    } else {
        /* :54-62  extend to the left by length(fullRead) - length(partialRead) = 11 */
        if (pos - (VO_GLEN - (VO_GLEN - VO_HALF1)) < 0) return 0;
        pos -= VO_HALF1;
    }
    /* :64-65 dedup on the (uint16 id, pos) map key */
    uint64_t key = make_key(c, (uint32_t)pos, key_mode);
    if (ks_has(present, key)) {
        if (key_mode == VO_KEY_REF16) {
            /* would this have been a distinct, acceptable hit?  counted for the report only */
            uint64_t wk = ((uint64_t)c << 32) | (uint32_t)pos;
            if (!ks_has(seen_wide, wk)) {
                /* evaluate acceptance to count a real collision */
                const uint8_t *W = chrom + pos;
                int ok = reverse ? rev_pam[W[0] * 5 + W[1]] : fwd_pam[W[21] * 5 + W[22]];
                if (ok) {
                    unsigned mm = 0;
                    for (unsigned i = 0; i < VO_GLEN && mm <= (unsigned)k; ++i) {
                        if (W[i] == 4) mm += k + 1;
                        mm += (P[i] != W[i]);
                    }
                    if (mm <= (unsigned)k) { (*collisions)++; ks_add(seen_wide, wk); }
                }
            }
        }
        return 0;
    }
    const uint8_t *W = chrom + pos;            /* :67 infixWithLength(chromosome, pos, 23) */
    /* :70-76 PAM on the genome window */
    if (!reverse && !fwd_pam[W[21] * 5 + W[22]]) return 0;
    if (reverse && !rev_pam[W[0] * 5 + W[1]]) return 0;
    /* :79-86 mismatch count over all 23 positions, N makes the window invalid */
    unsigned mm = 0;
    for (unsigned i = 0; i < VO_GLEN && mm <= (unsigned)k; ++i) {
        if (W[i] == 4) mm += k + 1;
        mm += (P[i] != W[i]);
    }
    if (mm > (unsigned)k) return 0;
    /* :88-125 record + insert */
    hit_t h; h.guide = guide; h.contig = c; h.pos = (uint32_t)pos; h.strand = (uint8_t)reverse; h.mm = (uint8_t)mm;
    h.disc = (uint32_t)out->n;
    if (hv_push(out, h)) return -1;
    if (ks_add(present, key)) return -1;
    if (key_mode == VO_KEY_REF16) ks_add(seen_wide, ((uint64_t)c << 32) | (uint32_t)pos);
    return 0;
}

/* searchAndVerify (bidir_mapping.cpp:31-148): enumerate every occurrence of `seed` with <= K
 * substitutions (find<0,K>(.., HammingDistance()), :129-146) inside one contig and call the delegate. */
static int search_and_verify(const uint8_t *text, const uint64_t *off, uint32_t n_contigs,
                             const uint8_t *P, const uint8_t *seed, int seed_len,
                             int reverse, int first_half, int k,
                             const uint8_t *fwd_pam, const uint8_t *rev_pam, int key_mode,
                             keyset *present, keyset *seen_wide, uint64_t *collisions,
                             uint32_t guide, hitvec *out)
{
    int K = k / 2;   /* :129-146: k 0,1->0; 2,3->1; 4,5->2; 6,7->3; 8->4 */
    for (uint32_t c = 0; c < n_contigs; ++c) {
        const uint8_t *chrom = text + off[c];
        uint64_t L = off[c + 1] - off[c];
        if (L < (uint64_t)seed_len) continue;
        for (uint64_t q = 0; q + seed_len <= L; ++q) {
            int e = 0;
            for (int i = 0; i < seed_len && e <= K; ++i) e += (seed[i] != chrom[q + i]);   /* text N never equals a Dna needle char */
            if (e > K) continue;
            if (delegate(text, off, c, q, P, reverse, first_half, k, fwd_pam, rev_pam, key_mode,
                         present, seen_wide, collisions, guide, out)) return -1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ scan mode */

#define PMASK 0x155555555555ull   /* bit 2i for i < 23 */

static uint64_t pack23(const uint8_t *P)
{
    uint64_t w = 0;
    for (int i = 0; i < VO_GLEN; ++i) w |= (uint64_t)(P[i] & 3) << (2 * i);
    return w;
}

/* Rules on one window: R1 fit (caller), R2 PAM, R3 no N + H <= k,
 * R4 last window of a contig additionally needs H2 <= K because the first-half seed path
 * rejects L_c <= p+23 (bidir_mapping.cpp:51) and only the second-half seed can reach it. */
static void scan_range(const uint8_t *text, const uint64_t *off, uint32_t n_contigs,
                       uint64_t ga, uint64_t gb, /* global start-position range [ga, gb) */
                       const uint64_t *Pf, const uint64_t *Pr, uint32_t n_guides,
                       int k, const uint8_t *fwd_pam, const uint8_t *rev_pam,
                       hitvec *out, uint64_t *count)
{
    int K = k / 2;
    /* contig containing ga */
    uint32_t lo = 0, hi = n_contigs;
    while (hi - lo > 1) { uint32_t mid = lo + (hi - lo) / 2; if (off[mid] <= ga) lo = mid; else hi = mid; }
    uint64_t cnt = 0;
    for (uint32_t c = lo; c < n_contigs && off[c] < gb; ++c) {
        uint64_t cs = off[c], ce = off[c + 1];
        if (ce - cs < VO_GLEN) continue;
        uint64_t s0 = cs > ga ? cs : ga;                 /* first start position (global) */
        uint64_t s1 = (ce - VO_GLEN + 1) < gb ? (ce - VO_GLEN + 1) : gb;   /* one past last start */
        if (s0 >= s1) continue;
        /* preload 22 bases */
        uint64_t w = 0; int run = 0;
        for (uint64_t j = s0; j < s0 + VO_GLEN - 1; ++j) {
            uint8_t b = text[j];
            run = (b > 3) ? 0 : run + 1;
            w = (w >> 2) | ((uint64_t)(b & 3) << 44);
        }
        for (uint64_t p = s0; p < s1; ++p) {
            uint8_t b = text[p + VO_GLEN - 1];
            run = (b > 3) ? 0 : run + 1;
            w = (w >> 2) | ((uint64_t)(b & 3) << 44);
            if (run < VO_GLEN) continue;                  /* R3: an N anywhere rejects the window */
            int f_ok = fwd_pam[((w >> 42) & 3) * 5 + ((w >> 44) & 3)];
            int r_ok = rev_pam[(w & 3) * 5 + ((w >> 2) & 3)];
            if (!f_ok && !r_ok) continue;
            int last = (p + VO_GLEN == ce);
            for (int s = 0; s < 2; ++s) {
                if (!(s ? r_ok : f_ok)) continue;
                const uint64_t *PP = s ? Pr : Pf;
                for (uint32_t g = 0; g < n_guides; ++g) {
                    uint64_t x = w ^ PP[g];
                    uint64_t y = (x | (x >> 1)) & PMASK;
                    int mm = __builtin_popcountll(y);
                    if (mm > k) continue;
                    if (last && __builtin_popcountll(y >> (2 * VO_HALF1)) > K) continue;   /* R4 */
                    cnt++;
                    if (out) {
                        hit_t h; h.guide = g; h.contig = c; h.pos = (uint32_t)(p - cs); h.strand = (uint8_t)s; h.mm = (uint8_t)mm; h.disc = 0;
                        hv_push(out, h);
                    }
                }
            }
        }
    }
    if (count) *count += cnt;
}

#define SCAN_CHUNK (1u << 20)

static int scan_all(const uint8_t *text, const uint64_t *off, uint32_t n_contigs,
                    const uint8_t *guides, uint32_t n_guides, int k, int extra_pam,
                    int n_threads, hitvec *out, uint64_t *count)
{
    uint8_t fwd_pam[25], rev_pam[25];
    build_pam_tables(extra_pam, fwd_pam, rev_pam);
    uint64_t *Pf = (uint64_t *)malloc((size_t)(n_guides ? n_guides : 1) * 8);
    uint64_t *Pr = (uint64_t *)malloc((size_t)(n_guides ? n_guides : 1) * 8);
    if (!Pf || !Pr) { free(Pf); free(Pr); return -1; }
    for (uint32_t g = 0; g < n_guides; ++g) {
        uint8_t P[VO_GLEN];
        make_pattern(guides + (size_t)g * VO_GLEN, 0, P); Pf[g] = pack23(P);
        make_pattern(guides + (size_t)g * VO_GLEN, 1, P); Pr[g] = pack23(P);
    }
    uint64_t total = n_contigs ? off[n_contigs] : 0;
    int64_t n_chunks = (int64_t)((total + SCAN_CHUNK - 1) / SCAN_CHUNK);
    if (n_threads < 1) n_threads = 1;
    hitvec *tv = (hitvec *)calloc((size_t)n_threads, sizeof(hitvec));
    uint64_t *tc = (uint64_t *)calloc((size_t)n_threads, sizeof(uint64_t));
    if (!tv || !tc) { free(Pf); free(Pr); free(tv); free(tc); return -1; }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
    for (int64_t ch = 0; ch < n_chunks; ++ch) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        uint64_t ga = (uint64_t)ch * SCAN_CHUNK, gb = ga + SCAN_CHUNK;
        if (gb > total) gb = total;
        scan_range(text, off, n_contigs, ga, gb, Pf, Pr, n_guides, k, fwd_pam, rev_pam,
                   out ? &tv[tid] : NULL, &tc[tid]);
    }
    uint64_t cnt = 0;
    for (int t = 0; t < n_threads; ++t) {
        cnt += tc[t];
        if (out) for (uint64_t i = 0; i < tv[t].n; ++i) hv_push(out, tv[t].v[i]);
        free(tv[t].v);
    }
    if (count) *count = cnt;
    free(tv); free(tc); free(Pf); free(Pr);
    return 0;
}

uint64_t vo_scan_count(const uint8_t *text, const uint64_t *off, uint32_t n_contigs,
                       const uint8_t *guides, uint32_t n_guides, int k, int extra_pam, int n_threads)
{
    uint64_t c = 0;
    if (scan_all(text, off, n_contigs, guides, n_guides, k, extra_pam, n_threads, NULL, &c)) return (uint64_t)-1;
    return c;
}

/* ------------------------------------------------------------------ ordering, flags, MD */

static int g_key_mode_for_sort;   /* qsort has no context argument; vo_map is not re-entrant across key modes */

/* std::map<TOccType, BamRecord> iteration order (bidir_mapping.cpp:154,170): (uint16 id, pos);
 * wide mode appends the high id bits so that nothing collides. */
static int cmp_pass(const void *a, const void *b)
{
    const hit_t *x = (const hit_t *)a, *y = (const hit_t *)b;
    if (x->guide != y->guide) return x->guide < y->guide ? -1 : 1;
    if (x->strand != y->strand) return x->strand < y->strand ? -1 : 1;
    uint32_t xa = x->contig & 0xFFFFu, ya = y->contig & 0xFFFFu;
    if (xa != ya) return xa < ya ? -1 : 1;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    if (g_key_mode_for_sort == VO_KEY_WIDE) {
        uint32_t xh = x->contig >> 16, yh = y->contig >> 16;
        if (xh != yh) return xh < yh ? -1 : 1;
    }
    if (x->disc != y->disc) return x->disc < y->disc ? -1 : 1;
    return 0;
}

/* getMDString(md, row(align,0)=mappedRegion, row(align,1)=fullRead) (bidir_mapping.cpp:113-119):
 * SeqAn 2.4 bam_io walks both rows; a match run is flushed as a number only when a run of
 * matches ends (or at the end), mismatching columns append the GENOME character. */
void vo_md_string(const uint8_t *window, const uint8_t *pattern, int md_style, char *out)
{
    static const char L[5] = { 'A', 'C', 'G', 'T', 'N' };
    int n = 0, run = 0;
    if (md_style == VO_MD_SEQAN) {
        char last = ' ';
        for (int i = 0; i < VO_GLEN; ++i) {
            char op = (window[i] == pattern[i]) ? 'M' : 'R';
            if (last != op) {
                if (last == 'M') n += sprintf(out + n, "%d", run);
                run = 0; last = op;
            }
            if (op != 'M') out[n++] = L[window[i] > 4 ? 4 : window[i]];
            ++run;
        }
        if (last == 'M') n += sprintf(out + n, "%d", run);
    } else {
        for (int i = 0; i < VO_GLEN; ++i) {
            if (window[i] == pattern[i]) { ++run; continue; }
            n += sprintf(out + n, "%d", run);
            out[n++] = L[window[i] > 4 ? 4 : window[i]];
            run = 0;
        }
        n += sprintf(out + n, "%d", run);
    }
    out[n] = 0;
}

static int res_push(vo_result *r, const hit_t *h, uint16_t flag, const uint8_t *text, const uint64_t *off,
                    const uint8_t *guides, int md_style)
{
    if (r->n == r->cap) {
        uint64_t nc = r->cap ? r->cap * 2 : 256;
        vo_record *nv = (vo_record *)realloc(r->rec, nc * sizeof(vo_record));
        if (!nv) return -1;
        r->rec = nv; r->cap = nc;
    }
    vo_record *o = &r->rec[r->n++];
    memset(o, 0, sizeof(*o));
    o->guide = h->guide; o->contig = h->contig; o->pos = h->pos; o->mm = h->mm;
    o->flag = (uint16_t)(flag | (h->strand ? 16 : 0));          /* BAM_FLAG_RC, :97-98 */
    uint8_t P[VO_GLEN];
    make_pattern(guides + (size_t)h->guide * VO_GLEN, h->strand, P);
    vo_md_string(text + off[h->contig] + h->pos, P, md_style, o->md);
    return 0;
}

/* searchAndVerifyEntireRead's write-out (bidir_mapping.cpp:164-187): walk the map in key order
 * keeping a running best; every record that is not strictly better than the best is written at
 * once with BAM_FLAG_SECONDARY; a strictly better one pushes the old best out (secondary) and
 * becomes the best; the final best is written last without the secondary bit. */
static int emit_pass(vo_result *r, const hit_t *v, uint64_t n, const uint8_t *text, const uint64_t *off,
                     const uint8_t *guides, int md_style)
{
    if (n == 0) return 0;
    uint64_t best = 0;
    for (uint64_t i = 1; i < n; ++i) {
        if (v[i].mm >= v[best].mm) {
            if (res_push(r, &v[i], 256, text, off, guides, md_style)) return -1;
        } else {
            if (res_push(r, &v[best], 256, text, off, guides, md_style)) return -1;
            best = i;
        }
    }
    return res_push(r, &v[best], 0, text, off, guides, md_style);
}

/* ------------------------------------------------------------------ entry point */

int vo_map(const uint8_t *text, const uint64_t *off, uint32_t n_contigs,
           const uint8_t *guides, uint32_t n_guides,
           int k, int extra_pam, int mode, int key_mode, int md_style,
           int n_threads, vo_result *out)
{
    if (!out) return 1;
    memset(out, 0, sizeof(*out));
    if (k < 0 || k > 8) return 2;                         /* bidir_mapping.cpp:234-238 */
    if (extra_pam > 15) return 2;
    uint8_t fwd_pam[25], rev_pam[25];
    build_pam_tables(extra_pam, fwd_pam, rev_pam);
    hitvec all = { 0, 0, 0 };

    if (mode == VO_MODE_LITERAL) {
        /* main loop, bidir_mapping.cpp:285-295: per guide, forward pass then reverse-complement pass */
        for (uint32_t g = 0; g < n_guides; ++g) {
            for (int reverse = 0; reverse < 2; ++reverse) {
                uint8_t P[VO_GLEN];
                make_pattern(guides + (size_t)g * VO_GLEN, reverse, P);
                keyset present, seen_wide;
                if (ks_init(&present) || ks_init(&seen_wide)) return 3;
                hitvec pass = { 0, 0, 0 };
                /* :157-158 first half = read[0, 11), extend right */
                int rc = search_and_verify(text, off, n_contigs, P, P, VO_HALF1, reverse, 1, k, fwd_pam, rev_pam,
                                           key_mode, &present, &seen_wide, &out->key16_collisions, g, &pass);
                /* :161-162 second half = read[11, 23), extend left */
                if (!rc) rc = search_and_verify(text, off, n_contigs, P, P + VO_HALF1, VO_GLEN - VO_HALF1, reverse, 0, k,
                                                fwd_pam, rev_pam, key_mode, &present, &seen_wide,
                                                &out->key16_collisions, g, &pass);
                free(present.k); free(seen_wide.k);
                if (rc) { free(pass.v); free(all.v); return 3; }
                for (uint64_t i = 0; i < pass.n; ++i) hv_push(&all, pass.v[i]);
                free(pass.v);
            }
        }
    } else {
        if (scan_all(text, off, n_contigs, guides, n_guides, k, extra_pam, n_threads, &all, NULL)) { free(all.v); return 3; }
    }

    g_key_mode_for_sort = key_mode;
    if (all.n) qsort(all.v, all.n, sizeof(hit_t), cmp_pass);

    /* scan mode + REF16: emulate "first found wins" by keeping the lowest full id (discovery order of the
     * reference is unknowable without SeqAn); count what was dropped. */
    if (mode == VO_MODE_SCAN && key_mode == VO_KEY_REF16 && all.n) {
        uint64_t w = 0;
        for (uint64_t i = 0; i < all.n; ++i) {
            if (w > 0) {
                hit_t *p = &all.v[w - 1], *q = &all.v[i];
                if (p->guide == q->guide && p->strand == q->strand && (p->contig & 0xFFFFu) == (q->contig & 0xFFFFu) && p->pos == q->pos) {
                    out->key16_collisions++;
                    if (q->contig < p->contig) *p = *q;
                    continue;
                }
            }
            all.v[w++] = all.v[i];
        }
        all.n = w;
    }

    uint64_t i = 0;
    while (i < all.n) {
        uint64_t j = i + 1;
        while (j < all.n && all.v[j].guide == all.v[i].guide && all.v[j].strand == all.v[i].strand) ++j;
        if (emit_pass(out, all.v + i, j - i, text, off, guides, md_style)) { free(all.v); return 3; }
        i = j;
    }
    free(all.v);
    return 0;
}

void vo_result_free(vo_result *r)
{
    if (!r) return;
    free(r->rec);
    memset(r, 0, sizeof(*r));
}

/* write(buffer, record, bamIOContext, Sam()) for the record built at bidir_mapping.cpp:88-123 [R9]:
 * QNAME FLAG RNAME POS(1-based) 255 23M * 0 0 SEQ(original guide, :106-108) I*23 NM:i:<mm> MD:Z:<md> */
int vo_format_sam(const vo_record *r, const char *qname, const char *rname,
                  const uint8_t *guide_codes, char *buf, size_t buflen)
{
    static const char L[4] = { 'A', 'C', 'G', 'T' };
    char seq[VO_GLEN + 1];
    for (int i = 0; i < VO_GLEN; ++i) seq[i] = L[guide_codes[i] & 3];
    seq[VO_GLEN] = 0;
    return snprintf(buf, buflen, "%s\t%u\t%s\t%u\t255\t23M\t*\t0\t0\t%s\tIIIIIIIIIIIIIIIIIIIIIII\tNM:i:%u\tMD:Z:%s\n",
                    qname, (unsigned)r->flag, rname, r->pos + 1u, seq, (unsigned)r->mm, r->md);
}
