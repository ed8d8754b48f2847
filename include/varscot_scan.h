/*
 * include/varscot_scan.h — C ABI of the B200-native off-target scan ("read_mapping" hot path).
 *
 * The reference has no in-process API for this path: its boundary is two executables,
 * `bidir_index` and `bidir_mapping` (VARSCOT_pipeline/read_mapping/bidir_index.cpp:10-52,
 * bidir_mapping.cpp:190-312), called by the driver script (VARSCOT_pipeline/VARSCOT:296-314).
 * This header is the layer those two executables (and any FFI binding: ctypes, cgo, JNI) sit on.
 * Plain C types only; caller-owned host buffers; callee-owned device buffers; int error codes.
 *
 * Entry point                      replaces (reference file:line)
 * -------------------------------  -----------------------------------------------------------------
 * vs_packer_* / vs_pack_text       bidir_index.cpp:36-47  readRecords -> Dna5 StringSet -> indexCreate
 * vs_text_save / vs_text_load      bidir_index.cpp:47 save(index, path); bidir_mapping.cpp:268 open(index, path)
 * vs_ctx_create / vs_text_upload   bidir_mapping.cpp:268  index resident in RAM -> packed text resident in HBM
 * vs_scan                          bidir_mapping.cpp:285-295 omp-parallel loop over guides calling
 *                                  searchAndVerifyEntireRead -> searchAndVerify (:31-148, :150-188):
 *                                  seed search + verify delegate, for both strands
 * vs_resolve_hits                  bidir_mapping.cpp:154,164-187  std::map order + primary/secondary flags
 * vs_md_string / vs_format_sam     bidir_mapping.cpp:111-123 getMDString + tags; :177-187 write(.., Sam())
 * vs_bidir_index_main              bidir_index.cpp:10-52   main (argv contract)
 * vs_bidir_mapping_main            bidir_mapping.cpp:190-312 main (argv contract, stdout lines, exit codes)
 */
#ifndef VARSCOT_SCAN_H
#define VARSCOT_SCAN_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_GLEN 23            /* guide length incl. PAM (CIGAR 23M hard-coded at bidir_mapping.cpp:102) */
#define VS_MAX_MISMATCHES 8   /* bidir_mapping.cpp:234-238 */

enum {
    VS_OK = 0,
    VS_ERR_ARG = 1,
    VS_ERR_CUDA = 2,
    VS_ERR_NOMEM = 3,
    VS_ERR_OVERFLOW = 4,     /* caller's hit buffer too small; *n_hits holds the required count */
    VS_ERR_NODEVICE = 5,
    VS_ERR_IO = 6
};

/* ---- packed text ---------------------------------------------------------------------------
 * The text (all contigs concatenated, "ConcatDirect" as bidir_index.cpp:12) is stored bit-sliced,
 * 32 bases per word, base j of a word in bit j:
 *   hi, lo : the two bits of the Dna code (A=00, C=01, G=10, T=11)
 *   nm     : 1 where the base is N (anything outside ACGT/U, R6) or padding past the end
 *   em     : 1 where the base is the LAST base of a contig
 * n_words = ceil(n_bases/32); the array always carries ONE extra pad word (nm = all ones) so that
 * word[i+1] is readable for every owned word i.
 */
typedef struct { uint32_t hi, lo, nm, em; } vs_word;

typedef struct vs_packer vs_packer;
vs_packer *vs_packer_new(void);
void       vs_packer_free(vs_packer *p);
/* append raw sequence characters (no newlines needed to be stripped: whitespace is skipped) */
int        vs_packer_append(vs_packer *p, const char *chars, size_t n);
/* close the current contig (contigs of length 0 are kept: they occupy an id, SeqAn StringSet semantics) */
int        vs_packer_end_contig(vs_packer *p);
uint64_t   vs_packer_num_bases(const vs_packer *p);
uint32_t   vs_packer_num_contigs(const vs_packer *p);
uint64_t   vs_packer_num_words(const vs_packer *p);           /* without the pad word */
/* pointers stay valid until the packer is freed or appended to; words has num_words+1 entries, offsets n_contigs+1 */
const vs_word  *vs_packer_words(vs_packer *p);
const uint64_t *vs_packer_offsets(vs_packer *p);

/* one-shot: ASCII text of all contigs concatenated + n_contigs+1 offsets -> out_words[ceil(n/32)+1] */
int vs_pack_text(const char *ascii, uint64_t n_bases, const uint64_t *contig_off, uint32_t n_contigs, vs_word *out_words);

/* packed-text cache at the -I prefix (files <prefix>.vsidx) */
int vs_text_save(const char *prefix, const vs_word *words, uint64_t n_bases, const uint64_t *contig_off, uint32_t n_contigs);
/* loads into malloc'd buffers the caller frees with vs_free */
int vs_text_load(const char *prefix, vs_word **words, uint64_t *n_bases, uint64_t **contig_off, uint32_t *n_contigs);
void vs_free(void *p);

/* ---- device context ------------------------------------------------------------------------- */
typedef struct vs_ctx vs_ctx;

int  vs_device_count(void);                     /* >= 0, or -VS_ERR_* */
int  vs_ctx_create(int device, vs_ctx **ctx);   /* one context per device per host thread */
void vs_ctx_destroy(vs_ctx *ctx);
const char *vs_last_error(const vs_ctx *ctx);   /* ctx may be NULL: last error of the calling thread */

/* Upload words [0, n_words) of a shard; words[n_words] must be readable (next word of the text, or the
 * pad word).  Window START positions in [0, 32*n_words) are owned by this context; hit positions are
 * reported as global_base + local start.  pinned != 0 promises `words` is page-locked. */
int vs_text_upload(vs_ctx *ctx, const vs_word *words, uint64_t n_words, uint64_t global_base);

/* page-locked host memory for fast uploads / hit downloads */
void *vs_host_alloc(size_t bytes);
void  vs_host_free(void *p);

typedef struct {
    uint32_t pos;    /* global start position in the concatenated text */
    uint32_t info;   /* guide << 8 | strand << 7 | mm  (strand 1 = reverse pass, flag bit 16) */
} vs_hit;

typedef struct {
    float    count_ms, extract_ms, score_ms, total_ms;   /* CUDA-event times on the context's stream */
    uint64_t n_cand_fwd, n_cand_rev;                      /* PAM-valid, N-free windows per strand */
    uint64_t n_blocks_fwd, n_blocks_rev;                  /* 32-candidate blocks scored per strand */
    uint64_t n_hits;
    uint32_t launches;                                    /* kernels launched by this call */
    uint32_t score_launches;
} vs_scan_stats;

/* Scan the uploaded text for every window within k mismatches of each guide, both strands (rules R1-R4,
 * SURVEY.md section 8a).  guides: n_guides x 23 Dna codes (0..3).  extra_pam: -1 or 4*x+y for -P XY.
 * Hits arrive UNORDERED.  If more than out_cap hits exist returns VS_ERR_OVERFLOW with *n_hits = needed;
 * the hits stay on the device and vs_scan_fetch() retrieves them without rescanning. */
int vs_scan(vs_ctx *ctx, const uint8_t *guides, uint32_t n_guides, int k, int extra_pam,
            vs_hit *out, uint64_t out_cap, uint64_t *n_hits, vs_scan_stats *stats);
int vs_scan_fetch(vs_ctx *ctx, vs_hit *out, uint64_t out_cap, uint64_t *n_hits);

/* Convenience used by the executables and bindings: shard a packed text by word ranges over the given
 * devices (devices == NULL or n_devices == 0 -> device 0; one host thread + one context per device, no
 * collective), scan, and return all hits unordered in a malloc'd array (release with vs_free). */
int vs_map_packed(const vs_word *words, uint64_t n_bases, const uint8_t *guides, uint32_t n_guides,
                  int k, int extra_pam, const int *devices, int n_devices,
                  vs_hit **hits, uint64_t *n_hits, vs_scan_stats *stats);

/* Shard plan used by vs_map_packed and by one-process-per-GPU callers: out[0..n] are word indices, shard i owns the
 * window starts of words [out[i], out[i+1]) and reads one halo word after them.  Tile-aligned, balanced. */
int vs_shard_bounds(uint64_t n_words, int n_shards, uint64_t *out);

/* ---- host-side resolution ------------------------------------------------------------------- */
typedef struct {
    uint32_t guide;
    uint32_t contig;   /* full 32-bit id */
    uint32_t pos;      /* 0-based inside the contig */
    uint16_t flag;     /* 0 / 16 / 256 / 272 */
    uint8_t  mm;
    uint8_t  pad;
} vs_record;

/* Sort hits into the reference's emission order (guide; forward pass then reverse pass; std::map key order
 * (contig & 0xFFFF, pos) extended by contig >> 16) and apply the running-best primary/secondary rule.
 * out must hold n records.  *key16_collisions counts records the reference's uint16 key would have merged. */
int vs_resolve_hits(const vs_hit *hits, uint64_t n, const uint64_t *contig_off, uint32_t n_contigs,
                    vs_record *out, uint64_t *key16_collisions);

#define VS_MD_SEQAN 0
#define VS_MD_SAMTOOLS 1
/* MD:Z value for the window at global position gpos against guide (codes) on the given strand. out >= 64 bytes. */
int vs_md_string(const vs_word *words, uint64_t gpos, const uint8_t *guide, int strand, int md_style, char *out);
/* one 13-column SAM line (R9) incl. '\n'; returns the length written (excluding NUL), or -1 if buflen is too small */
int vs_format_sam(const vs_record *r, const char *qname, const char *rname, const uint8_t *guide,
                  const char *md, char *buf, size_t buflen);

/* ---- the two executables as library calls ---------------------------------------------------- */
int vs_bidir_index_main(int argc, char **argv);
int vs_bidir_mapping_main(int argc, char **argv);

/* ---- microbenchmarks used by bench.py for the roofline denominators ---------------------------- */
/* thread-level LOP3 instructions per second of the device (alu pipe), and LDS 32-bit lane-words per second */
int vs_measure_int_peaks(vs_ctx *ctx, double *lop3_per_s, double *lds_words_per_s);

#ifdef __cplusplus
}
#endif
#endif
