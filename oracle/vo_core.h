/*
 * oracle/vo_core.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the semantics of VARSCOT's read_mapping stage
 * (reference: VARSCOT_pipeline/read_mapping/bidir_mapping.cpp).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use anything under oracle/.  The product (varscot_b200/, the C-ABI
 * library and the CLI executables) never links or calls it.
 *
 * PARITY UNPINNED: the reference's search arithmetic lives in SeqAn 2.4.0rc2
 * (seqan/seqan tag seqan-v2.4.0rc2, VARSCOT_pipeline/Dockerfile:41), which is
 * not vendored under /root/reference and not installed; the only golden output
 * (workflow/guideseq-data/bidir_guideseq.sam) is a git-LFS pointer.  So this
 * oracle cannot be checked against reference outputs here.  It is instead
 * written twice (a literal seed-and-verify restatement and a rule-based scan)
 * and the two are tested against each other and against hand-derived
 * known-answer vectors (tests/golden/).
 */
#ifndef VO_CORE_H
#define VO_CORE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VO_GLEN 23          /* guide length incl. PAM; CIGAR "23M" is hard-coded, bidir_mapping.cpp:102 */
#define VO_HALF1 11         /* length(read)/2, bidir_mapping.cpp:157 */
#define VO_MD_MAX 64

/* search mode */
#define VO_MODE_LITERAL 0   /* seed (half, <=K errors) -> delegate, as bidir_mapping.cpp:31-148 */
#define VO_MODE_SCAN    1   /* distilled rules R1-R4 evaluated on every window (fast, OpenMP) */

/* dedup / sort key */
#define VO_KEY_REF16 0      /* literal TOccType = Pair<uint16_t,uint32_t> (bidir_mapping.cpp:13): first found wins */
#define VO_KEY_WIDE  1      /* order (id & 0xFFFF, pos, id >> 16), nothing dropped */

/* MD style */
#define VO_MD_SEQAN    0    /* SeqAn bam_io getMDString: no "0" between adjacent mismatches, none leading/trailing */
#define VO_MD_SAMTOOLS 1    /* SAM-spec style: [0-9]+(([A-Z])[0-9]+)* */

typedef struct {
    uint32_t guide;         /* index of the guide in input order */
    uint32_t contig;        /* full 32-bit contig id (record.r.rID, bidir_mapping.cpp:99) */
    uint32_t pos;           /* 0-based begin (record.r.beginPos, :100) */
    uint16_t flag;          /* 0 / 16 / 256 / 272 (:92-98, :167-187) */
    uint8_t  mm;            /* NM (:121) */
    uint8_t  pad;
    char     md[VO_MD_MAX]; /* MD:Z value (:113-122) */
} vo_record;

typedef struct {
    vo_record *rec;         /* emission order: guide, forward pass, reverse pass (:285-295), R8 inside a pass */
    uint64_t   n;
    uint64_t   cap;
    uint64_t   key16_collisions;  /* distinct (contig,pos) hits of one pass sharing (contig & 0xFFFF, pos) */
} vo_result;

/* alphabet conversions (SeqAn Dna / Dna5 char tables) */
uint8_t vo_text_code(char c);   /* A,C,G,T(U) -> 0..3 case-insensitive, everything else -> 4 (N)   [R6] */
uint8_t vo_guide_code(char c);  /* A,C,G,T(U) -> 0..3 case-insensitive, everything else -> 0 (A)   [R5] */

/*
 * text      Dna5 codes (0..4), all contigs concatenated
 * off       n_contigs+1 offsets into text
 * guides    n_guides x 23 Dna codes (0..3)
 * k         -M, 0..8
 * extra_pam -1, or 4*x+y for "-P XY" in Dna codes (codes > 3 never match)
 * Returns 0 on success, non-zero on bad arguments / out of memory.
 */
int vo_map(const uint8_t *text, const uint64_t *off, uint32_t n_contigs,
           const uint8_t *guides, uint32_t n_guides,
           int k, int extra_pam, int mode, int key_mode, int md_style,
           int n_threads, vo_result *out);

void vo_result_free(vo_result *r);

/* Count-only rule-based scan for CPU-baseline timing: returns number of hits (R1-R4), no records. */
uint64_t vo_scan_count(const uint8_t *text, const uint64_t *off, uint32_t n_contigs,
                       const uint8_t *guides, uint32_t n_guides,
                       int k, int extra_pam, int n_threads);

/* MD string of window (genome, Dna5 codes) against pattern (Dna codes), genome-forward orientation. */
void vo_md_string(const uint8_t *window, const uint8_t *pattern, int md_style, char *out);

/* Format one record as the 13-column SAM line of R9 (no trailing NUL issues; returns length). */
int vo_format_sam(const vo_record *r, const char *qname, const char *rname,
                  const uint8_t *guide_codes, char *buf, size_t buflen);

int vo_num_procs(void);

#ifdef __cplusplus
}
#endif
#endif
