/*
 * oracle/vo_cli.c — CPU ORACLE command line (test infrastructure, NOT product code).
 *
 * `oracle_bidir_mapping -G genome.fa -R guides.fa -M k [-P XY] -O out.sam [--mode scan|literal]
 *                       [--key ref16|wide] [--md-style seqan|samtools] [-T threads]`
 * restates main() of VARSCOT_pipeline/read_mapping/bidir_mapping.cpp:190-312 closely enough that
 * the product CLI's SAM can be diffed byte-for-byte against this one.  -I is accepted and ignored
 * (the oracle has no index).  PARITY UNPINNED, see vo_core.h.
 */
#include "vo_core.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    char   **names;  /* whole header line after '>' (SeqAn readRecord id; bidir_mapping.cpp:272-280) */
    uint64_t *off;   /* n+1 */
    uint8_t *codes;  /* concatenated */
    uint32_t n;
    uint64_t total;
} fasta_t;

static int read_fasta(const char *path, int guide_alphabet, fasta_t *f)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return -1;
    memset(f, 0, sizeof(*f));
    uint64_t cap = 1 << 20, ncap = 16;
    f->codes = (uint8_t *)malloc(cap);
    f->names = (char **)malloc(ncap * sizeof(char *));
    f->off = (uint64_t *)malloc((ncap + 1) * sizeof(uint64_t));
    char *line = NULL; size_t lcap = 0; ssize_t len;
    int have = 0;
    while ((len = getline(&line, &lcap, fp)) >= 0) {
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        if (len > 0 && line[0] == '>') {
            if (f->n == ncap) {
                ncap *= 2;
                f->names = (char **)realloc(f->names, ncap * sizeof(char *));
                f->off = (uint64_t *)realloc(f->off, (ncap + 1) * sizeof(uint64_t));
            }
            f->names[f->n] = strdup(line + 1);
            f->off[f->n] = f->total;
            f->n++;
            have = 1;
            continue;
        }
        if (!have) continue;
        for (ssize_t i = 0; i < len; ++i) {
            char c = line[i];
            if (c == ' ' || c == '\t' || c == '\v' || c == '\f') continue;
            if (f->total == cap) { cap *= 2; f->codes = (uint8_t *)realloc(f->codes, cap); }
            f->codes[f->total++] = guide_alphabet ? vo_guide_code(c) : vo_text_code(c);
        }
    }
    free(line);
    fclose(fp);
    f->off[f->n] = f->total;
    return 0;
}

int main(int argc, char **argv)
{
    const char *G = NULL, *R = NULL, *O = NULL, *P = NULL;
    int k = -1, threads = 1, mode = VO_MODE_SCAN, key = VO_KEY_WIDE, md = VO_MD_SEQAN;
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        const char *v = (i + 1 < argc) ? argv[i + 1] : NULL;
        if ((!strcmp(a, "-G") || !strcmp(a, "--genome")) && v) { G = v; ++i; }
        else if ((!strcmp(a, "-I") || !strcmp(a, "--index")) && v) { ++i; }
        else if ((!strcmp(a, "-R") || !strcmp(a, "--reads")) && v) { R = v; ++i; }
        else if ((!strcmp(a, "-M") || !strcmp(a, "--mismatches")) && v) { k = atoi(v); ++i; }
        else if ((!strcmp(a, "-T") || !strcmp(a, "--threads")) && v) { threads = atoi(v); ++i; }
        else if ((!strcmp(a, "-O") || !strcmp(a, "--output")) && v) { O = v; ++i; }
        else if ((!strcmp(a, "-P") || !strcmp(a, "--pam")) && v) { P = v; ++i; }
        else if (!strcmp(a, "--mode") && v) { mode = !strcmp(v, "literal") ? VO_MODE_LITERAL : VO_MODE_SCAN; ++i; }
        else if (!strcmp(a, "--key") && v) { key = !strcmp(v, "ref16") ? VO_KEY_REF16 : VO_KEY_WIDE; ++i; }
        else if (!strcmp(a, "--md-style") && v) { md = !strcmp(v, "samtools") ? VO_MD_SAMTOOLS : VO_MD_SEQAN; ++i; }
        else { fprintf(stderr, "oracle_bidir_mapping: unknown argument %s\n", a); return 1; }
    }
    if (!G || !R || !O || k < 0) { fprintf(stderr, "oracle_bidir_mapping: -G -R -M -O are required\n"); return 1; }
    if (k > 8) { fprintf(stderr, "Error: Maximum number of mismatches must lie between 0 and 8.\n"); return 1; }
    int extra = -1;
    if (P && P[0]) {
        if (strlen(P) != 2) { fprintf(stderr, "oracle_bidir_mapping: -P needs two letters\n"); return 1; }
        uint8_t x = vo_text_code(P[0]), y = vo_text_code(P[1]);
        extra = (x > 3 || y > 3) ? -1 : x * 4 + y;    /* a PAM containing N can never equal a window without N */
    }
    fasta_t guides, genome;
    if (read_fasta(R, 1, &guides)) { fprintf(stderr, "cannot open %s\n", R); return 1; }
    printf("Reads loaded (total: %u).\n", guides.n);
    if (read_fasta(G, 0, &genome)) { fprintf(stderr, "cannot open %s\n", G); return 1; }
    printf("Index loaded.\n");
    for (uint32_t g = 0; g < guides.n; ++g)
        if (guides.off[g + 1] - guides.off[g] != VO_GLEN) { fprintf(stderr, "guide %u is not 23 nt\n", g); return 1; }
    vo_result res;
    int rc = vo_map(genome.codes, genome.off, genome.n, guides.codes, guides.n, k, extra, mode, key, md, threads, &res);
    if (rc) { fprintf(stderr, "vo_map failed (%d)\n", rc); return 1; }
    FILE *out = fopen(O, "wb");
    if (!out) { fprintf(stderr, "ERROR: Could not open output path.\n"); return 1; }
    char buf[4096 + 256];
    for (uint64_t i = 0; i < res.n; ++i) {
        const vo_record *r = &res.rec[i];
        char small[512];
        const char *qn = guides.names[r->guide], *rn = genome.names[r->contig];
        size_t need = strlen(qn) + strlen(rn) + 256;
        char *b = need <= sizeof(buf) ? buf : (char *)malloc(need);
        (void)small;
        int n = vo_format_sam(r, qn, rn, guides.codes + (size_t)r->guide * VO_GLEN, b, need);
        fwrite(b, 1, (size_t)n, out);
        if (b != buf) free(b);
    }
    fclose(out);
    if (res.key16_collisions)
        fprintf(stderr, "oracle: %llu uint16 contig-id key collisions\n", (unsigned long long)res.key16_collisions);
    vo_result_free(&res);
    return 0;
}
