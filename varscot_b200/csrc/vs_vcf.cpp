// varscot_b200/csrc/vs_vcf.cpp — rows f1 and f4 of SURVEY.md section 8f.
// f1: drop-in for the reference's `vcf_loader`
// (VARSCOT_pipeline/variant_processing/vcf_loader.cpp:11-77), the producer of the hot path's second input:
// VCF (one sample) + genome FASTA -> "SNP genome" multi-FASTA of variant haplotype segments.
// Host-only C++ (no SeqAn, no CUDA).  Same argv: vcf_loader FILE.vcf SNPGENOME.fa GENOME.fa SAMPLE SEQLENGTH THREADS
//
//   read_vcf / process_record   process_vcf.h:54-209, 226-269   GT -> variants with allele codes 0/1/2/-1
//   find_max_overlap            overlap_sequences.h:35-162      clusters of variants within seqLength (+ deletions)
//   all_combinations            write_fasta.h:88-229            haplotypes per cluster (phased: <= 2, unphased: 2^u tuples)
//   fasta_id                    write_fasta.h:30-65             <chr>_<start>_REF | <chr>_<start>_ALT(_<pos>_<ref>_<alt>)+
//   generate_sequences          write_fasta.h:303-399           ref gaps interleaved with allele strings
// Inputs on which the reference has undefined behaviour are handled as listed in oracle/vcf_oracle.py (D1-D6).
#include "../../include/varscot_scan.h"
#include "vs_genome.h"
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <map>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

namespace {

using namespace vsg;

struct Variant {                 // process_vcf.h:32-42
    std::string ref, alt;
    unsigned chr = 0;
    long pos = 0, start = 0, end = 0;
    int type = 0;                // 0 substitution, 1 insertion, 2 deletion
    int allele = 0;              // 0 first, 1 second, 2 both, -1 unphased
};
typedef std::vector<Variant> Record;

// `is >> first [>> sep >> second]` (process_vcf.h:93-113)
bool read_int(const std::string &t, size_t &i, long &v)
{
    size_t j = i;
    while (j < t.size() && (t[j] == ' ' || t[j] == '\t')) ++j;
    size_t k = j;
    if (k < t.size() && (t[k] == '+' || t[k] == '-')) ++k;
    size_t d = k;
    while (k < t.size() && std::isdigit((unsigned char)t[k])) ++k;
    if (k == d) return false;
    v = strtol(t.substr(j, k - j).c_str(), nullptr, 10);
    i = k;
    return true;
}

// processRecord, process_vcf.h:54-209
bool process_record(unsigned rid, long pos, const std::string &ref, const std::string &alt, const std::string &format,
                    const std::vector<std::string> &samples, unsigned sample, Record &out, std::string &err)
{
    out.clear();
    if (sample >= samples.size()) { err = "ERROR: Sample index out of range."; return false; }
    std::vector<std::string> entries = split(samples[sample], ':'), fmt = split(format, ':');
    size_t gt_pos = fmt.size();
    for (size_t i = 0; i < fmt.size(); ++i) if (fmt[i] == "GT") { gt_pos = i; break; }
    if (gt_pos >= fmt.size() || gt_pos >= entries.size()) return true;                 // D2
    std::vector<std::string> alts = split(alt, ',');
    const std::string &g = entries[gt_pos];
    size_t i = 0;
    long first = -1, second = -1;
    bool phased = true;
    if (!read_int(g, i, first) || first > (long)alts.size()) return true;
    bool have_second = false;
    if (i < g.size()) {
        char sep = g[i];
        size_t j = i + 1;
        if (read_int(g, j, second) && second <= (long)alts.size()) { have_second = true; if (sep == '/') phased = false; }
    }
    if (!have_second) second = first;                                                  // haploid (Y chromosome)
    if (first < 0 || second < 0) return true;
    Variant vs;
    vs.chr = rid; vs.pos = pos; vs.ref = dna5(ref);
    if (first == 0 && second == 0) return true;
    if (first > 0 && second > 0 && first != second) {
        const std::string &a = alts[(size_t)first - 1], &b = alts[(size_t)second - 1];
        if (a != "." && b != ".") {
            vs.allele = 0; vs.alt = dna5(a); out.push_back(vs);
            vs.allele = 1; vs.alt = dna5(b); out.push_back(vs);
        } else if (a != ".") { vs.allele = 0; vs.alt = dna5(a); out.push_back(vs); }
        else if (b != ".") { vs.allele = 1; vs.alt = dna5(b); out.push_back(vs); }       // D1
        else return true;
    } else {
        if (alts[0] == ".") return true;
        if (first == 0) { vs.allele = 1; vs.alt = dna5(alts[(size_t)second - 1]); }
        else if (second == 0) { vs.allele = 0; vs.alt = dna5(alts[(size_t)first - 1]); }
        else { vs.allele = 2; vs.alt = dna5(alts[(size_t)first - 1]); }
        out.push_back(vs);
    }
    for (Variant &v : out) {
        if (!phased && first != second) v.allele = -1;
        v.type = v.ref.size() > v.alt.size() ? 2 : (v.ref.size() == v.alt.size() ? 0 : 1);
    }
    return true;
}

// processVcfFile, process_vcf.h:226-269
bool read_vcf(const char *path, unsigned sample, std::vector<Record> &all, std::vector<std::string> &chr_table, std::string &err)
{
    FILE *f = fopen(path, "rb");
    if (!f) { err = "ERROR: Could not open VCF file."; return false; }
    std::map<std::string, unsigned> chr_id;
    auto intern = [&](const std::string &name) {
        auto it = chr_id.find(name);
        if (it != chr_id.end()) return it->second;
        unsigned id = (unsigned)chr_table.size();
        chr_table.push_back(name); chr_id[name] = id;
        return id;
    };
    char *line = nullptr; size_t cap = 0; ssize_t len;
    bool ok = true;
    while ((len = getline(&line, &cap, f)) >= 0) {
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        if (len == 0) continue;
        if (line[0] == '#') {
            if (!strncmp(line, "##contig=<", 10)) {
                const char *id = strstr(line, "ID=");
                if (id) {
                    std::string name(id + 3);
                    size_t e = name.find_first_of(",>");
                    intern(name.substr(0, e));
                }
            }
            continue;
        }
        std::vector<std::string> fld = split(std::string(line, (size_t)len), '\t');
        if (fld.size() < 10) continue;
        unsigned rid = intern(fld[0]);
        std::vector<std::string> samples(fld.begin() + 9, fld.end());
        Record r;
        if (!process_record(rid, strtol(fld[1].c_str(), nullptr, 10) - 1, fld[3], fld[4], fld[8], samples, sample, r, err)) { ok = false; break; }
        if (!r.empty()) all.push_back(std::move(r));
    }
    free(line);
    fclose(f);
    return ok;
}

// findMaxOverlap, overlap_sequences.h:35-162
void find_max_overlap(std::vector<Record> &all, const std::vector<unsigned> &idx, long seq_len,
                      std::vector<std::pair<unsigned, unsigned>> &regions, std::vector<unsigned> &centers)
{
    const long n = (long)idx.size();
    std::vector<long> max_del((size_t)n, 0);
    for (long i = 0; i < n; ++i)
        for (const Variant &v : all[idx[(size_t)i]])
            if (v.type == 2) max_del[(size_t)i] = std::max(max_del[(size_t)i], (long)v.ref.size() - (long)v.alt.size());
    auto pos = [&](long i) { return all[idx[(size_t)i]][0].pos; };
    long r1 = 0, r2 = 0;
    for (long i = 0; i < n; ++i) {
        long wsl, wsr;
        if (r2 > i) {
            long right = r2;
            wsr = seq_len + max_del[(size_t)i];
            if (right < n) for (long d = i + 1; d <= right; ++d) wsr += max_del[(size_t)d];
            while (right < n && pos(right) - pos(i) < wsr) { wsr += max_del[(size_t)right]; ++right; }
            if (right == r2) {
                for (Variant &v : all[centers.back()]) v.end = pos(i) + wsr;            // D4
                continue;
            }
            r2 = right;
            long left = i - 1;
            wsl = seq_len + max_del[(size_t)left];
            while (left >= 0 && pos(i) - pos(left) < wsl) { --left; wsl += left >= 0 ? max_del[(size_t)left] : 0; }   // D3
            if (left + 1 == r1) {
                for (Variant &v : all[centers.back()]) v.end = pos(i) + wsr;
                regions.back().second = (unsigned)right;
                continue;
            }
            r1 = left + 1;
        } else {
            wsr = seq_len + max_del[(size_t)i];
            long right = i + 1;
            while (right < n && pos(right) - pos(i) < wsr) { wsr += max_del[(size_t)right]; ++right; }
            r2 = right;
            wsl = seq_len;
            r1 = i;
        }
        regions.emplace_back((unsigned)r1, (unsigned)r2);
        centers.push_back(idx[(size_t)i]);
        for (Variant &v : all[idx[(size_t)i]]) { v.start = std::max(0L, v.pos - wsl + 1); v.end = v.pos + wsr; }   // D5
    }
}

// getFastaID, write_fasta.h:30-65
std::string fasta_id(const std::vector<Record> &all, const std::vector<unsigned> &idx, unsigned first, unsigned center,
                     const std::vector<int> &comb, const std::string &chr)
{
    std::string id = chr + "_" + std::to_string(all[center][0].start) + "_";
    bool all_ref = std::all_of(comb.begin(), comb.end(), [](int c) { return c == -1; });
    if (all_ref) return id + "REF";
    id += "ALT";
    for (size_t i = 0; i < comb.size(); ++i)
        if (comb[i] != -1) {
            const Variant &v = all[idx[first + i]][(size_t)comb[i]];
            id += "_" + std::to_string(v.pos) + "_" + v.ref + "_" + v.alt;
        }
    return id;
}

// allCombinations, write_fasta.h:88-229
void all_combinations(const std::vector<Record> &all, const std::vector<unsigned> &idx, unsigned first, unsigned last, unsigned center,
                      const std::string &chr, std::vector<std::vector<std::string>> &combos, std::vector<std::string> &ids)
{
    const unsigned size = last - first;
    std::vector<unsigned> unphased;
    std::vector<std::string> fa(size), sa(size);
    std::vector<int> ifa(size, 0), isa(size, 0);
    for (unsigned i = first; i < last; ++i) {
        const Record &rec = all[idx[i]];
        const unsigned j = i - first;
        if (rec[0].allele == -1) unphased.push_back(j);
        else if (rec.size() == 2) { fa[j] = rec[0].alt; ifa[j] = 0; sa[j] = rec[1].alt; isa[j] = 1; }
        else if (rec[0].allele == 0) { fa[j] = rec[0].alt; ifa[j] = 0; sa[j] = rec[0].ref; isa[j] = -1; }
        else if (rec[0].allele == 1) { fa[j] = rec[0].ref; ifa[j] = -1; sa[j] = rec[0].alt; isa[j] = 0; }
        else { fa[j] = rec[0].alt; ifa[j] = 0; sa[j] = rec[0].alt; }                     // :140-146: second index stays 0
    }
    auto emit = [&]() {
        combos.push_back(fa);
        ids.push_back(fasta_id(all, idx, first, center, ifa, chr));
        if (ifa != isa) { combos.push_back(sa); ids.push_back(fasta_id(all, idx, first, center, isa, chr)); }
    };
    if (!unphased.empty()) {
        const unsigned u = (unsigned)unphased.size();
        for (unsigned long t = 0; t < (1ul << u); ++t) {        // lexicographic tuples, as the stack of :155-214 produces them
            for (unsigned b = 0; b < u; ++b) {
                const int bit = (int)((t >> (u - 1 - b)) & 1);
                const unsigned j = unphased[b];
                const Record &rec = all[idx[first + j]];
                if (rec.size() == 2) { fa[j] = sa[j] = rec[(size_t)bit].alt; ifa[j] = isa[j] = bit; }
                else if (bit == 0) { fa[j] = sa[j] = rec[0].ref; ifa[j] = isa[j] = -1; }
                else { fa[j] = sa[j] = rec[0].alt; ifa[j] = isa[j] = 0; }
            }
            emit();
        }
    } else emit();
}

}  // namespace

extern "C" int vs_vcf_loader_main(int argc, char **argv)
{
    if (argc != 7) { fprintf(stderr, "USAGE: vcf_loader FILE.vcf SNPGENOME.fa GENOME.fa SAMPLE SEQLENGTH THREADS\n"); return 1; }
    auto cast = [](const char *s, unsigned &v) {
        if (!*s) return false;
        char *end = nullptr;
        long x = strtol(s, &end, 10);
        if (!end || *end || x < 0) return false;
        v = (unsigned)x;
        return true;
    };
    unsigned sample, seq_len, threads;
    if (!cast(argv[4], sample)) { fprintf(stderr, "ERROR: Cannot cast %s into an unsigned.\n", argv[4]); return 1; }
    if (!cast(argv[5], seq_len)) { fprintf(stderr, "ERROR: Cannot cast %s into an unsigned.\n", argv[5]); return 1; }
    if (!cast(argv[6], threads)) { fprintf(stderr, "ERROR: Cannot cast %s into an unsigned.\n", argv[6]); return 1; }
    printf("Process records\n");
    std::vector<Record> all;
    std::vector<std::string> chr_table;
    std::string err;
    if (!read_vcf(argv[1], sample, all, chr_table, err)) { printf("%s\n", err.c_str()); return 1; }
    printf("Compute overlap sequences\n");
    const size_t nchr = chr_table.size();
    std::vector<std::vector<unsigned>> idx(nchr);
    for (size_t i = 0; i < all.size(); ++i) idx[all[i][0].chr].push_back((unsigned)i);
    std::vector<std::vector<std::pair<unsigned, unsigned>>> regions(nchr);
    std::vector<std::vector<unsigned>> centers(nchr);
    for (size_t c = 0; c < nchr; ++c) {
        std::stable_sort(idx[c].begin(), idx[c].end(), [&](unsigned a, unsigned b) { return all[a][0].pos < all[b][0].pos; });   // D6
        find_max_overlap(all, idx[c], (long)seq_len, regions[c], centers[c]);
    }
    printf("Write fasta\n");
    fflush(stdout);
    FILE *out = fopen(argv[2], "wb");
    if (!out) { printf("ERROR: Could not open single FASTA output file.\n"); return 1; }
    Genome g;
    if (!open_genome(argv[3], g, err)) { printf("%s\n", err.c_str()); fclose(out); return 1; }
    std::string buf;
    for (size_t c = 0; c < nchr; ++c) {
        if (regions[c].empty()) continue;
        auto it = g.by_name.find(chr_table[c]);
        if (it == g.by_name.end()) { printf("ERROR: Index out of range.\n"); break; }       // write_fasta.h:249-252 -> caught, file truncated
        const size_t gi = it->second;
        for (size_t r = 0; r < regions[c].size(); ++r) {
            const unsigned r1 = regions[c][r].first, r2 = regions[c][r].second, center = centers[c][r];
            const Variant &cv = all[center][0];
            std::vector<std::string> base;
            for (unsigned i = r1; i <= r2; ++i) {
                long b, e;
                if (i == r1) { b = cv.start; e = all[idx[c][i]][0].pos; }
                else {
                    const Variant &p = all[idx[c][i - 1]][0];
                    b = p.pos + (long)p.ref.size();
                    e = i == r2 ? cv.end : all[idx[c][i]][0].pos;
                }
                base.push_back(extract(g, gi, b, e));
            }
            std::vector<std::vector<std::string>> combos;
            std::vector<std::string> ids;
            all_combinations(all, idx[c], r1, r2, center, chr_table[c], combos, ids);
            for (size_t k = 0; k < combos.size(); ++k) {
                std::string seq;
                for (size_t j = 0; j < combos[k].size(); ++j) { seq += base[j]; seq += combos[k][j]; }
                seq += base[combos[k].size()];
                buf.clear();
                buf += ">"; buf += ids[k]; buf += "\n";
                for (size_t p = 0; p < seq.size(); p += 70) { buf.append(seq, p, 70); buf += "\n"; }     // SeqAn writes 70 columns
                if (seq.empty()) buf += "\n";
                fwrite(buf.data(), 1, buf.size(), out);
            }
        }
    }
    fclose(out);
    (void)threads;
    return 0;
}

// ---- row f4: fasta_writer (VARSCOT_pipeline/variant_processing/fasta_writer.cpp:8-41, extract_fasta_ontargets.h) -------
// `fasta_writer OUTPUT1.fa OUTPUT2.fa ONTARGETS.bed GENOME.fa`: BED6 on-targets -> 23-nt guide FASTA (OUTPUT1, what
// bidir_mapping -R reads) and the 30-nt flanking FASTA for TUSCAN (OUTPUT2).  Half-open 0-based BED coordinates,
// reverse complement on '-', flanks +4/+3 on '+', +3/+4 on '-' (extract_fasta_ontargets.h:44-53).  A flank that would
// start before the contig (unsigned wrap in the reference) is clamped at 0.
namespace {

bool write_ontargets(const char *out_path, const char *bed_path, const Genome &g, bool flanking, std::string &err)
{
    FILE *out = fopen(out_path, "wb");
    if (!out) { err = "ERROR: Could not open output file."; return false; }
    FILE *bed = fopen(bed_path, "rb");
    if (!bed) { fclose(out); err = "ERROR: Could not open BED file."; return false; }
    char *line = nullptr; size_t cap = 0; ssize_t len;
    bool ok = true;
    while ((len = getline(&line, &cap, bed)) >= 0) {
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        if (len == 0 || line[0] == '#' || !strncmp(line, "track", 5) || !strncmp(line, "browser", 7)) continue;
        std::vector<std::string> f = split(std::string(line, (size_t)len), '\t');
        if (f.size() < 6) continue;
        auto it = g.by_name.find(f[0]);
        if (it == g.by_name.end()) { err = "ERROR: Index out of range."; ok = false; break; }     // extract_fasta_ontargets.h:37-40
        long b = strtol(f[1].c_str(), nullptr, 10), e = strtol(f[2].c_str(), nullptr, 10);
        const char strand = f[5].empty() ? '.' : f[5][0];
        if (flanking && strand == '+') { b -= 4; e += 3; }
        else if (flanking && strand == '-') { b -= 3; e += 4; }
        std::string seq = extract(g, it->second, b, e);
        if (strand == '-') {
            std::reverse(seq.begin(), seq.end());
            for (char &c : seq) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
        }
        std::string buf = ">" + f[3] + "\n";
        for (size_t p = 0; p < seq.size(); p += 70) { buf.append(seq, p, 70); buf += "\n"; }
        if (seq.empty()) buf += "\n";
        fwrite(buf.data(), 1, buf.size(), out);
    }
    free(line);
    fclose(bed);
    fclose(out);
    return ok;
}

}  // namespace

extern "C" int vs_fasta_writer_main(int argc, char **argv)
{
    if (argc != 5) { fprintf(stderr, "USAGE: extract_fasta_ontargets OUTPUT1.fa OUTPUT2.fa ONTARGETS.bed GENOME.fa\n"); return 1; }
    Genome g;
    std::string err;
    if (!open_genome(argv[4], g, err)) { printf("%s\n", err.c_str()); return 1; }
    if (!write_ontargets(argv[1], argv[3], g, false, err)) { printf("%s\n", err.c_str()); return 1; }
    if (!write_ontargets(argv[2], argv[3], g, true, err)) { printf("%s\n", err.c_str()); return 1; }
    return 0;
}
