// varscot_b200/csrc/vs_merge.cpp — row f3 of SURVEY.md section 8f: drop-ins for the reference's `bam_merger` and
// `bam_merger_ref_only` (VARSCOT_pipeline/variant_processing/bam_merger.cpp:8-62, bam_merger_ref_only.cpp:8-55), the
// consumers of the mapper's SAM files: they drop the on-target itself, reference hits that lie wholly inside a variant
// segment and adjacent duplicates, remap variant-segment hits to chromosome coordinates, and write the user-visible
// hit table (+ the 442-feature matrix for the random-forest classifier).  Host-only C++ (no SeqAn, no CUDA).
//
//   read_sam               filter_output_bam.h:362-418   one PotentialOffTarget per SAM record, sequence from the FAI
//   mismatch_positions     filter_output_bam.h:330-349   `while (is >> num >> base)` over the MD value
//   read_ontargets         filter_output_bam.h:449-483
//   read_snp_table         filter_output_bam.h:425-439   id split on '_' + sequence length
//   filter_ref / filter_snp / snp_type   filter_output_bam.h:70-124, 279-317, 189-263
//   mit_score              mit_score.h:12-68
//   feature_record / feature_names / read_tuscan   feature_matrix.h:25-126, 140-204, 206-230
//   table writers          merge_output_bam.h:46-215, 244-460, 462-722
// Inputs on which the reference is undefined are handled as listed in oracle/merge_oracle.py (M1-M3).
#include "../../include/varscot_scan.h"
#include "vs_genome.h"
#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <cstdlib>
#include <stdexcept>

namespace {

using namespace vsg;

struct Pot {                    // PotentialOffTarget, filter_output_bam.h:23-32
    std::string chr, target, snp_type = "REF", sequence;
    std::vector<int> mm;
    long pos = 0;
    char strand = '+';
};

bool same(const Pot &a, const Pot &b)       // comp(), filter_output_bam.h:40-49
{
    return a.target == b.target && a.chr == b.chr && a.pos == b.pos && a.strand == b.strand && a.sequence == b.sequence &&
           a.mm == b.mm && a.snp_type == b.snp_type;
}

std::string fetch(const Genome &g, const std::string &name, long b, long e, char strand)
{
    auto it = g.by_name.find(name);
    if (it == g.by_name.end()) throw std::out_of_range("ERROR: Index out of range.");
    std::string s = extract(g, it->second, b, e);
    return strand == '-' ? revcomp(s) : s;
}

std::vector<int> mismatch_positions(const std::string &md)
{
    std::vector<int> out;
    size_t i = 0, n = md.size();
    long pos = 0;
    for (;;) {
        while (i < n && (md[i] == ' ' || md[i] == '\t' || md[i] == '\n')) ++i;
        size_t j = i;
        while (j < n && std::isdigit((unsigned char)md[j])) ++j;
        if (j == i) break;                              // `is >> num` fails
        long num = strtol(md.substr(i, j - i).c_str(), nullptr, 10);
        while (j < n && (md[j] == ' ' || md[j] == '\t' || md[j] == '\n')) ++j;
        if (j >= n) break;                              // `>> base` fails at the end
        pos += num + 1;
        out.push_back((int)(pos - 1));
        i = j + 1;
    }
    if (out.empty()) out.push_back(-1);
    return out;
}

std::vector<Pot> read_sam(const char *path, const Genome &g)
{
    FILE *f = fopen(path, "rb");
    if (!f) throw std::runtime_error("ERROR: Could not open BAM file.");
    std::vector<Pot> out;
    char *line = nullptr; size_t cap = 0; ssize_t len;
    try {
        while ((len = getline(&line, &cap, f)) >= 0) {
            while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
            if (len == 0 || line[0] == '@') continue;
            std::vector<std::string> fld = split(std::string(line, (size_t)len), '\t');
            if (fld.size() < 11) continue;
            Pot p;
            p.target = fld[0]; p.chr = fld[2];
            p.pos = strtol(fld[3].c_str(), nullptr, 10) - 1;
            p.strand = (strtol(fld[1].c_str(), nullptr, 10) & 16) ? '-' : '+';
            std::string key = p.chr.substr(0, p.chr.find_first_of(" \t"));
            p.sequence = fetch(g, key, p.pos, p.pos + 23, p.strand);
            std::string md;
            for (size_t t = 11; t < fld.size(); ++t) if (fld[t].compare(0, 5, "MD:Z:") == 0) md = fld[t].substr(5);
            p.mm = mismatch_positions(md);
            out.push_back(std::move(p));
        }
    } catch (...) { free(line); fclose(f); throw; }
    free(line);
    fclose(f);
    return out;
}

void read_ontargets(const char *path, const Genome &g, std::map<std::string, Pot> &on, std::map<std::string, unsigned> &count)
{
    FILE *f = fopen(path, "rb");
    if (!f) throw std::runtime_error("ERROR: Could not open BED file.");
    char *line = nullptr; size_t cap = 0; ssize_t len;
    try {
        while ((len = getline(&line, &cap, f)) >= 0) {
            while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
            if (len == 0 || line[0] == '#' || !strncmp(line, "track", 5) || !strncmp(line, "browser", 7)) continue;
            std::vector<std::string> fld = split(std::string(line, (size_t)len), '\t');
            if (fld.size() < 6) continue;
            Pot p;
            p.target = fld[3]; p.chr = fld[0]; p.pos = strtol(fld[1].c_str(), nullptr, 10);
            p.strand = fld[5].empty() ? '.' : fld[5][0];
            p.sequence = fetch(g, p.chr, p.pos, p.pos + 23, p.strand);
            p.mm = {-1};
            on.emplace(p.target, p);               // the first record of a name wins (std::map::insert)
            count.emplace(p.target, 0u);
        }
    } catch (...) { free(line); fclose(f); throw; }
    free(line);
    fclose(f);
}

struct SnpInfo { std::vector<std::string> id; long start = 0, length = 0; };

void read_snp_table(const char *path, std::vector<SnpInfo> &table, std::map<std::string, std::vector<size_t>> &by_chr)
{
    FILE *f = fopen(path, "rb");
    if (!f) throw std::runtime_error("ERROR: Could not open variant genome FASTA file.");
    char *line = nullptr; size_t cap = 0; ssize_t len;
    bool have = false;
    SnpInfo cur;
    auto flush = [&]() {
        if (!have) return;
        cur.start = cur.id.size() > 1 ? atol(cur.id[1].c_str()) : 0;
        by_chr[cur.id[0]].push_back(table.size());
        table.push_back(cur);
    };
    while ((len = getline(&line, &cap, f)) >= 0) {
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        if (len > 0 && line[0] == '>') {
            flush();
            cur = SnpInfo();
            cur.id = split(std::string(line + 1, (size_t)len - 1), '_');
            have = true;
        } else if (have) {
            for (ssize_t i = 0; i < len; ++i) if (!std::isspace((unsigned char)line[i])) ++cur.length;
        }
    }
    flush();
    free(line);
    fclose(f);
}

// getSnpType, filter_output_bam.h:189-263
void snp_type(const std::vector<std::string> &fid, long &pos, long seq_len, std::string &type)
{
    std::string vars;
    long count = 0;
    bool start_found = false;
    for (size_t i = 3; i + 2 < fid.size(); i += 3) {
        const long p = atol(fid[i].c_str()), lr = (long)fid[i + 1].size(), la = (long)fid[i + 2].size();
        auto inside = [&](long q) { return pos <= q && pos + seq_len > q; };
        if (lr == la) {
            if (inside(p)) { vars += fid[i] + ","; start_found = true; }
        } else if (lr < la) {
            if (inside(p + 1) || inside(p + la - 1)) { vars += fid[i] + ","; start_found = true; }
            else if (!start_found) count -= la - lr;
        } else {
            if (inside(p + 1) || inside(p + lr - 1)) { vars += fid[i] + ","; start_found = true; }
            else if (!start_found) count += lr - la;
        }
    }
    pos += count;
    if (!vars.empty()) type = "VAR_" + fid[0] + "_" + vars.substr(0, vars.size() - 1);
}

const double MATRIX_M[20] = {0, 0, 0.014, 0, 0, 0.395, 0.317, 0, 0.389, 0.079, 0.445, 0.508, 0.613, 0.851, 0.732, 0.828, 0.615, 0.804, 0.685, 0.583};

double mit_score(const std::vector<int> &mm)       // calcMitScore, mit_score.h:12-68
{
    if (mm.size() == 1 && mm[0] == -1) return 100;
    const size_t nm = mm.back() < 20 ? mm.size() : mm.size() - 1;
    if (nm == 0) return 100;
    const double s3 = 1.0 / std::pow((double)nm, 2);
    double s1 = 1, sum = 0;
    for (size_t i = 0; i < nm; ++i) {
        s1 *= 1 - ((mm[i] >= 0 && mm[i] < 20) ? MATRIX_M[mm[i]] : 0.0);       // M1
        if (i > 0) sum += mm[i] - mm[i - 1];
    }
    double s2 = 1;
    if (nm >= 2) {
        const double avg = sum / (double)(nm - 1);
        s2 = 1 / (((19 - avg) / 19) * 4 + 1);
    }
    return s1 * s2 * s3 * 100;
}

std::string fmt_double(double x)                    // std::ostream << double, default precision
{
    char b[64];
    snprintf(b, sizeof b, "%g", x);
    return b;
}

const char *PAIRS[16] = {"AA", "AC", "AG", "AT", "CA", "CC", "CG", "CT", "GA", "GC", "GG", "GT", "TA", "TC", "TG", "TT"};
const char *MTYPE_NAMES[12] = {"AtoC", "AtoG", "AtoT", "CtoA", "CtoG", "CtoT", "GtoA", "GtoC", "GtoT", "TtoA", "TtoC", "TtoG"};

inline int code(char c) { return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; }      // M2: N folds to A

void feature_record(const std::string &on, const std::string &off, unsigned (&f)[442])   // feature_matrix.h:25-126
{
    memset(f, 0, sizeof f);
    bool prec = false;
    for (size_t i = 0; i + 2 < off.size(); ++i) {
        if (i < 19) {
            const int pr = code(off[i]) * 4 + code(off[i + 1]);
            f[120 + i * 16 + (size_t)pr] = 1;
            f[424 + pr]++;
        }
        f[36 + i * 4 + (size_t)code(off[i])] = 1;
        if (i < on.size() && on[i] != off[i]) {
            f[0]++;
            f[i + 1] = 1;
            if (i > 7 && i < 20) f[441]++;
            if (prec) f[440]++;
            prec = true;
            const int a = code(on[i]), b = code(off[i]);
            const bool transition = (a == 0 && b == 2) || (a == 1 && b == 3) || (a == 2 && b == 0) || (a == 3 && b == 1);
            if (transition) f[34]++; else f[35]++;
            // mismatch type index over the 12 ordered pairs of different letters (AC AG AT CA CG CT GA GC GT TA TC TG)
            const int idx = a == b ? 0 : a * 3 + (b > a ? b - 1 : b);
            f[22 + idx] = 1;
        } else prec = false;
    }
}

std::vector<std::string> feature_names(unsigned seq_len)        // feature_matrix.h:140-204
{
    std::vector<std::string> n(443);
    const char L[4] = {'A', 'C', 'G', 'T'};
    n[0] = "totalMismatches";
    for (unsigned i = 1; i + 1 < seq_len && i < 22; ++i) n[i] = "mismatchPos" + std::to_string(i);
    for (unsigned i = 0; i < 12; ++i) n[22 + i] = MTYPE_NAMES[i];
    n[34] = "transitionNumber"; n[35] = "transversionNumber";
    for (unsigned i = 1; i + 2 < seq_len && i <= 20; ++i)
        for (unsigned j = 0; j < 4; ++j) n[36 + (i - 1) * 4 + j] = std::string(1, L[j]) + std::to_string(i);
    n[116] = "PAMA"; n[117] = "PAMC"; n[118] = "PAMG"; n[119] = "PAMT";
    for (unsigned i = 1; i + 3 < seq_len && i <= 19; ++i)
        for (unsigned j = 0; j < 16; ++j) n[120 + (i - 1) * 16 + j] = PAIRS[j] + std::to_string(i);
    for (unsigned j = 0; j < 16; ++j) n[424 + j] = PAIRS[j];
    n[440] = "adjacentMismatches"; n[441] = "seedMismatches"; n[442] = "ontargetActivity";
    return n;
}

std::map<std::string, double> read_tuscan(const char *path)       // feature_matrix.h:206-230
{
    FILE *f = fopen(path, "rb");
    if (!f) throw std::runtime_error("ERROR: Could not open on-target activity file.");
    std::map<std::string, double> out;
    char *line = nullptr; size_t cap = 0; ssize_t len;
    while ((len = getline(&line, &cap, f)) >= 0) {
        char name[4096], seq[4096]; double score;
        char *end = nullptr;
        char third[256];
        if (sscanf(line, "%4095s %4095s %255s", name, seq, third) == 3) {
            score = strtod(third, &end);
            if (end != third) out.emplace(name, score);              // `is >> score` needs a leading number
        }
    }
    free(line);
    fclose(f);
    return out;
}

struct Writer {
    FILE *table = nullptr, *matrix = nullptr;
    bool use_mit = true, with_variants = true;
    const std::map<std::string, Pot> *on = nullptr;
    std::map<std::string, unsigned> *count = nullptr;
    const std::map<std::string, double> *activity = nullptr;

    void row(const Pot &o)
    {
        unsigned &c = count->at(o.target);
        ++c;
        const std::string name = o.target + "_" + std::to_string(c);
        std::string r = o.chr + "\t" + std::to_string(o.pos) + "\t" + std::to_string(o.pos + 23) + "\t" + name + "\t" +
                        (use_mit ? fmt_double(mit_score(o.mm)) : std::string(".")) + "\t" + std::string(1, o.strand) + "\t" + o.sequence + "\t";
        if (o.mm.size() == 1 && o.mm[0] == -1) r += "0\t";
        else {
            r += std::to_string(o.mm.size()) + "\t";
            for (size_t j = 0; j < o.mm.size(); ++j) { if (j) r += ","; r += std::to_string(o.mm[j]); }
        }
        if (with_variants) r += "\t" + o.snp_type;
        r += "\n";
        fwrite(r.data(), 1, r.size(), table);
        if (matrix) {
            auto a = activity->find(o.target);
            if (a == activity->end()) throw std::runtime_error("ERROR: no on-target activity for " + o.target);      // M3
            unsigned f[442];
            feature_record(on->at(o.target).sequence, o.sequence, f);
            std::string m = name + "\t";
            for (unsigned x : f) { m += std::to_string(x); m += "\t"; }
            m += fmt_double(a->second) + "\n";
            fwrite(m.data(), 1, m.size(), matrix);
        }
    }
};

const char *HEADER = "#Chr\tStart\tEnd\tTargetsite\tScore\tStrand\tSequence\tMismatch_Number\tMismatch_Positions";

bool cast_unsigned(const char *s, unsigned &v)
{
    if (!*s) return false;
    char *end = nullptr;
    long x = strtol(s, &end, 10);
    if (!end || *end || x < 0) return false;
    v = (unsigned)x;
    return true;
}

void open_outputs(Writer &w, const char *table, const char *matrix, unsigned seq_len, bool variants)
{
    w.table = fopen(table, "wb");
    if (!w.table) throw std::runtime_error("ERROR: Could not open output file.");
    std::string h = std::string(HEADER) + (variants ? "\tVariants\n" : "\n");
    fwrite(h.data(), 1, h.size(), w.table);
    if (matrix) {
        w.matrix = fopen(matrix, "wb");
        if (w.matrix) {
            std::vector<std::string> n = feature_names(seq_len);
            std::string m;
            for (size_t i = 0; i < n.size(); ++i) { m += n[i]; m += i + 1 < n.size() ? "\t" : "\n"; }
            fwrite(m.data(), 1, m.size(), w.matrix);
        }
    }
}

}  // namespace

extern "C" int vs_bam_merger_main(int argc, char **argv)
{
    if (argc != 13) {
        fprintf(stderr, "USAGE: bam_merger RESULT_MERGED.txt FEATURE_MATRIX.txt RESULT_REF.bam RESULT_SNP.bam ONTARGETS.bed GENOME.fa VARIANT_GENOME.fa TUSCAN_REGRESSION.txt NUMMISMATCHES SEQLENGTH THREADS MIT\n");
        return 1;
    }
    unsigned k, seq_len, threads, mit;
    for (int i = 9; i <= 12; ++i) {
        unsigned &dst = i == 9 ? k : i == 10 ? seq_len : i == 11 ? threads : mit;
        if (!cast_unsigned(argv[i], dst)) { fprintf(stderr, "ERROR: Cannot cast %s into an unsigned.\n", argv[i]); return 1; }
    }
    Writer w;
    try {
        std::vector<SnpInfo> table;
        std::map<std::string, std::vector<size_t>> by_chr;
        read_snp_table(argv[7], table, by_chr);
        Genome ref, snp;
        std::string err;
        if (!open_genome(argv[6], ref, err)) throw std::runtime_error("ERROR: Reference index could not be loaded or built.");
        if (!open_genome(argv[7], snp, err)) throw std::runtime_error("ERROR: Variant index could not be loaded or built.");
        std::map<std::string, Pot> on;
        std::map<std::string, unsigned> count;
        read_ontargets(argv[5], ref, on, count);
        printf("Process reference off-targets\n");
        std::vector<Pot> ro = read_sam(argv[3], ref);
        std::vector<size_t> valid_ref;
        // filterRefAlignment (filter_output_bam.h:70-124) drops a reference hit that lies wholly inside a variant segment
        // of its chromosome: exists j with start_j <= pos and pos + seq_len <= start_j + length_j.  The reference tests
        // every segment of the chromosome per hit (O(hits x segments): hours at 5 M variants); the same predicate here is
        // one binary search per hit in the segments sorted by start with a running maximum of their ends.
        struct Cover { std::vector<long> start, max_end; };
        std::map<std::string, Cover> cover;
        for (const auto &kv : by_chr) {
            std::vector<std::pair<long, long>> se;
            se.reserve(kv.second.size());
            for (size_t j : kv.second) se.emplace_back(table[j].start, table[j].start + table[j].length);
            std::sort(se.begin(), se.end());
            Cover &c = cover[kv.first];
            c.start.reserve(se.size()); c.max_end.reserve(se.size());
            long m = std::numeric_limits<long>::min();
            for (const auto &x : se) { m = std::max(m, x.second); c.start.push_back(x.first); c.max_end.push_back(m); }
        }
        for (size_t i = 0; i < ro.size(); ++i) {
            const Pot &o = ro[i];
            bool ok = !same(o, on.at(o.target));
            if (ok) {
                auto it = cover.find(o.chr);
                if (it != cover.end()) {
                    const Cover &c = it->second;
                    const size_t below = (size_t)(std::upper_bound(c.start.begin(), c.start.end(), o.pos) - c.start.begin());   // segments with start <= pos
                    if (below && c.max_end[below - 1] >= o.pos + (long)seq_len) ok = false;
                }
            }
            if (ok) valid_ref.push_back(i);
        }
        printf("Process variant off-targets\n");
        std::vector<Pot> so = read_sam(argv[4], snp);
        std::vector<size_t> valid_snp;
        for (size_t i = 0; i < so.size(); ++i) {                         // filterSnpAlignment, filter_output_bam.h:279-317
            Pot &o = so[i];
            std::vector<std::string> fid = split(o.chr, '_');
            o.chr = fid[0];
            o.pos += fid.size() > 1 ? atol(fid[1].c_str()) : 0;
            snp_type(fid, o.pos, (long)seq_len, o.snp_type);
            bool ok = !same(o, on.at(o.target));
            if (i > 0 && same(o, so[i - 1])) ok = false;
            if (ok) valid_snp.push_back(i);
        }
        std::map<std::string, double> activity = read_tuscan(argv[8]);
        w.use_mit = mit == 0; w.with_variants = true; w.on = &on; w.count = &count; w.activity = &activity;
        open_outputs(w, argv[1], mit == 0 ? nullptr : argv[2], seq_len, true);
        for (size_t i : valid_ref) w.row(ro[i]);
        for (size_t i : valid_snp) w.row(so[i]);
        printf("Merging output files finished\n");
        (void)k; (void)threads;
    } catch (const std::exception &e) {
        printf("%s\n", e.what());
        if (w.table) fclose(w.table);
        if (w.matrix) fclose(w.matrix);
        return 1;
    }
    if (w.table) fclose(w.table);
    if (w.matrix) fclose(w.matrix);
    return 0;
}

extern "C" int vs_bam_merger_ref_only_main(int argc, char **argv)
{
    if (argc != 10) {
        fprintf(stderr, "USAGE: bam_merger RESULT_MERGED.txt FEATURE_MATRIX.txt RESULT_REF.bam ONTARGETS.bed GENOME.fa TUSCAN_REGRESSION.txt NUMMISMATCHES SEQLENGTH MIT\n");
        return 1;
    }
    unsigned k, seq_len, mit;
    for (int i = 7; i <= 9; ++i) {
        unsigned &dst = i == 7 ? k : i == 8 ? seq_len : mit;
        if (!cast_unsigned(argv[i], dst)) { fprintf(stderr, "ERROR: Cannot cast %s into an unsigned.\n", argv[i]); return 1; }
    }
    Writer w;
    try {
        Genome ref;
        std::string err;
        if (!open_genome(argv[5], ref, err)) throw std::runtime_error("ERROR: Reference index could not be loaded or built.");
        printf("Read reference BAM file\n");
        std::vector<Pot> ro = read_sam(argv[3], ref);
        std::map<std::string, Pot> on;
        std::map<std::string, unsigned> count;
        read_ontargets(argv[4], ref, on, count);
        std::map<std::string, double> activity = read_tuscan(argv[6]);
        w.use_mit = mit == 0; w.with_variants = false; w.on = &on; w.count = &count; w.activity = &activity;
        open_outputs(w, argv[1], mit == 0 ? nullptr : argv[2], seq_len, false);
        for (const Pot &o : ro)
            if (!same(o, on.at(o.target))) w.row(o);
        printf("Writing reference output finished.\n");
        (void)k;
    } catch (const std::exception &e) {
        printf("%s\n", e.what());
        if (w.table) fclose(w.table);
        if (w.matrix) fclose(w.matrix);
        return 1;
    }
    if (w.table) fclose(w.table);
    if (w.matrix) fclose(w.matrix);
    return 0;
}
