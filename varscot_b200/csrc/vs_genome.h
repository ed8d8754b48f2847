// varscot_b200/csrc/vs_genome.h — host helpers shared by the variant_processing drop-ins (vs_vcf.cpp, vs_merge.cpp):
// SeqAn's Dna5 conversion, string splitting and FAI-style random access to a FASTA (write_fasta.h:245-271,435-448;
// extract_fasta_ontargets.h:30-70,101-112: open the .fai, else build it and save it next to the FASTA).
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <map>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

namespace vsg {

inline std::string dna5(const std::string &s)
{
    std::string o(s.size(), 'N');
    for (size_t i = 0; i < s.size(); ++i) {
        char u = (char)std::toupper((unsigned char)s[i]);
        o[i] = (u == 'A' || u == 'C' || u == 'G' || u == 'T') ? u : (u == 'U' ? 'T' : 'N');
    }
    return o;
}

inline std::vector<std::string> split(const std::string &s, char c)
{
    std::vector<std::string> out;
    size_t b = 0;
    for (;;) {
        size_t e = s.find(c, b);
        out.push_back(s.substr(b, e == std::string::npos ? std::string::npos : e - b));
        if (e == std::string::npos) break;
        b = e + 1;
    }
    return out;
}

// ---- FAI access to the genome (write_fasta.h:435-448: open the .fai, else build and save it) -------------------
struct FaiEntry { std::string name; long length, offset, line_bases, line_width; };
struct Genome {
    std::vector<FaiEntry> entries;
    std::map<std::string, size_t> by_name;
    const char *data = nullptr; size_t size = 0; int fd = -1;
    ~Genome() { if (data) munmap((void *)data, size); if (fd >= 0) close(fd); }
};

inline bool load_fai(const std::string &path, Genome &g)
{
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return false;
    char name[4096]; long a, b, c, d;
    while (fscanf(f, "%4095[^\t]\t%ld\t%ld\t%ld\t%ld\n", name, &a, &b, &c, &d) == 5) g.entries.push_back({name, a, b, c, d});
    fclose(f);
    return !g.entries.empty();
}

inline bool build_fai(Genome &g)
{
    FaiEntry cur; bool have = false;
    size_t i = 0;
    while (i < g.size) {
        const char *nl = (const char *)memchr(g.data + i, '\n', g.size - i);
        size_t e = nl ? (size_t)(nl - g.data) : g.size;
        size_t raw = e - i + (nl ? 1 : 0);
        size_t len = e - i;
        if (len && g.data[e - 1] == '\r') --len;
        if (len && g.data[i] == '>') {
            if (have) g.entries.push_back(cur);
            std::string h(g.data + i + 1, len - 1);
            cur = FaiEntry{h.substr(0, h.find_first_of(" \t")), 0, (long)(i + raw), 0, 0};
            have = true;
        } else if (have && len) {
            if (cur.line_bases == 0) { cur.line_bases = (long)len; cur.line_width = (long)raw; }
            cur.length += (long)len;
        }
        i += raw;
    }
    if (have) g.entries.push_back(cur);
    return true;
}

inline bool open_genome(const char *path, Genome &g, std::string &err)
{
    g.fd = open(path, O_RDONLY);
    struct stat st;
    if (g.fd < 0 || fstat(g.fd, &st) != 0) { err = "ERROR: Index could not be loaded or built."; return false; }
    g.size = (size_t)st.st_size;
    if (g.size) {
        void *p = mmap(nullptr, g.size, PROT_READ, MAP_PRIVATE, g.fd, 0);
        if (p == MAP_FAILED) { err = "ERROR: Index could not be loaded or built."; return false; }
        g.data = (const char *)p;
    }
    std::string fai = std::string(path) + ".fai";
    if (!load_fai(fai, g)) {
        build_fai(g);
        if (FILE *f = fopen(fai.c_str(), "w")) {
            for (const FaiEntry &e : g.entries) fprintf(f, "%s\t%ld\t%ld\t%ld\t%ld\n", e.name.c_str(), e.length, e.offset, e.line_bases, e.line_width);
            fclose(f);
        } else { err = "ERROR: Index could not be written do disk."; return false; }
    }
    for (size_t i = 0; i < g.entries.size(); ++i) g.by_name.emplace(g.entries[i].name, i);
    return true;
}

// extractSequenceFromIndex, write_fasta.h:245-271
inline std::string extract(const Genome &g, size_t idx, long b, long e)
{
    const FaiEntry &f = g.entries[idx];
    b = std::min(std::max(b, 0L), f.length); e = std::min(std::max(e, 0L), f.length);
    if (b > e) e = b;
    std::string out;
    out.reserve((size_t)(e - b));
    for (long p = b; p < e; ++p) {
        long off = f.offset + (f.line_bases ? (p / f.line_bases) * f.line_width + p % f.line_bases : p);
        if ((size_t)off >= g.size) break;
        out.push_back(g.data[off]);
    }
    return dna5(out);
}

inline std::string revcomp(std::string s)
{
    std::reverse(s.begin(), s.end());
    for (char &c : s) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    return s;
}

}  // namespace vsg
