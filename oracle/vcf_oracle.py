"""ORACLE for row f1: a pure-Python restatement of the reference's `vcf_loader`
(test infrastructure, NOT product code; small inputs only).

Follows VARSCOT_pipeline/variant_processing/{vcf_loader.cpp, process_vcf.h, overlap_sequences.h, write_fasta.h}
function by function; each function cites the lines it restates.  PARITY UNPINNED: the reference needs SeqAn
(vcf_io, seq_io/FAI) and cannot be built here, and it ships no expected output for this stage.

Deliberate deviations, all on inputs where the reference has undefined behaviour (SURVEY.md 8f "avoid its UB inputs"):
  D1 process_vcf.h:145-151 writes variants[1] after resize(1): here the single variant is kept with allele = 1.
  D2 process_vcf.h:73-83 leaves positionGT uninitialised when FORMAT has no GT: here the record is skipped.
  D3 overlap_sequences.h:114-115 reads maxDeletion[-1]: here that read yields 0.
  D4 overlap_sequences.h:99-103,122-126 index allele j of record i with the CENTER's allele count: all alleles of a
     record share `pos`, so allele 0 is read instead.
  D5 overlap_sequences.h:158 `start = pos - windowSizeLeft + 1` wraps for variants closer than 22 bp to the contig
     start: here start is clamped at 0.
  D6 overlap_sequences.h:233-237 sorts with std::sort (unstable) by pos: here ties keep file order.
"""
from __future__ import annotations

from dataclasses import dataclass, field


def dna5(s: str) -> str:
    """SeqAn Dna5 conversion: ACGT case-insensitive, U -> T, everything else N."""
    out = []
    for c in s:
        u = c.upper()
        out.append(u if u in "ACGT" else ("T" if u == "U" else "N"))
    return "".join(out)


@dataclass
class VariantSequence:            # process_vcf.h:32-42
    ref: str = ""
    alt: str = ""
    chr: int = 0
    pos: int = 0
    start: int = 0
    end: int = 0
    variantType: int = 0          # 0 substitution, 1 insertion, 2 deletion
    allele: int = 0               # 0 first, 1 second, 2 both, -1 unphased


def parse_gt(text: str):
    """`is >> firstAllele [>> sep >> secondAllele]` of process_vcf.h:93-113 (istream integer extraction)."""
    i, n = 0, len(text)

    def read_int(i):
        j = i
        while j < n and text[j] in " \t":
            j += 1
        k = j
        if k < n and text[k] in "+-":
            k += 1
        d = k
        while k < n and text[k].isdigit():
            k += 1
        if k == d:
            return None, i
        return int(text[j:k]), k

    first, i = read_int(i)
    if first is None:
        return None
    if i >= n:
        return first, None, None
    sep = text[i]
    second, j = read_int(i + 1)
    if second is None:
        return first, None, None
    return first, sep, second


def process_record(rec: dict, sample_index: int) -> list:
    """processRecord, process_vcf.h:54-209.  rec: chrom index, pos (0-based), ref, alt, format, samples."""
    vs = VariantSequence(ref=dna5(rec["ref"]), chr=rec["rid"], pos=rec["pos"])
    if sample_index >= len(rec["samples"]):
        raise IndexError("ERROR: Sample index out of range.")                      # :61-64
    entries = rec["samples"][sample_index].split(":")
    fmt = rec["format"].split(":")
    if "GT" not in fmt:
        return []                                                                  # D2
    pos_gt = fmt.index("GT")
    if pos_gt >= len(entries):
        return []
    alts = rec["alt"].split(",")
    gt = parse_gt(entries[pos_gt])
    if gt is None:
        return []                                                                  # :110-113
    first, sep, second = gt
    if first > len(alts):
        return []
    phased = True
    if second is not None and second <= len(alts):
        if sep == "/":
            phased = False                                                         # :99-102
    else:
        second = first                                                             # :104-108 (haploid)
    variants = []
    if first == 0 and second == 0:
        return []                                                                  # :116-120
    if first > 0 and second > 0 and first != second:                               # :121-158
        a, b = alts[first - 1], alts[second - 1]
        if a != "." and b != ".":
            v0 = VariantSequence(**{**vs.__dict__, "allele": 0, "alt": dna5(a)})
            v1 = VariantSequence(**{**vs.__dict__, "allele": 1, "alt": dna5(b)})
            variants = [v0, v1]
        elif a != ".":
            variants = [VariantSequence(**{**vs.__dict__, "allele": 0, "alt": dna5(a)})]
        elif b != ".":
            variants = [VariantSequence(**{**vs.__dict__, "allele": 1, "alt": dna5(b)})]     # D1
        else:
            return []
    else:
        if alts[0] == ".":
            return []                                                              # :160-163
        if first == 0:
            variants = [VariantSequence(**{**vs.__dict__, "allele": 1, "alt": dna5(alts[second - 1])})]
        elif second == 0:
            variants = [VariantSequence(**{**vs.__dict__, "allele": 0, "alt": dna5(alts[first - 1])})]
        else:
            variants = [VariantSequence(**{**vs.__dict__, "allele": 2, "alt": dna5(alts[first - 1])})]
    for v in variants:                                                             # :187-207
        if not phased and first != second:
            v.allele = -1
        if len(v.ref) > len(v.alt):
            v.variantType = 2
        elif len(v.ref) == len(v.alt):
            v.variantType = 0
        else:
            v.variantType = 1
    return variants


def read_vcf(path: str):
    """processVcfFile, process_vcf.h:226-269: records in file order; contig table = ##contig IDs, then first appearance."""
    chr_table, records = [], []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\r\n")
            if line.startswith("##"):
                if line.startswith("##contig=<") and "ID=" in line:
                    cid = line.split("ID=", 1)[1].split(",", 1)[0].rstrip(">")
                    if cid not in chr_table:
                        chr_table.append(cid)
                continue
            if line.startswith("#") or not line:
                continue
            f_ = line.split("\t")
            if len(f_) < 10:
                continue
            if f_[0] not in chr_table:
                chr_table.append(f_[0])
            records.append({"rid": chr_table.index(f_[0]), "pos": int(f_[1]) - 1, "ref": f_[3], "alt": f_[4], "format": f_[8],
                            "samples": f_[9:]})
    return chr_table, records


def find_max_overlap(all_variants, sorted_index, seq_length):
    """findMaxOverlap, overlap_sequences.h:35-162.  Returns (regions [(I1, I2)], center variant indices)."""
    n = len(sorted_index)
    max_del = [0] * n
    for i in range(n):                                                             # :41-52
        for v in all_variants[sorted_index[i]]:
            if v.variantType == 2:
                max_del[i] = max(max_del[i], len(v.ref) - len(v.alt))
    pos = lambda i: all_variants[sorted_index[i]][0].pos
    regions, centers = [], []
    r1 = r2 = 0
    for i in range(n):
        if r2 > i:                                                                 # :68
            idx_right = r2
            wsr = seq_length + max_del[i]
            if idx_right < n:
                for d in range(i + 1, idx_right + 1):
                    wsr += max_del[d]                                              # :78-84
            while idx_right < n and pos(idx_right) - pos(i) < wsr:                 # :86-94
                wsr += max_del[idx_right]
                idx_right += 1
            if idx_right == r2:                                                    # :97-104
                for v in all_variants[centers[-1]]:
                    v.end = pos(i) + wsr                                           # D4
                continue
            r2 = idx_right
            idx_left = i - 1                                                       # :108-116
            wsl = seq_length + max_del[idx_left]
            while idx_left >= 0 and pos(i) - pos(idx_left) < wsl:
                idx_left -= 1
                wsl += max_del[idx_left] if idx_left >= 0 else 0                   # D3
            if idx_left + 1 == r1:                                                 # :120-128
                for v in all_variants[centers[-1]]:
                    v.end = pos(i) + wsr
                regions[-1] = (regions[-1][0], idx_right)
                continue
            r1 = idx_left + 1
        else:                                                                      # :131-151
            wsr = seq_length + max_del[i]
            idx_right = i + 1
            while idx_right < n and pos(idx_right) - pos(i) < wsr:
                wsr += max_del[idx_right]
                idx_right += 1
            r2 = idx_right
            wsl = seq_length
            r1 = i
        regions.append((r1, r2))
        centers.append(sorted_index[i])
        for v in all_variants[sorted_index[i]]:                                    # :156-160
            v.start = max(0, v.pos - wsl + 1)                                      # D5
            v.end = v.pos + wsr
    return regions, centers


def get_fasta_id(all_variants, sorted_index, first, center, combination, chr_name):
    """getFastaID, write_fasta.h:30-65."""
    parts = [chr_name, "_", str(all_variants[center][0].start), "_"]
    if all(c == -1 for c in combination):
        parts.append("REF")
    else:
        parts.append("ALT")
        for i, c in enumerate(combination):
            if c != -1:
                v = all_variants[sorted_index[first + i]][c]
                parts += ["_", str(v.pos), "_", v.ref, "_", v.alt]
    return "".join(parts)


def all_combinations(all_variants, sorted_index, first, last, center, chr_name):
    """allCombinations, write_fasta.h:88-229.  Returns ([allele string lists], [ids])."""
    size = last - first
    unphased = []
    fa, sa = [""] * size, [""] * size
    ifa, isa = [0] * size, [0] * size
    for i in range(first, last):
        rec = all_variants[sorted_index[i]]
        j = i - first
        if rec[0].allele == -1:
            unphased.append(j)
        elif len(rec) == 2:
            fa[j], ifa[j], sa[j], isa[j] = rec[0].alt, 0, rec[1].alt, 1
        elif rec[0].allele == 0:
            fa[j], ifa[j], sa[j], isa[j] = rec[0].alt, 0, rec[0].ref, -1
        elif rec[0].allele == 1:
            fa[j], ifa[j], sa[j], isa[j] = rec[0].ref, -1, rec[0].alt, 0
        else:
            fa[j], ifa[j], sa[j] = rec[0].alt, 0, rec[0].alt        # :140-146: indexVariantsSecond stays 0
    combos, ids = [], []

    def emit():
        combos.append(list(fa))
        ids.append(get_fasta_id(all_variants, sorted_index, first, center, ifa, chr_name))
        if ifa != isa:
            combos.append(list(sa))
            ids.append(get_fasta_id(all_variants, sorted_index, first, center, isa, chr_name))

    if unphased:
        u = len(unphased)
        for t in range(2 ** u):                                      # the stack of :155-214 enumerates tuples in lexicographic order
            tup = [(t >> (u - 1 - b)) & 1 for b in range(u)]
            for b, j in enumerate(unphased):
                rec = all_variants[sorted_index[first + j]]
                if len(rec) == 2:
                    fa[j] = sa[j] = rec[tup[b]].alt
                    ifa[j] = isa[j] = tup[b]
                elif tup[b] == 0:
                    fa[j] = sa[j] = rec[0].ref
                    ifa[j] = isa[j] = -1
                else:
                    fa[j] = sa[j] = rec[0].alt
                    ifa[j] = isa[j] = 0
            emit()
    else:
        emit()
    return combos, ids


def generate_variant_sequences(genome, all_variants, sorted_index, chr_name, region, center):
    """generateVariantSequences, write_fasta.h:303-399, normal case (startVariant/endVariant cannot occur with D5)."""
    r1, r2 = region
    seq = genome[chr_name]

    def extract(b, e):                                               # extractSequenceFromIndex, :245-271
        L = len(seq)
        b, e = min(b, L), min(e, L)
        if b > e:
            e = b
        return dna5(seq[b:e])

    c = all_variants[center][0]
    base = []
    for i in range(r1, r2 + 1):
        if i == r1:
            b, e = c.start, all_variants[sorted_index[i]][0].pos
        elif i == r2:
            p = all_variants[sorted_index[i - 1]][0]
            b, e = p.pos + len(p.ref), c.end
        else:
            p = all_variants[sorted_index[i - 1]][0]
            b, e = p.pos + len(p.ref), all_variants[sorted_index[i]][0].pos
        base.append(extract(b, e))
    combos, ids = all_combinations(all_variants, sorted_index, r1, r2, center, chr_name)
    seqs = []
    for alle in combos:
        s = []
        for j in range(len(alle)):
            s.append(base[j]); s.append(alle[j])
        s.append(base[len(alle)])
        seqs.append("".join(s))
    return seqs, ids


def read_genome(path: str) -> dict:
    """FAI view of the genome: name = first word of the header."""
    g, name, parts = {}, None, []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    g[name] = "".join(parts)
                name, parts = line[1:].split()[0] if line[1:].split() else "", []
            elif name is not None:
                parts.append("".join(line.split()))
    if name is not None:
        g[name] = "".join(parts)
    return g


def vcf_loader(vcf_path: str, genome_path: str, sample: int = 0, seq_length: int = 23):
    """main of vcf_loader.cpp:11-77.  Returns [(id, sequence)] in output order."""
    chr_table, records = read_vcf(vcf_path)
    all_variants = []
    for r in records:
        v = process_record(r, sample)
        if v:
            all_variants.append(v)
    genome = read_genome(genome_path)
    out = []
    for ci, cname in enumerate(chr_table):                           # getVariantOverlapRanges, overlap_sequences.h:183-240
        idx = [i for i, v in enumerate(all_variants) if v[0].chr == ci]
        idx.sort(key=lambda i: all_variants[i][0].pos)               # D6 (stable)
        regions, centers = find_max_overlap(all_variants, idx, seq_length)
        if regions and cname not in genome:
            raise IndexError("ERROR: Index out of range.")           # write_fasta.h:249-252
        for reg, cen in zip(regions, centers):                       # writeFastaFile, write_fasta.h:451-463
            seqs, ids = generate_variant_sequences(genome, all_variants, idx, cname, reg, cen)
            out += list(zip(ids, seqs))
    return out


def write_fasta(path: str, records, width: int = 70):
    """SeqAn writeRecords: sequence lines wrapped at 70 columns."""
    with open(path, "w") as f:
        for rid, seq in records:
            f.write(">" + rid + "\n")
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + "\n")
            if not seq:
                f.write("\n")


def fasta_writer(bed_path: str, genome_path: str, flanking: bool):
    """writeFastaOntargets, extract_fasta_ontargets.h:86-138 (+ extractSequenceFromIndex :30-70): [(name, sequence)]."""
    genome = read_genome(genome_path)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    out = []
    with open(bed_path) as f:
        for line in f:
            line = line.rstrip("\r\n")
            if not line or line.startswith(("#", "track", "browser")):
                continue
            fld = line.split("\t")
            if len(fld) < 6:
                continue
            if fld[0] not in genome:
                raise IndexError("ERROR: Index out of range.")
            seq = genome[fld[0]]
            b, e, strand = int(fld[1]), int(fld[2]), fld[5][:1]
            if flanking and strand == "+":
                b, e = b - 4, e + 3
            elif flanking and strand == "-":
                b, e = b - 3, e + 4
            b, e = min(max(b, 0), len(seq)), min(max(e, 0), len(seq))
            if b > e:
                e = b
            s_ = dna5(seq[b:e])
            if strand == "-":
                s_ = "".join(comp[c] for c in reversed(s_))
            out.append((fld[3], s_))
    return out
