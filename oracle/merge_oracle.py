"""ORACLE for row f3: pure-Python restatement of the reference's `bam_merger` and
`bam_merger_ref_only` (test infrastructure, NOT product code).

Follows VARSCOT_pipeline/variant_processing/{bam_merger.cpp, bam_merger_ref_only.cpp, merge_output_bam.h,
filter_output_bam.h, feature_matrix.h, mit_score.h}; each function cites the lines it restates.  PARITY UNPINNED (the
reference needs SeqAn bam_io / seq_io and ships no expected output).

Deviations, on inputs where the reference is undefined:
  M1 mit_score.h:44 indexes matrixM with positions >= 20 when two PAM positions mismatch: such positions contribute a
     factor 1 (as if matrixM were 0 there).
  M2 feature_matrix.h:52-54,97: std::map::operator[] on a pair / mismatch type containing N inserts a new key while its
     size() is used in the same expression: sequences with N cannot be hits; N is folded to A here.
  M3 an on-target without a TUSCAN activity makes the reference terminate (std::map::at): here an error is raised.
"""
from __future__ import annotations

from dataclasses import dataclass, field

from .vcf_oracle import dna5, read_genome

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


@dataclass
class Pot:                      # PotentialOffTarget, filter_output_bam.h:23-32
    chr: str = ""
    target: str = ""
    snpType: str = "REF"
    sequence: str = ""
    mismatchPos: list = field(default_factory=list)
    pos: int = 0
    strand: str = "+"

    def key(self):              # comp(), filter_output_bam.h:40-49
        return (self.target, self.chr, self.pos, self.strand, self.sequence, tuple(self.mismatchPos), self.snpType)


def fetch(genome: dict, name: str, b: int, e: int, strand: str) -> str:
    """extractSequenceFromIndex without flanks (extract_fasta_ontargets.h:30-70)."""
    if name not in genome:
        raise IndexError("ERROR: Index out of range.")
    seq = genome[name]
    b, e = min(max(b, 0), len(seq)), min(max(e, 0), len(seq))
    if b > e:
        e = b
    s = dna5(seq[b:e])
    return "".join(COMP[c] for c in reversed(s)) if strand == "-" else s


def mismatch_positions(md: str) -> list:
    """getMismatchPositions, filter_output_bam.h:330-349: `while (is >> num >> base) pos += num + 1`."""
    out, pos, i, n = [], 0, 0, len(md)
    while True:
        while i < n and md[i] in " \t\n":
            i += 1
        j = i
        while j < n and md[j].isdigit():
            j += 1
        if j == i:
            break                                   # `is >> num` fails
        num = int(md[i:j])
        while j < n and md[j] in " \t\n":
            j += 1
        if j >= n:
            break                                   # `>> base` fails at the end of the string
        pos += num + 1
        out.append(pos - 1)
        i = j + 1
    return out if out else [-1]


def read_sam(path: str, genome: dict) -> list:
    """readBamFile, filter_output_bam.h:362-418 (text SAM, header-less)."""
    out = []
    with open(path) as f:
        for line in f:
            if not line.strip() or line.startswith("@"):
                continue
            fld = line.rstrip("\r\n").split("\t")
            p = Pot(target=fld[0], chr=fld[2], pos=int(fld[3]) - 1, strand="-" if int(fld[1]) & 16 else "+")
            p.sequence = fetch(genome, p.chr.split()[0] if p.chr.split() else p.chr, p.pos, p.pos + 23, p.strand)
            md = ""
            for t in fld[11:]:
                if t.startswith("MD:Z:"):
                    md = t[5:]
            p.mismatchPos = mismatch_positions(md)
            out.append(p)
    return out


def read_ontargets(bed: str, genome: dict):
    """readOntargets, filter_output_bam.h:449-483: first record of a name wins (std::map::insert)."""
    on, count = {}, {}
    with open(bed) as f:
        for line in f:
            line = line.rstrip("\r\n")
            if not line or line.startswith(("#", "track", "browser")):
                continue
            fld = line.split("\t")
            if len(fld) < 6:
                continue
            p = Pot(target=fld[3], chr=fld[0], pos=int(fld[1]), strand=fld[5][:1], mismatchPos=[-1])
            p.sequence = fetch(genome, p.chr, p.pos, p.pos + 23, p.strand)
            on.setdefault(p.target, p)
            count.setdefault(p.target, 0)
    return on, count


def read_snp_fasta(path: str):
    """readRecords(ids, seqs) + getSnpInfoTable, filter_output_bam.h:425-439: id split on '_' + sequence length."""
    ids, lens, name, n = [], [], None, 0
    with open(path) as f:
        for line in f:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    ids.append(name); lens.append(n)
                name, n = line[1:], 0
            elif name is not None:
                n += len("".join(line.split()))
    if name is not None:
        ids.append(name); lens.append(n)
    table = [i.split("_") + [str(l)] for i, l in zip(ids, lens)]
    chr_map = {}
    for t in table:
        chr_map.setdefault(t[0], len(chr_map))
    return chr_map, table


def filter_ref(off: list, on: dict, chr_map: dict, table: list, seq_len: int) -> list:
    """filterRefAlignment, filter_output_bam.h:70-124."""
    by_chr = {}
    for t in table:
        by_chr.setdefault(t[0], []).append(t)
    valid = []
    for i, o in enumerate(off):
        ok = o.key() != on[o.target].key()
        if ok and o.chr in chr_map:
            for t in by_chr.get(o.chr, []):
                if o.pos >= int(t[1]) and o.pos + seq_len <= int(t[1]) + int(t[-1]):
                    ok = False
                    break
        if ok:
            valid.append(i)
    return valid


def snp_type(fid: list, pos: int, seq_len: int):
    """getSnpType, filter_output_bam.h:189-263. Returns (snpType or None, adjusted position)."""
    variants, count, start_found = [], 0, False
    for i in range(3, len(fid) - 2, 3):
        p, r, a = int(fid[i]), fid[i + 1], fid[i + 2]
        if len(r) == len(a):
            if pos <= p and pos + seq_len > p:
                variants.append(fid[i]); start_found = True
        elif len(r) < len(a):
            if (pos <= p + 1 and pos + seq_len > p + 1) or (pos <= p + len(a) - 1 and pos + seq_len > p + len(a) - 1):
                variants.append(fid[i]); start_found = True
            elif not start_found:
                count -= len(a) - len(r)
        else:
            if (pos <= p + 1 and pos + seq_len > p + 1) or (pos <= p + len(r) - 1 and pos + seq_len > p + len(r) - 1):
                variants.append(fid[i]); start_found = True
            elif not start_found:
                count += len(r) - len(a)
    st = ("VAR_" + fid[0] + "_" + ",".join(variants)) if variants else None
    return st, pos + count


def filter_snp(off: list, on: dict, seq_len: int) -> list:
    """filterSnpAlignment, filter_output_bam.h:279-317 (modifies the records in place)."""
    valid = []
    for i, o in enumerate(off):
        fid = o.chr.split("_")
        o.chr = fid[0]
        o.pos = o.pos + int(fid[1])
        st, o.pos = snp_type(fid, o.pos, seq_len)
        if st is not None:
            o.snpType = st
        ok = o.key() != on[o.target].key()
        if i > 0 and o.key() == off[i - 1].key():
            ok = False
        if ok:
            valid.append(i)
    return valid


MATRIX_M = [0, 0, 0.014, 0, 0, 0.395, 0.317, 0, 0.389, 0.079, 0.445, 0.508, 0.613, 0.851, 0.732, 0.828, 0.615, 0.804, 0.685, 0.583]


def mit_score(mm: list) -> float:
    """calcMitScore, mit_score.h:12-68."""
    if mm == [-1]:
        return 100.0
    nm = len(mm) if mm[-1] < 20 else len(mm) - 1
    if nm == 0:
        return 100.0
    s3 = 1.0 / float(nm ** 2)
    s1, dist = 1.0, []
    for i in range(nm):
        s1 *= 1 - (MATRIX_M[mm[i]] if 0 <= mm[i] < 20 else 0.0)        # M1
        if i > 0:
            dist.append(mm[i] - mm[i - 1])
    if nm < 2:
        s2 = 1.0
    else:
        avg = float(sum(dist)) / float(len(dist))
        s2 = 1 / (((19 - avg) / 19) * 4 + 1)
    return s1 * s2 * s3 * 100


def fmt_double(x: float) -> str:
    """std::ostream << double with default precision (6 significant digits, %g)."""
    return "%g" % x


PAIRS = ["AA", "AC", "AG", "AT", "CA", "CC", "CG", "CT", "GA", "GC", "GG", "GT", "TA", "TC", "TG", "TT"]
MTYPES = ["AC", "AG", "AT", "CA", "CG", "CT", "GA", "GC", "GT", "TA", "TC", "TG"]
TRANSITIONS = {"AG", "CT", "GA", "TC"}


def feature_record(on_t: str, off_t: str) -> list:
    """featureMatrixRecord, feature_matrix.h:25-126."""
    f = [0] * 442
    fold = lambda c: c if c in "ACGT" else "A"                           # M2
    prec = False
    for i in range(len(off_t) - 2):
        if i < 19:
            pr = fold(off_t[i]) + fold(off_t[i + 1])
            f[120 + i * 16 + PAIRS.index(pr)] = 1
            f[424 + PAIRS.index(pr)] += 1
        f[36 + i * 4 + "ACGT".index(fold(off_t[i]))] = 1
        if on_t[i] != off_t[i]:
            f[0] += 1
            f[i + 1] = 1
            if 7 < i < 20:
                f[441] += 1
            if prec:
                f[440] += 1
            prec = True
            mt = fold(on_t[i]) + fold(off_t[i])
            if mt in TRANSITIONS:
                f[34] += 1
            else:
                f[35] += 1
            if mt in MTYPES:
                f[22 + MTYPES.index(mt)] = 1
            else:
                f[22] = 1                                               # same letter after folding N: operator[] default 0
        else:
            prec = False
    return f


def feature_names(seq_len: int = 23) -> list:
    """getFeatureNames, feature_matrix.h:140-204."""
    n = [""] * 443
    n[0] = "totalMismatches"
    for i in range(1, seq_len - 1):
        n[i] = "mismatchPos" + str(i)
    for i, t in enumerate(["AtoC", "AtoG", "AtoT", "CtoA", "CtoG", "CtoT", "GtoA", "GtoC", "GtoT", "TtoA", "TtoC", "TtoG"]):
        n[22 + i] = t
    n[34], n[35] = "transitionNumber", "transversionNumber"
    for i in range(1, seq_len - 2):
        for j, c in enumerate("ACGT"):
            n[36 + (i - 1) * 4 + j] = c + str(i)
    n[116:120] = ["PAMA", "PAMC", "PAMG", "PAMT"]
    for i in range(1, seq_len - 3):
        for j, p in enumerate(PAIRS):
            n[120 + (i - 1) * 16 + j] = p + str(i)
    for j, p in enumerate(PAIRS):
        n[424 + j] = p
    n[440], n[441], n[442] = "adjacentMismatches", "seedMismatches", "ontargetActivity"
    return n


def read_tuscan(path: str) -> dict:
    """readTuscanResult, feature_matrix.h:206-230: `is >> target >> sequence >> score`; first entry of a name wins."""
    out = {}
    with open(path) as f:
        for line in f:
            p = line.split()
            if len(p) >= 3:
                try:
                    out.setdefault(p[0], float(p[2]))
                except ValueError:
                    pass
    return out


def _rows(records, idx, on, count, activity, use_mit, with_variants, fm_rows):
    rows = []
    for i in idx:
        o = records[i]
        count[o.target] += 1
        name = f"{o.target}_{count[o.target]}"
        score = fmt_double(mit_score(o.mismatchPos)) if use_mit else "."
        if o.mismatchPos == [-1]:
            mmcol = "0\t"
        else:
            mmcol = f"{len(o.mismatchPos)}\t" + ",".join(str(x) for x in o.mismatchPos)
        row = f"{o.chr}\t{o.pos}\t{o.pos + 23}\t{name}\t{score}\t{o.strand}\t{o.sequence}\t{mmcol}"
        rows.append(row + ("\t" + o.snpType if with_variants else "") + "\n")
        if fm_rows is not None:
            if o.target not in activity:
                raise KeyError(f"no on-target activity for {o.target}")                     # M3
            feats = feature_record(on[o.target].sequence, o.sequence)
            fm_rows.append(name + "\t" + "".join(str(x) + "\t" for x in feats) + fmt_double(activity[o.target]) + "\n")
    return rows


HEADER = "#Chr\tStart\tEnd\tTargetsite\tScore\tStrand\tSequence\tMismatch_Number\tMismatch_Positions"


def bam_merger(ref_sam, snp_sam, bed, genome_fa, snp_fa, tuscan, seq_len=23, mit=0):
    """mergeResults, merge_output_bam.h:46-215 (mit == 0) / :244-460 (feature-matrix mode).  Returns (table text, matrix text or None)."""
    genome, snp_genome = read_genome(genome_fa), read_genome(snp_fa)
    chr_map, table = read_snp_fasta(snp_fa)
    on, count = read_ontargets(bed, genome)
    ref = read_sam(ref_sam, genome)
    v_ref = filter_ref(ref, on, chr_map, table, seq_len)
    snp = read_sam(snp_sam, snp_genome)
    v_snp = filter_snp(snp, on, seq_len)
    activity = read_tuscan(tuscan)
    fm = None if mit == 0 else []
    rows = _rows(ref, v_ref, on, count, activity, mit == 0, True, fm) + _rows(snp, v_snp, on, count, activity, mit == 0, True, fm)
    text = HEADER + "\tVariants\n" + "".join(rows)
    matrix = None if fm is None else "\t".join(feature_names(seq_len)) + "\n" + "".join(fm)
    return text, matrix


def bam_merger_ref_only(ref_sam, bed, genome_fa, tuscan, seq_len=23, mit=0):
    """processRefOnly, merge_output_bam.h:462-722: on-target removal only, no Variants column."""
    genome = read_genome(genome_fa)
    ref = read_sam(ref_sam, genome)
    on, count = read_ontargets(bed, genome)
    activity = read_tuscan(tuscan)
    idx = [i for i, o in enumerate(ref) if o.key() != on[o.target].key()]
    fm = None if mit == 0 else []
    rows = _rows(ref, idx, on, count, activity, mit == 0, False, fm)
    text = HEADER + "\n" + "".join(rows)
    matrix = None if fm is None else "\t".join(feature_names(seq_len)) + "\n" + "".join(fm)
    return text, matrix
