"""Row f3: the `bam_merger` / `bam_merger_ref_only` drop-ins against the Python oracle
(oracle/merge_oracle.py) and hand-derived vectors.  CPU only: the SAM inputs come from the ORACLE mapper here, so this
test needs no GPU (the GPU chain test feeds the same binaries from the CUDA mapper).  Parity is unpinned."""
import os
import subprocess

import numpy as np
import pytest

from oracle import merge_oracle as MO
from oracle import vcf_oracle as VO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VP = os.path.join(ROOT, "build", "variant_processing_build")
ORACLE_MAPPER = os.path.join(ROOT, "oracle", "oracle_bidir_mapping")
LUT = "ACGT"


def test_md_parser_reproduces_the_reference_quirk():
    # filter_output_bam.h:338-342: `while (is >> num >> base)`: exact with samtools-style MD, under-counts SeqAn-style MD
    assert MO.mismatch_positions("23") == [-1]
    assert MO.mismatch_positions("5A17") == [5]
    assert MO.mismatch_positions("22A") == [22]                 # SeqAn style, trailing mismatch: num = 22, base = A
    assert MO.mismatch_positions("0C0T21") == [0, 1]
    assert MO.mismatch_positions("CT21") == [-1]                # SeqAn style, leading mismatch: nothing parsed
    assert MO.mismatch_positions("5AC16") == [5]                # SeqAn style, adjacent mismatches: the second is lost
    assert MO.mismatch_positions("3G4T13") == [3, 8]


def test_mit_score_vectors():
    # mit_score.h: 100 for a perfect match or a single PAM mismatch; hand-computed otherwise
    assert MO.mit_score([-1]) == 100 and MO.mit_score([21]) == 100
    assert MO.fmt_double(MO.mit_score([19])) == "%g" % ((1 - 0.583) * 100)
    s = (1 - 0.395) * (1 - 0.445) * (1 / (((19 - 5) / 19) * 4 + 1)) * (1 / 4) * 100
    assert abs(MO.mit_score([5, 10]) - s) < 1e-12
    # PAM mismatch last: excluded from nm
    assert abs(MO.mit_score([5, 10, 22]) - s) < 1e-12
    assert MO.fmt_double(100.0) == "100" and MO.fmt_double(0.000012345678) == "1.23457e-05"


def test_feature_record_vectors():
    on = "GAGTCCGAGCAGAAGAAGAAGGG"
    f = MO.feature_record(on, on)
    assert f[0] == 0 and sum(f[1:36]) == 0 and sum(f[36:120]) == 21 and sum(f[120:424]) == 19 and sum(f[424:440]) == 19
    off = "GAGTTCGAGCAGAAGAAGAAGGG"                              # C -> T at position 4 (transition)
    f = MO.feature_record(on, off)
    assert f[0] == 1 and f[5] == 1 and f[34] == 1 and f[35] == 0 and f[22 + MO.MTYPES.index("CT")] == 1 and f[441] == 0
    off = "GAGTCCGAGTTGAAGAAGAAGGG"                              # positions 9, 10: adjacent, in the seed, C->T, A->T
    f = MO.feature_record(on, off)
    assert f[0] == 2 and f[440] == 1 and f[441] == 2 and f[34] == 1 and f[35] == 1
    names = MO.feature_names(23)
    assert len(names) == 443 and all(names) and names[1] == "mismatchPos1" and names[21] == "mismatchPos21"
    assert names[36] == "A1" and names[115] == "T20" and names[120] == "AA1" and names[423] == "TT19" and names[442] == "ontargetActivity"


def build_case(tmp_path, seed=5):
    """Genome with on-targets, near copies (some repaired by variants), a VCF; maps with the ORACLE mapper."""
    rng = np.random.default_rng(seed)
    seq = list("".join(rng.choice(list(LUT), 80000)))
    bed_rows, vcf_rows = [], []
    for gi, p in enumerate([1000, 9000, 20000, 41000, 60000]):
        seq[p + 21:p + 23] = "GG"
        bed_rows.append(("chr1", p, p + 23, f"guide{gi}", 0, "+"))
        for copy, offs in enumerate(([3, 12], [0, 1], [7], [5, 22])):
            q = p + 1500 * (copy + 1)
            w = seq[p:p + 23]
            for off in offs:
                w[off] = LUT[(LUT.index(w[off]) + 1) % 4]
            if offs == [5, 22]:
                w[22] = "A"                                                   # GA PAM: a PAM-position mismatch
            seq[q:q + 23] = w
            if copy == 0:
                vcf_rows.append((q + 12, seq[q + 12], seq[p + 12], "0|1" if gi % 2 else "1|1"))
            if copy == 2:
                vcf_rows.append((q - 5, "".join(seq[q - 5:q - 2]), seq[q - 5], "0|1"))   # deletion left of the site
    for p in (500, 510, 15000, 33333):
        vcf_rows.append((p, seq[p], LUT[(LUT.index(seq[p]) + 2) % 4], "0|1"))
    vcf_rows.sort()
    g, bed, vcf = str(tmp_path / "genome.fa"), str(tmp_path / "t.bed"), str(tmp_path / "v.vcf")
    s = "".join(seq)
    with open(g, "w") as f:
        f.write(">chr1\n" + "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n")
    open(bed, "w").write("".join("\t".join(map(str, r)) + "\n" for r in bed_rows))
    open(vcf, "w").write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n" +
                         "".join(f"chr1\t{p + 1}\t.\t{r}\t{a}\t.\t.\t.\tGT\t{gt}\n" for p, r, a, gt in vcf_rows))
    guides, flank, snp = str(tmp_path / "guides.fa"), str(tmp_path / "flank.fa"), str(tmp_path / "snp.fa")
    assert subprocess.run([os.path.join(VP, "fasta_writer"), guides, flank, bed, g]).returncode == 0
    assert subprocess.run([os.path.join(VP, "vcf_loader"), vcf, snp, g, "0", "23", "1"], capture_output=True).returncode == 0
    tus = str(tmp_path / "activity.txt")
    open(tus, "w").write("ID  Sequence  Score  Dir\n" + "".join(f"guide{i}  {'A' * 30}  {0.5 + i * 0.37:.11f}  +\n" for i in range(5)))
    return g, bed, snp, guides, tus


@pytest.mark.parametrize("md_style", ["seqan", "samtools"])
@pytest.mark.parametrize("mit", [0, 1])
def test_bam_merger_equals_oracle(tmp_path, md_style, mit):
    g, bed, snp, guides, tus = build_case(tmp_path)
    ref_sam, snp_sam = str(tmp_path / "ref.sam"), str(tmp_path / "snp.sam")
    for fa, sam in ((g, ref_sam), (snp, snp_sam)):
        r = subprocess.run([ORACLE_MAPPER, "-G", fa, "-R", guides, "-M", "4", "-O", sam, "--md-style", md_style], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    out, fm = str(tmp_path / "out.txt"), str(tmp_path / "fm.txt")
    r = subprocess.run([os.path.join(VP, "bam_merger"), out, fm, ref_sam, snp_sam, bed, g, snp, tus, "4", "23", "2", str(mit)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Merging output files finished" in r.stdout
    exp_text, exp_fm = MO.bam_merger(ref_sam, snp_sam, bed, g, snp, tus, 23, mit)
    got = open(out).read()
    assert got == exp_text
    lines = got.splitlines()
    assert lines[0] == "#Chr\tStart\tEnd\tTargetsite\tScore\tStrand\tSequence\tMismatch_Number\tMismatch_Positions\tVariants"
    assert len(lines) > 10 and all(len(l.split("\t")) == 10 for l in lines)
    assert any("\tVAR_chr1_" in l for l in lines) and any(l.endswith("\tREF") for l in lines)
    # the on-targets themselves are gone: no row is a perfect match at an on-target start
    assert not any(l.split("\t")[1] in ("1000", "9000", "20000", "41000", "60000") and l.split("\t")[7] == "0" and l.endswith("REF") for l in lines[1:])
    if mit:
        assert open(fm).read() == exp_fm
        assert all(l.split("\t")[4] == "." for l in lines[1:])
        assert all(len(l.split("\t")) == 444 for l in open(fm).read().splitlines()[1:])
    else:
        assert all(l.split("\t")[4] != "." for l in lines[1:])
    if md_style == "samtools":
        # with exact MD parsing the mismatch-number column equals the count of listed positions
        for l in lines[1:]:
            c = l.split("\t")
            assert int(c[7]) == (len(c[8].split(",")) if c[8] else 0)


@pytest.mark.parametrize("mit", [0, 1])
def test_bam_merger_ref_only_equals_oracle(tmp_path, mit):
    g, bed, snp, guides, tus = build_case(tmp_path, seed=6)
    ref_sam = str(tmp_path / "ref.sam")
    assert subprocess.run([ORACLE_MAPPER, "-G", g, "-R", guides, "-M", "4", "-O", ref_sam, "--md-style", "samtools"], capture_output=True).returncode == 0
    out, fm = str(tmp_path / "out.txt"), str(tmp_path / "fm.txt")
    r = subprocess.run([os.path.join(VP, "bam_merger_ref_only"), out, fm, ref_sam, bed, g, tus, "4", "23", str(mit)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    exp_text, exp_fm = MO.bam_merger_ref_only(ref_sam, bed, g, tus, 23, mit)
    assert open(out).read() == exp_text
    assert open(out).readline() == "#Chr\tStart\tEnd\tTargetsite\tScore\tStrand\tSequence\tMismatch_Number\tMismatch_Positions\n"
    if mit:
        assert open(fm).read() == exp_fm


def test_ref_hits_inside_overlapping_segments(tmp_path):
    """filterRefAlignment (filter_output_bam.h:99-108) drops a reference hit that lies wholly inside ANY variant segment of
    its chromosome.  The product answers that with a binary search over the segments sorted by start + a running maximum of
    their ends; the oracle keeps the reference's scan over all segments.  Nested, overlapping and merely adjacent segments —
    the nearest segment by start often does not cover the hit while an earlier, longer one does."""
    g, bed, snp, guides, tus = build_case(tmp_path, seed=11)
    ref_sam, snp_sam = str(tmp_path / "ref.sam"), str(tmp_path / "snp.sam")
    assert subprocess.run([ORACLE_MAPPER, "-G", g, "-R", guides, "-M", "5", "-O", ref_sam, "--md-style", "samtools"], capture_output=True).returncode == 0
    open(snp_sam, "w").close()
    genome = "".join(l.strip() for l in open(g) if not l.startswith(">"))
    hits = sorted({int(l.split("\t")[3]) - 1 for l in open(ref_sam)})
    assert len(hits) > 20
    rng = np.random.default_rng(3)
    segs = set()
    for i, p in enumerate(hits):
        kind = i % 5
        if kind == 0:
            segs.add((p - 30, 90))                                   # covers
        elif kind == 1:
            segs.add((p + 1, 60)); segs.add((max(0, p - 40), 62))    # starts after the hit / ends one base short
        elif kind == 2:
            segs.add((max(0, p - 400), 900))                         # a long early segment covers ...
            for d in (300, 200, 100, 10, 0):
                segs.add((max(0, p - d), 22))                        # ... the nearer ones are too short
        elif kind == 3:
            segs.add((p, 23))                                        # exactly the window
        # kind 4: no segment near this hit
    for _ in range(300):
        a = int(rng.integers(0, len(genome) - 200))
        segs.add((a, int(rng.integers(1, 120))))
    with open(snp, "w") as f:
        for a, n in sorted(segs, key=lambda x: (rng.random(), x)):   # file order is not sorted
            f.write(f">chr1_{a}_REF\n{genome[a:a + n]}\n")
    if os.path.exists(snp + ".fai"):
        os.remove(snp + ".fai")
    out, fm = str(tmp_path / "out.txt"), str(tmp_path / "fm.txt")
    r = subprocess.run([os.path.join(VP, "bam_merger"), out, fm, ref_sam, snp_sam, bed, g, snp, tus, "5", "23", "1", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    exp_text, _ = MO.bam_merger(ref_sam, snp_sam, bed, g, snp, tus, 23, 0)
    got = open(out).read()
    assert got == exp_text
    kept = {int(l.split("\t")[1]) for l in got.splitlines()[1:]}
    dropped = set(hits) - kept
    assert len(dropped) >= len(hits) // 2 and len(kept) >= len(hits) // 6


def test_usage_errors():
    assert subprocess.run([os.path.join(VP, "bam_merger")], capture_output=True).returncode == 1
    assert subprocess.run([os.path.join(VP, "bam_merger_ref_only"), "a"], capture_output=True).returncode == 1
    r = subprocess.run([os.path.join(VP, "bam_merger")] + ["x"] * 8 + ["4", "23", "z", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot cast z into an unsigned" in r.stderr
