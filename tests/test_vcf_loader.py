"""Row f1: the `vcf_loader` drop-in against the Python oracle (oracle/vcf_oracle.py, a function-by-function
restatement of process_vcf.h / overlap_sequences.h / write_fasta.h) and against hand-derived vectors (appendix A).
CPU only.  Parity is unpinned: the reference needs SeqAn and ships no expected output for this stage."""
import os
import subprocess

import numpy as np
import pytest

from oracle import vcf_oracle as VO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "variant_processing_build", "vcf_loader")
HDR = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n"


def write_genome(path, contigs, width=60):
    with open(path, "w") as f:
        for name, seq in contigs:
            f.write(f">{name} description\n")
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + "\n")


def run(tmp_path, genome, vcf_body, header=HDR, sample=0):
    g, v, o = str(tmp_path / "g.fa"), str(tmp_path / "in.vcf"), str(tmp_path / "snp.fa")
    write_genome(g, genome)
    open(v, "w").write(header + vcf_body)
    r = subprocess.run([EXE, v, o, g, str(sample), "23", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Process records" in r.stdout and "Write fasta" in r.stdout
    got = open(o).read()
    exp_records = VO.vcf_loader(v, g, sample, 23)
    exp_path = str(tmp_path / "exp.fa")
    VO.write_fasta(exp_path, exp_records)
    assert got == open(exp_path).read()
    return exp_records


def rand_seq(rng, n):
    return "".join(rng.choice(list("ACGT"), n))


def test_isolated_snv_phased_het_gives_ref_and_alt_45mers(tmp_path):
    rng = np.random.default_rng(1)
    seq = rand_seq(rng, 500)
    p = 200                                   # 0-based position of the variant
    alt = "A" if seq[p] != "A" else "C"
    recs = run(tmp_path, [("chr1", seq)], f"chr1\t{p + 1}\t.\t{seq[p]}\t{alt}\t.\t.\t.\tGT\t0|1\n")
    # appendix A.2/A.3: region [p-22, p+23), 45 bp, variant at offset 22; allele code 1 -> first haplotype REF, second ALT
    assert [r[0] for r in recs] == [f"chr1_{p - 22}_REF", f"chr1_{p - 22}_ALT_{p}_{seq[p]}_{alt}"]
    assert recs[0][1] == seq[p - 22:p + 23]
    assert recs[1][1] == seq[p - 22:p] + alt + seq[p + 1:p + 23]
    assert len(recs[0][1]) == 45
    # 1|0 swaps the emission order
    recs = run(tmp_path, [("chr1", seq)], f"chr1\t{p + 1}\t.\t{seq[p]}\t{alt}\t.\t.\t.\tGT:DP\t1|0:7\n")
    assert [r[0].split("_")[2] for r in recs] == ["ALT", "REF"]


def test_hom_alt_and_unphased_and_indels(tmp_path):
    rng = np.random.default_rng(2)
    seq = rand_seq(rng, 2000)
    body = ""
    body += f"chr1\t{300 + 1}\t.\t{seq[300]}\tG{seq[300]}\t.\t.\t.\tGT\t1|1\n"                       # hom-alt insertion (1 bp)
    body += f"chr1\t{600 + 1}\t.\t{seq[600:604]}\t{seq[600]}\t.\t.\t.\tGT\t0|1\n"                    # het deletion of 3 bp
    body += f"chr1\t{900 + 1}\t.\t{seq[900]}\t{'T' if seq[900] != 'T' else 'A'}\t.\t.\t.\tGT\t0/1\n"  # unphased het SNV
    body += f"chr1\t{1200 + 1}\t.\t{seq[1200]}\tAC,AG\t.\t.\t.\tGT\t1|2\n"                           # two ALT alleles
    recs = run(tmp_path, [("chr1", seq)], body)
    ids = [r[0] for r in recs]
    lens = [len(r[1]) for r in recs]
    # hom-alt: one contig (A.3); insertion: ALT is 45 + 1
    assert ids[0].startswith("chr1_278_ALT_300_") and lens[0] == 46
    # deletion of d = 3: REF haplotype 45 + 3, ALT 45 (A.2)
    assert ids[1] == "chr1_578_REF" and lens[1] == 48 and ids[2].startswith("chr1_578_ALT_600_") and lens[2] == 45
    # unphased het: tuples (ref, ref) then (alt, alt): REF then ALT
    assert ids[3] == "chr1_878_REF" and ids[4].startswith("chr1_878_ALT_900_")
    # 1|2: first haplotype ALT[0], second ALT[1], both 46 long
    assert ids[5].endswith("_AC") and ids[6].endswith("_AG") and lens[5] == lens[6] == 46
    assert len(recs) == 7


def test_neighbouring_variants_merge_into_one_region(tmp_path):
    rng = np.random.default_rng(3)
    seq = rand_seq(rng, 1000)

    def snv(p, gt):
        a = "A" if seq[p] != "A" else "C"
        return f"chr1\t{p + 1}\t.\t{seq[p]}\t{a}\t.\t.\t.\tGT\t{gt}\n", a

    l1, a1 = snv(400, "0|1")
    l2, a2 = snv(410, "1|0")
    recs = run(tmp_path, [("chr1", seq)], l1 + l2)
    # two records < 23 bp apart: one region [p0-22, p1+23) (A.2), two phased haplotypes
    assert len(recs) == 2 and all(len(r[1]) == 410 + 23 - (400 - 22) for r in recs)
    assert recs[0][0] == f"chr1_378_ALT_410_{seq[410]}_{a2}" and recs[1][0] == f"chr1_378_ALT_400_{seq[400]}_{a1}"
    assert recs[0][1] == seq[378:410] + a2 + seq[411:433]
    assert recs[1][1] == seq[378:400] + a1 + seq[401:433]
    # far apart: two independent regions
    l3, _ = snv(700, "0|1")
    recs = run(tmp_path, [("chr1", seq)], l1 + l3)
    assert [r[0].split("_")[1] for r in recs] == ["378", "378", "678", "678"]


def test_skipped_records_contig_order_and_multi_sample(tmp_path):
    rng = np.random.default_rng(4)
    g = [("chrA", rand_seq(rng, 400)), ("chrB", rand_seq(rng, 400))]
    hdr = "##fileformat=VCFv4.2\n##contig=<ID=chrB,length=400>\n##contig=<ID=chrA,length=400>\n" \
          "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\n"
    a, b = g[0][1], g[1][1]
    body = ""
    body += f"chrA\t101\t.\t{a[100]}\t{'G' if a[100] != 'G' else 'T'}\t.\t.\t.\tGT\t0|0\t0|1\n"      # sample 0 is 0|0 -> skipped
    body += f"chrA\t201\t.\t{a[200]}\t.\t.\t.\t.\tGT\t0|1\t0|1\n"                                   # ALT '.' -> skipped
    body += f"chrA\t251\t.\t{a[250]}\t{'G' if a[250] != 'G' else 'T'}\t.\t.\t.\tDP\t5\t6\n"         # no GT -> skipped
    body += f"chrB\t151\t.\t{b[150]}\t{'G' if b[150] != 'G' else 'T'}\t.\t.\t.\tGT\t1\t.\n"        # haploid GT: 1 -> hom
    body += f"chrA\t301\t.\t{a[300]}\t{'G' if a[300] != 'G' else 'T'}\t.\t.\t.\tGT\t.|.\t1|1\n"    # unparsable first allele
    recs0 = run(tmp_path, g, body, header=hdr, sample=0)
    assert [r[0].split("_")[0] for r in recs0] == ["chrB"] and recs0[0][0].split("_")[2] == "ALT"
    recs1 = run(tmp_path, g, body, header=hdr, sample=1)
    # sample 1: chrA@100 (het), chrA@300 (hom); contig table order comes from the ##contig lines: chrB first (empty), then chrA
    assert [r[0].split("_")[0:3:2] for r in recs1] == [["chrA", "REF"], ["chrA", "ALT"], ["chrA", "ALT"]]


@pytest.mark.parametrize("seed", range(8))
def test_random_vcfs_product_equals_oracle(tmp_path, seed):
    rng = np.random.default_rng(100 + seed)
    contigs = [(f"chr{i + 1}", rand_seq(rng, int(rng.integers(300, 3000)))) for i in range(3)]
    lines = []
    for name, seq in contigs:
        L = len(seq)
        n = int(rng.integers(1, 40))
        pos = sorted(set(int(x) for x in rng.integers(25, L - 40, n)))
        # dense clusters to exercise the range-merging branches
        if len(pos) > 3:
            pos += [pos[0] + 3, pos[1] + 11, pos[2] + 22, pos[2] + 23]
            pos = sorted(set(p for p in pos if p < L - 40))
        for p in pos:
            kind = rng.random()
            if kind < 0.6:
                ref = seq[p]; alts = [rng.choice([c for c in "ACGT" if c != ref])]
            elif kind < 0.8:
                ref = seq[p]; alts = [ref + rand_seq(rng, int(rng.integers(1, 8)))]
            else:
                d = int(rng.integers(1, 9)); ref = seq[p:p + 1 + d]; alts = [seq[p]]
            if rng.random() < 0.15:
                alts.append(seq[p] + rand_seq(rng, 2))
            gts = ["0|1", "1|0", "1|1", "0/1", "1/1", "0|0", "1/0"] + (["1|2", "2|1", "1/2", "0|2"] if len(alts) > 1 else [])
            lines.append(f"{name}\t{p + 1}\t.\t{ref}\t{','.join(alts)}\t.\t.\t.\tGT:GQ\t{rng.choice(gts)}:30\n")
    recs = run(tmp_path, contigs, "".join(lines))
    assert len(recs) > 0
    for rid, seq in recs:
        assert set(seq) <= set("ACGTN") and len(seq) >= 23


def test_usage_and_argument_errors(tmp_path):
    assert subprocess.run([EXE], capture_output=True).returncode == 1
    r = subprocess.run([EXE, "a.vcf", "o.fa", "g.fa", "x", "23", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot cast x into an unsigned" in r.stderr
    r = subprocess.run([EXE, str(tmp_path / "missing.vcf"), str(tmp_path / "o.fa"), str(tmp_path / "g.fa"), "0", "23", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not open VCF file" in r.stdout


def test_segments_feed_the_packer(tmp_path):
    """The FASTA vcf_loader writes is what bidir_index packs (ids without whitespace, 70-column lines)."""
    import varscot_b200 as V
    rng = np.random.default_rng(9)
    seq = rand_seq(rng, 3000)
    body = "".join(f"chr1\t{p + 1}\t.\t{seq[p]}\t{'A' if seq[p] != 'A' else 'C'}\t.\t.\t.\tGT\t0|1\n" for p in range(100, 2900, 97))
    recs = run(tmp_path, [("chr1", seq)], body)
    t = V.PackedText.from_fasta(str(tmp_path / "snp.fa"))
    assert t.names == [r[0] for r in recs] and t.n_bases == sum(len(r[1]) for r in recs)
    assert all("_" not in n.split("_", 1)[0] for n in t.names)


FW = os.path.join(ROOT, "build", "variant_processing_build", "fasta_writer")


def test_fasta_writer_guides_and_flanks(tmp_path):
    """Row f4: BED6 on-targets -> 23-nt guide FASTA + 30-nt flanking FASTA (extract_fasta_ontargets.h:44-53,65-69)."""
    rng = np.random.default_rng(5)
    contigs = [("chr1", rand_seq(rng, 500)), ("chr11", rand_seq(rng, 400).lower())]
    g, bed = str(tmp_path / "g.fa"), str(tmp_path / "t.bed")
    write_genome(g, contigs)
    rows = [("chr1", 100, 123, "EMX1", 7, "+"), ("chr11", 50, 73, "CD151", 7, "-"), ("chr1", 2, 25, "edge", 0, "+"), ("chr1", 480, 503, "end", 0, "-")]
    open(bed, "w").write("# comment\n" + "".join("\t".join(map(str, r)) + "\n" for r in rows))
    o1, o2 = str(tmp_path / "guides.fa"), str(tmp_path / "flank.fa")
    r = subprocess.run([FW, o1, o2, bed, g], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rc = lambda s: "".join({"A": "T", "C": "G", "G": "C", "T": "A"}[c] for c in reversed(s))
    s1, s2 = contigs[0][1], contigs[1][1].upper()
    exp1 = [("EMX1", s1[100:123]), ("CD151", rc(s2[50:73])), ("edge", s1[2:25]), ("end", rc(s1[480:500]))]
    exp2 = [("EMX1", s1[96:126]), ("CD151", rc(s2[47:77])), ("edge", s1[0:28]), ("end", rc(s1[477:500]))]
    for path, exp, flank in ((o1, exp1, False), (o2, exp2, True)):
        got = open(path).read()
        assert got == "".join(f">{n}\n{s}\n" for n, s in exp)
        assert VO.fasta_writer(bed, g, flank) == exp
    assert len(exp1[0][1]) == 23 and len(exp2[0][1]) == 30
    assert subprocess.run([FW, o1], capture_output=True).returncode == 1
    assert os.path.exists(g + ".fai")
