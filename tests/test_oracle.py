"""CPU tests of the ORACLE (test infrastructure): its two restatements against each other, against the
committed golden fixtures and against hand-derived known-answer vectors.
The reference holds no golden outputs for this path (parity "unpinned"), so these KATs are derived by hand
from bidir_mapping.cpp's rules, each citing the lines it exercises."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import GLEN, make_case, revcomp_codes

HERE = os.path.dirname(os.path.abspath(__file__))
EMX1 = "GAGTCCGAGCAGAAGAAGAAGGG"
LUT = np.frombuffer(b"ACGT", dtype=np.uint8)


def run(text, guides, k, pam=None, offsets=None, **kw):
    codes = O.text_codes(text.encode() if isinstance(text, str) else text)
    off = np.array([0, len(codes)] if offsets is None else offsets, dtype=np.uint64)
    g = O.guide_codes(guides) if isinstance(guides[0], str) else guides
    return O.map_guides(codes, off, g, k, pam=pam, **kw)


def both(text, guides, k, pam=None, offsets=None):
    a = run(text, guides, k, pam, offsets, mode=O.MODE_LITERAL)
    b = run(text, guides, k, pam, offsets, mode=O.MODE_SCAN)
    assert a.rows() == b.rows()
    return a


@pytest.mark.parametrize("seed", range(12))
def test_literal_equals_scan_random(seed):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 500, int(rng.integers(1, 7))).tolist()
    for k in (0, 1, 2, 3, 4, 5, 6, 7, 8):
        case = make_case(seed * 10 + k, lens, 3, k, pam=[None, "AG", "CT"][seed % 3])
        codes = O.text_codes(case.ascii)
        a = O.map_guides(codes, case.offsets, case.guides, k, pam=case.pam, mode=O.MODE_LITERAL)
        b = O.map_guides(codes, case.offsets, case.guides, k, pam=case.pam, mode=O.MODE_SCAN)
        assert a.rows() == b.rows()


def _spec_rows(ascii_bytes, offsets, guide_strs, k, pam):
    """Third, independent statement of the path, written from the rules R1-R9 of SURVEY.md 8a (not from the C oracle):
    plain Python over strings, one window at a time.  Small inputs only."""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rc = lambda x: "".join(comp[c] for c in reversed(x))
    text = "".join(c if c in "ACGT" else "N" for c in ascii_bytes.decode().upper().replace("U", "T"))        # R6
    fwd = {"GG", "GA"} | ({pam.upper()} if pam else set())                                                    # R2
    rev = {rc(x) for x in fwd}
    rows = []
    for gi, g in enumerate(guide_strs):
        g = "".join(c if c in "ACGT" else "A" for c in g.upper().replace("U", "T"))                           # R5
        for strand, pat in ((0, g), (1, rc(g))):
            recs = []
            for ci in range(len(offsets) - 1):
                c = text[int(offsets[ci]):int(offsets[ci + 1])]
                for p in range(len(c) - 22):                                                                  # R1
                    w = c[p:p + 23]
                    if "N" in w:                                                                              # R3
                        continue
                    if (w[21:23] not in fwd) if strand == 0 else (w[0:2] not in rev):                         # R2
                        continue
                    mm = sum(a != b for a, b in zip(pat, w))
                    if mm > k:                                                                                # R3
                        continue
                    if p + 23 == len(c) and sum(a != b for a, b in zip(pat[11:], w[11:])) > k // 2:           # R4
                        continue
                    md, run = "", 0                                                                           # R9, SeqAn style
                    for a, b in zip(pat, w):
                        if a == b:
                            run += 1
                        else:
                            md += (str(run) if run else "") + b
                            run = 0
                    md += str(run) if run else ""
                    recs.append(((ci & 0xFFFF, p, ci >> 16), ci, p, mm, md))
            recs.sort()                                                                                       # R8: std::map order
            out, best = [], None
            for _, ci, p, mm, md in recs:                                                                     # running best
                if best is None:
                    best = (ci, p, mm, md)
                elif mm >= best[2]:
                    out.append((gi, 16 * strand + 256, ci, p, mm, md))
                else:
                    out.append((gi, 16 * strand + 256, best[0], best[1], best[2], best[3]))
                    best = (ci, p, mm, md)
            if best is not None:
                out.append((gi, 16 * strand, best[0], best[1], best[2], best[3]))
            rows += out
    return rows


@pytest.mark.parametrize("seed", range(8))
def test_python_spec_restatement_agrees_with_the_c_oracle(seed):
    rng = np.random.default_rng(500 + seed)
    lens = [int(x) for x in rng.choice([0, 5, 22, 23, 24, 45, 46, 300, 900], int(rng.integers(2, 7)))] + [700]
    k = int(rng.integers(0, 9))
    pam = [None, "AG", "TT", "cc"][seed % 4]
    case = make_case(900 + seed, lens, 3, k, pam=pam, n_frac=0.01, iupac_frac=0.005)
    exp = _spec_rows(case.ascii, case.offsets, case.guide_strs, k, pam)
    for mode in (O.MODE_SCAN, O.MODE_LITERAL):
        got = O.map_guides(O.text_codes(case.ascii), case.offsets, case.guides, k, pam=pam, mode=mode).rows()
        assert got == exp, (seed, mode)
    assert len(exp) > 0


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "case_*.json"))))
def test_golden_fixtures(path):
    d = json.load(open(path))
    r = run(d["ascii"], d["guides"], d["k"], d["pam"], d["offsets"])
    assert [list(x) for x in r.rows()] == d["rows"]
    assert len(d["rows"]) > 0


def test_kat_forward_perfect_hit_sam_line():
    # appendix C: forward perfect hit at 0-based pos 1000 -> POS 1001, FLAG 0, NM 0, MD 23 (bidir_mapping.cpp:88-123)
    rng = np.random.default_rng(0)
    text = bytearray(LUT[rng.integers(0, 4, 3000)].tobytes())
    # avoid chance PAM-valid near matches: irrelevant at k=0 except exact duplicates
    text[1000:1023] = EMX1.encode()
    r = both(bytes(text), [EMX1], 0)
    assert r.rows() == [(0, 0, 0, 1000, 0, "23")]
    import ctypes as C
    buf = C.create_string_buffer(512)
    rec = O._Rec(0, 0, 1000, 0, 0, 0, b"23")
    O.lib().vo_format_sam.restype = C.c_int
    n = O.lib().vo_format_sam(C.byref(rec), b"EMX1", b"chr2", O.guide_codes([EMX1]).ctypes.data, buf, 512)
    assert buf.raw[:n].decode() == "EMX1\t0\tchr2\t1001\t255\t23M\t*\t0\t0\tGAGTCCGAGCAGAAGAAGAAGGG\tIIIIIIIIIIIIIIIIIIIIIII\tNM:i:0\tMD:Z:23\n"


def test_kat_reverse_pass_md_and_flag():
    # appendix C: window CCCTTATTCTTCTGCTCGGACTC (genome A where revcomp(EMX1) has C at offset 5): FLAG 16, NM 1, MD 5A17
    w = "CCCTTATTCTTCTGCTCGGACTC"
    text = "T" * 40 + w + "T" * 40
    r = both(text, [EMX1], 1)
    assert r.rows() == [(0, 16, 0, 40, 1, "5A17")]
    assert run(text, [EMX1], 1, md_style=O.MD_SAMTOOLS).md == ["5A17"]


def test_kat_genome_pam_ga_counts_one_mismatch():
    # genome PAM GA under a guide ending GG: one mismatch at offset 22 -> MD 22A (SeqAn) / 22A0 (samtools); :70-86
    w = EMX1[:22] + "A"
    text = "T" * 30 + w + "T" * 30
    assert both(text, [EMX1], 1).rows() == [(0, 0, 0, 30, 1, "22A")]
    assert run(text, [EMX1], 1, md_style=O.MD_SAMTOOLS).md == ["22A0"]
    assert len(both(text, [EMX1], 0)) == 0
    # adjacent mismatches: no 0 between them in SeqAn style
    w2 = "CT" + EMX1[2:]
    t2 = "A" * 30 + w2 + "A" * 30
    assert both(t2, [EMX1], 2).md == ["CT21"]
    assert run(t2, [EMX1], 2, md_style=O.MD_SAMTOOLS).md == ["0C0T21"]


def test_kat_pam_is_checked_on_the_genome_not_the_guide():
    # guide with a non-GG tail still needs GG/GA on the genome (R2); its own tail then counts as mismatches (R3)
    g = EMX1[:20] + "TTT"
    text = "A" * 30 + EMX1[:20] + "TGG" + "A" * 30
    assert both(text, [g], 2).rows() == [(0, 0, 0, 30, 2, "21GG")]
    assert len(both(text, [g], 1)) == 0
    text2 = "A" * 30 + g + "A" * 30          # exact copy of the guide, but genome PAM is TT
    assert len(both(text2, [g], 8)) == 0


def test_kat_extra_pam_and_its_reverse_complement():
    site = EMX1[:21] + "AG"
    text = "T" * 30 + site + "T" * 30
    assert len(both(text, [EMX1], 4)) == 0
    assert both(text, [EMX1], 4, pam="AG").rows() == [(0, 0, 0, 30, 1, "21A1")]
    rc = "".join("ACGT"[3 - "ACGT".index(c)] for c in reversed(site))       # starts with CT = revcomp(AG)
    text = "A" * 30 + rc + "A" * 30
    assert len(both(text, [EMX1], 4)) == 0
    assert [x[:5] for x in both(text, [EMX1], 4, pam="AG").rows()] == [(0, 16, 0, 30, 1)]


def test_kat_r4_last_window_needs_second_half_seed():
    # appendix C R4: contig of length 45, k=4 -> K=2.  Last window (p=22) with H=3: reported iff H2 <= 2 (:48-53)
    g = O.guide_codes([EMX1])[0]
    rng = np.random.default_rng(4)

    def contig(mm_pos, at_end):
        w = g.copy()
        for p in mm_pos:
            w[p] = (w[p] + 1) % 4
        pad = rng.integers(0, 4, 22).astype(np.uint8)
        pad[:] = 3                                         # T filler: no PAM-valid chance hits
        c = np.concatenate([pad, w]) if at_end else np.concatenate([w, pad, [3]])
        return LUT[c].tobytes()

    # H2 = 3 > K : dropped only when the window is the last one of the contig
    h2_3 = [12, 13, 14]
    assert len(both(contig(h2_3, True), [EMX1], 4)) == 0
    assert [x[3:5] for x in both(contig(h2_3, False), [EMX1], 4).rows()] == [(0, 3)]
    # H1 = 3, H2 = 0 : kept in both places
    h1_3 = [1, 2, 3]
    assert [x[3:5] for x in both(contig(h1_3, True), [EMX1], 4).rows()] == [(22, 3)]
    assert [x[3:5] for x in both(contig(h1_3, False), [EMX1], 4).rows()] == [(0, 3)]
    # L == 23: the single window is a last window; L < 23: no window
    assert [x[3:5] for x in both(LUT[g].tobytes(), [EMX1], 4).rows()] == [(0, 0)]
    assert len(both(LUT[g[:22]].tobytes(), [EMX1], 4)) == 0
    # k -> K (bidir_mapping.cpp:129-146): with H2 = 2 the last window needs K >= 2, i.e. k >= 4
    h2_2 = [12, 13]
    for k, expect in ((2, 0), (3, 0), (4, 1), (5, 1), (8, 1)):
        assert len(both(contig(h2_2, True), [EMX1], k)) == expect


def test_kat_n_in_window_rejects_and_alphabets():
    # R3/R6: any N (or IUPAC, '*') in the window rejects it; lowercase is folded; R5: guide N -> A
    for i in range(GLEN):
        w = list(EMX1); w[i] = "N"
        assert len(both("T" * 30 + "".join(w) + "T" * 30, [EMX1], 8)) == 0
    assert len(both("T" * 30 + EMX1[:5] + "R" + EMX1[6:] + "T" * 30, [EMX1], 8)) == 0
    assert both("t" * 30 + EMX1.lower() + "t" * 30, [EMX1], 0).rows() == [(0, 0, 0, 30, 0, "23")]
    gN = "N" + EMX1[1:]                                 # becomes A...: one mismatch against the G in the genome
    assert both("T" * 30 + EMX1 + "T" * 30, [gN], 1).rows() == [(0, 0, 0, 30, 1, "G22")]
    assert O.guide_codes(["acgtnACGTNxyzuU" + "A" * 8]).tolist()[0][:15] == [0, 1, 2, 3, 0, 0, 1, 2, 3, 0, 0, 0, 0, 3, 3]


def test_kat_r8_running_best_order_and_flags():
    # appendix C R8: key-ordered records with mm = [3,4,2,5,1] are emitted r1 r0 r3 r2 r4(primary) (:164-187)
    g = O.guide_codes([EMX1])[0]

    def site(mm):
        w = g.copy()
        for p in range(mm):
            w[p] = (w[p] + 1) % 4
        return LUT[w].tobytes()

    def text_for(mms):
        return b"".join(b"T" * 10 + site(m) for m in mms) + b"T" * 10

    r = both(text_for([3, 4, 2, 5, 1]), [EMX1], 5)
    pos = [10 + 33 * i for i in range(5)]
    assert [(x[3], x[1], x[4]) for x in r.rows()] == [(pos[1], 256, 4), (pos[0], 256, 3), (pos[3], 256, 5), (pos[2], 256, 2), (pos[4], 0, 1)]
    r = both(text_for([2, 2, 2]), [EMX1], 5)
    assert [(x[3], x[1]) for x in r.rows()] == [(43, 256), (76, 256), (10, 0)]


def test_both_strands_same_position_palindromic_pam():
    # a window starting CC and ending GG can hit on both passes; forward records precede reverse ones (:285-295)
    g = "CCAAACCCGGGTTTAAACCCTGG"
    rc = "".join("ACGT"[3 - "ACGT".index(c)] for c in reversed(g))
    text = "T" * 30 + g + "T" * 10 + rc + "T" * 30
    r = both(text, [g], 0)
    assert [(x[1], x[3]) for x in r.rows()] == [(0, 30), (16, 63)]


def test_key16_collisions_are_counted_not_hidden():
    # R7: contigs c and c + 65536 share the uint16 key; REF16 keeps one and counts, WIDE keeps both adjacent
    nct = 65536 + 2
    codes = np.full((nct, 45), 3, dtype=np.uint8)
    g = O.guide_codes([EMX1])[0]
    codes[1, 5:28] = g
    codes[65537, 5:28] = g
    off = (np.arange(nct + 1) * 45).astype(np.uint64)
    wide = O.map_guides(codes.reshape(-1), off, g.reshape(1, -1), 0, key_mode=O.KEY_WIDE)
    assert [(x[2], x[3]) for x in wide.rows()] == [(65537, 5), (1, 5)] or [(x[2], x[3]) for x in wide.rows()] == [(1, 5), (65537, 5)]
    assert sorted(x[1] for x in wide.rows()) == [0, 256]
    ref = O.map_guides(codes.reshape(-1), off, g.reshape(1, -1), 0, key_mode=O.KEY_REF16)
    assert len(ref) == 1 and ref.key16_collisions == 1


def test_scan_count_matches_records():
    case = make_case(5, [30000, 45, 45, 100], 4, 6)
    codes = O.text_codes(case.ascii)
    n = O.scan_count(codes, case.offsets, case.guides, 6)
    assert n == len(O.map_guides(codes, case.offsets, case.guides, 6))


def test_argument_errors():
    with pytest.raises(RuntimeError):
        run("ACGT" * 20, [EMX1], 9)
