"""Seeded synthetic cases shared by the tests, smoke() and the golden-fixture generator."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

GLEN = 23
_L = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class Case:
    ascii: bytes            # all contigs concatenated (may contain N, lowercase, IUPAC)
    offsets: np.ndarray     # n_contigs + 1, uint64
    guides: np.ndarray      # (n, 23) Dna codes
    guide_strs: list
    k: int
    pam: str | None
    names: list


def revcomp_codes(g):
    return (3 - g[::-1]).astype(np.uint8)


def make_case(seed, contig_lens, n_guides, k, pam=None, plant=True, n_frac=0.002, lower_frac=0.01, iupac_frac=0.0005,
              guide_pam="GG") -> Case:
    rng = np.random.default_rng(seed)
    lens = np.asarray(contig_lens, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    total = int(lens.sum())
    codes = rng.integers(0, 4, total).astype(np.uint8)
    guides = rng.integers(0, 4, (n_guides, GLEN)).astype(np.uint8)
    if guide_pam:
        guides[:, 21] = "ACGT".index(guide_pam[0]); guides[:, 22] = "ACGT".index(guide_pam[1])
    if plant:
        # near-matches with 0..k+1 mismatches on both strands, incl. contig starts / ends (R4) and PAM slots
        for g in range(n_guides):
            for c in range(len(lens)):
                L = int(lens[c])
                if L < GLEN:
                    continue
                spots = {0, L - GLEN}
                for _ in range(3):
                    spots.add(int(rng.integers(0, L - GLEN + 1)))
                for p in spots:
                    if rng.random() < 0.35:
                        continue
                    w = guides[g].copy()
                    nm = int(rng.integers(0, k + 2))
                    if nm:
                        idx = rng.choice(GLEN, size=min(nm, GLEN), replace=False)
                        w[idx] = (w[idx] + rng.integers(1, 4, len(idx))) % 4
                    if rng.random() < 0.3:      # force a legal alternative PAM on the site
                        w[21] = 2; w[22] = int(rng.choice([0, 2]))
                    if rng.random() < 0.5:
                        w = revcomp_codes(w)
                    a = int(off[c]) + p
                    codes[a:a + GLEN] = w
    asc = _L[codes].copy()
    n_n = int(total * n_frac)
    if n_n:
        asc[rng.integers(0, total, n_n)] = ord("N")
        # one run of N
        if total > 2000:
            s = int(rng.integers(0, total - 600)); asc[s:s + 500] = ord("N")
    n_i = int(total * iupac_frac)
    if n_i:
        asc[rng.integers(0, total, n_i)] = rng.choice(np.frombuffer(b"RYKMSWBDHV*-", dtype=np.uint8), n_i)
    n_l = int(total * lower_frac)
    if n_l:
        idx = rng.integers(0, total, n_l)
        asc[idx] = np.frombuffer(bytes(asc[idx]).lower(), dtype=np.uint8)
    gs = ["".join("ACGT"[int(b)] for b in g) for g in guides]
    names = [f"ctg{i}" for i in range(len(lens))]
    return Case(bytes(asc), off, guides, gs, k, pam, names)


def make_repeat_case(seed, n_bases, n_guides, k, pam=None, unit_len=37, mut_every=40) -> Case:
    """A LOW-COMPLEXITY text: tandem repeats of one unit that carries a forward PAM (GG at offset 21, plus `pam` at another phase) and
    a reverse one (CC), lightly mutated.  Thousands of candidates share their PAM + neighbouring bases — one bucket of the bucketed
    index spans several batches, which uniform random text never produces below ~10^7 bases — and the guides (windows of the unit with
    0..3 substitutions, both strands) hit at every repeat: a dense-hit case as well."""
    rng = np.random.default_rng(seed)
    unit = rng.integers(0, 4, unit_len).astype(np.uint8)
    unit[21] = 2; unit[22] = 2
    unit[30] = 1; unit[31] = 1
    if pam:
        unit[26] = "ACGT".index(pam[0]); unit[27] = "ACGT".index(pam[1])      # the window starting at offset 5 ends on the extra PAM
    codes = np.tile(unit, (n_bases + unit_len - 1) // unit_len)[:n_bases].copy()
    mut = rng.integers(0, n_bases, max(1, n_bases // mut_every))
    codes[mut] = rng.integers(0, 4, len(mut))
    cut = sorted({0, n_bases, n_bases // 3, min(n_bases, n_bases // 3 + 45)})
    off = np.array(cut, dtype=np.uint64)
    two = np.tile(unit, 3)
    guides = np.zeros((n_guides, GLEN), dtype=np.uint8)
    for g in range(n_guides):
        start = [0, 5, 30 - 0, 9][g % 4] if pam else [0, 30, 0, 30][g % 4]
        w = two[start:start + GLEN].copy()
        if start == 30:
            w = revcomp_codes(w)                              # the reverse-strand site, written as the guide that finds it
        idx = rng.choice(GLEN, size=int(rng.integers(0, 4)), replace=False)
        w[idx] = (w[idx] + rng.integers(1, 4, len(idx))) % 4
        guides[g] = w
    asc = _L[codes]
    gs = ["".join("ACGT"[int(b)] for b in g) for g in guides]
    return Case(bytes(asc), off, guides, gs, k, pam, [f"ctg{i}" for i in range(len(off) - 1)])


def write_fasta(path, names, ascii_bytes, offsets, width=70):
    with open(path, "wb") as f:
        for i, nm in enumerate(names):
            f.write(b">" + nm.encode() + b"\n")
            s = ascii_bytes[int(offsets[i]):int(offsets[i + 1])]
            for j in range(0, len(s), width):
                f.write(s[j:j + width] + b"\n")


def write_guides(path, ids, guide_strs):
    with open(path, "w") as f:
        for i, g in zip(ids, guide_strs):
            f.write(f">{i}\n{g}\n")
