"""Seeded synthetic workloads of BASELINE.json / SURVEY.md section 8d as BIT PLANES — numpy only, no product import.

This file is loaded two ways: as `varscot_b200.synth_core` by varscot_b200/synth.py (which wraps the planes into
PackedText objects), and BY PATH (importlib) by `bench.py --impl reference`, so that the reference arm generates the very
same text without loading the CUDA library.  A text is a `Planes` record: hi / lo base planes, the N plane `nm`, the
contig-end plane `em` (n_words + 1 words each), contig offsets and the number of bases.

The variant segments follow the shapes vcf_loader emits: an isolated SNV at p gives the 45-mer
[p-22, p+23) once as a pure-REF copy and once with the ALT base at offset 22; an insertion of d bases gives REF 45 /
ALT 45+d; a deletion gives REF 45+d / ALT 45; hom-alt records emit only the ALT contig.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Planes:
    hi: np.ndarray
    lo: np.ndarray
    nm: np.ndarray
    em: np.ndarray
    offsets: np.ndarray
    n_bases: int
    names: list | None = None

    @property
    def n_words(self) -> int:
        return (self.n_bases + 31) // 32


# hg38 chromosome lengths in Mbp (1..22, X, Y): the "human-like length ladder" of config 2
HUMAN_LADDER = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09, 133.28,
                114.36, 107.04, 101.99, 90.34, 83.26, 80.37, 58.62, 64.44, 46.71, 50.82, 156.04, 57.23]


def _set_bit_range(arr: np.ndarray, start: int, end: int):
    """Set bits [start, end) of a little-endian uint32 bit array."""
    if end <= start:
        return
    w0, w1 = start >> 5, (end - 1) >> 5
    m0 = np.uint32((0xFFFFFFFF << (start & 31)) & 0xFFFFFFFF)
    m1 = np.uint32(0xFFFFFFFF >> (31 - ((end - 1) & 31)))
    if w0 == w1:
        arr[w0] |= m0 & m1
    else:
        arr[w0] |= m0
        arr[w0 + 1:w1] = 0xFFFFFFFF
        arr[w1] |= m1


def _set_bits(arr: np.ndarray, idx: np.ndarray):
    np.bitwise_or.at(arr, (idx >> 5).astype(np.int64), (np.uint32(1) << (idx & 31).astype(np.uint32)))


def contig_ladder(total_bases: int, n_contigs: int = 24) -> np.ndarray:
    frac = np.array((HUMAN_LADDER * ((n_contigs + 23) // 24))[:n_contigs], dtype=np.float64)
    lens = np.floor(frac / frac.sum() * total_bases).astype(np.int64)
    lens[0] += total_bases - lens.sum()
    return lens


def synth_genome(seed: int, total_bases: int, n_contigs: int = 24, n_frac: float = 0.05) -> Planes:
    """Uniform-random ACGT contigs on a human-like length ladder; per contig N runs at both ends (10 kb) and
    one central run so that about n_frac of all bases are N.  total_bases is rounded down to a multiple of 32."""
    total_bases = int(total_bases) // 32 * 32
    rng = np.random.default_rng(seed)
    lens = contig_ladder(total_bases, n_contigs)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    nw = total_bases // 32
    raw = rng.integers(0, 1 << 32, size=(2, nw + 1), dtype=np.uint64).astype(np.uint32)
    nm = np.zeros(nw + 1, dtype=np.uint32)
    em = np.zeros(nw + 1, dtype=np.uint32)
    for c in range(n_contigs):
        s, e = int(off[c]), int(off[c + 1])
        L = e - s
        if L <= 0:
            continue
        tel = min(10_000, L // 20)
        _set_bit_range(nm, s, s + tel)
        _set_bit_range(nm, e - tel, e)
        cen = int(max(0.0, n_frac * L - 2 * tel))
        if cen > 0:
            m = s + L // 3
            _set_bit_range(nm, m, m + cen)
        em[(e - 1) >> 5] |= np.uint32(1) << np.uint32((e - 1) & 31)
    nm[nw] = 0xFFFFFFFF
    names = [f"chr{c + 1}" for c in range(n_contigs)]
    hi, lo = raw[0], raw[1]
    hi[:nw] &= ~nm[:nw]; lo[:nw] &= ~nm[:nw]          # N and padding read as 00 in the base planes
    hi[nw] = 0; lo[nw] = 0
    return Planes(hi, lo, nm, em, off, total_bases, names)


def _gather(plane: np.ndarray, starts: np.ndarray, length: int) -> np.ndarray:
    """bits [start, start+length) (length <= 57) of a uint32 bit plane, as uint64."""
    w = (starts >> 5).astype(np.int64)
    s = (starts & 31).astype(np.uint64)
    a = plane[w].astype(np.uint64) | (plane[w + 1].astype(np.uint64) << np.uint64(32))
    b = plane[w + 2].astype(np.uint64)
    lo = a >> s
    hi = np.where(s > 0, b << ((np.uint64(64) - s) & np.uint64(63)), np.uint64(0))
    return (lo | hi) & np.uint64((1 << length) - 1)


def synth_variant_segments(genome: Planes, seed: int, n_variants: int, snv_frac: float = 0.90, ins_frac: float = 0.05,
                           hom_frac: float = 0.20, max_indel: int = 10) -> Planes:
    """The "SNP genome": per variant the REF/ALT haplotype windows cut from the packed genome."""
    rng = np.random.default_rng(seed)
    nwg = genome.n_words
    # planes padded by two words so that _gather may read w+2
    hi = np.concatenate([genome.hi[:nwg + 1], np.zeros(2, np.uint32)])
    lo = np.concatenate([genome.lo[:nwg + 1], np.zeros(2, np.uint32)])
    nm = np.concatenate([genome.nm[:nwg + 1], np.full(2, 0xFFFFFFFF, np.uint32)])
    em = np.concatenate([genome.em[:nwg + 1], np.zeros(2, np.uint32)])
    span = 45 + max_indel
    # candidate positions: window [p-22, p-22+span) must be N-free and inside one contig
    n_try = int(n_variants * 1.25) + 1000
    p = np.sort(rng.integers(22, genome.n_bases - span, n_try).astype(np.int64))
    ok = (_gather(nm, p - 22, span) == 0) & (_gather(em, p - 22, span - 1) == 0)
    p = np.unique(p[ok])
    if len(p) > n_variants:
        p = np.sort(rng.choice(p, n_variants, replace=False))
    n = len(p)
    kind = rng.random(n)
    is_snv = kind < snv_frac
    is_ins = (~is_snv) & (kind < snv_frac + ins_frac)
    is_del = ~(is_snv | is_ins)
    d = np.where(is_snv, 0, rng.integers(1, max_indel + 1, n)).astype(np.uint64)
    hom = rng.random(n) < hom_frac
    wh = _gather(hi, p - 22, span)
    wl = _gather(lo, p - 22, span)
    one = np.uint64(1)
    m23, m22 = np.uint64((1 << 23) - 1), np.uint64((1 << 22) - 1)
    # REF haplotype: 45 bases, or 45 + d for deletions
    ref_len = np.where(is_del, 45 + d, 45).astype(np.int64)
    ref_mask = (one << ref_len.astype(np.uint64)) - one
    ref_h, ref_l = wh & ref_mask, wl & ref_mask
    # ALT haplotype
    alt_len = np.where(is_ins, 45 + d, 45).astype(np.int64)
    # SNV: change the base at offset 22 to a different one
    delta = rng.integers(1, 4, n).astype(np.uint64)
    code = (((wh >> np.uint64(22)) & one) << one) | ((wl >> np.uint64(22)) & one)
    ncode = (code + delta) & np.uint64(3)
    bit22 = one << np.uint64(22)
    snv_h = (wh & ~bit22 | ((ncode >> one) & one) << np.uint64(22)) & np.uint64((1 << 45) - 1)
    snv_l = (wl & ~bit22 | (ncode & one) << np.uint64(22)) & np.uint64((1 << 45) - 1)
    # insertion: left 23 bases, d random bases, right 22 bases
    insb = rng.integers(0, 1 << 20, size=(2, n), dtype=np.uint64) & ((one << d) - one)
    right_h, right_l = (wh >> np.uint64(23)) & m22, (wl >> np.uint64(23)) & m22
    ins_h = (wh & m23) | (insb[0] << np.uint64(23)) | (right_h << (np.uint64(23) + d))
    ins_l = (wl & m23) | (insb[1] << np.uint64(23)) | (right_l << (np.uint64(23) + d))
    # deletion: left 23 bases, skip d, right 22 bases
    del_h = (wh & m23) | (((wh >> (np.uint64(23) + d)) & m22) << np.uint64(23))
    del_l = (wl & m23) | (((wl >> (np.uint64(23) + d)) & m22) << np.uint64(23))
    alt_h = np.where(is_snv, snv_h, np.where(is_ins, ins_h, del_h))
    alt_l = np.where(is_snv, snv_l, np.where(is_ins, ins_l, del_l))
    # emission order: per variant REF (unless hom-alt) then ALT
    emit_ref = ~hom
    cnt = emit_ref.astype(np.int64) + 1
    total = int(cnt.sum())
    first = np.cumsum(cnt) - cnt
    seg_h = np.zeros(total, np.uint64); seg_l = np.zeros(total, np.uint64); seg_len = np.zeros(total, np.int64)
    ri = first[emit_ref]
    seg_h[ri], seg_l[ri], seg_len[ri] = ref_h[emit_ref], ref_l[emit_ref], ref_len[emit_ref]
    ai = first + emit_ref
    seg_h[ai], seg_l[ai], seg_len[ai] = alt_h, alt_l, alt_len
    off = np.concatenate([[0], np.cumsum(seg_len)]).astype(np.uint64)
    n_bases = int(off[-1])
    nw = (n_bases + 31) // 32
    cols = np.arange(span, dtype=np.uint64)
    keep = cols[None, :] < seg_len[:, None].astype(np.uint64)
    planes = {}
    for name, seg in (("hi", seg_h), ("lo", seg_l)):
        bits = ((seg[:, None] >> cols[None, :]) & one).astype(np.uint8)[keep]
        packed = np.packbits(bits, bitorder="little")
        buf = np.zeros((nw + 1) * 4, np.uint8)
        buf[:len(packed)] = packed
        planes[name] = buf.view("<u4")
    emw = np.zeros(nw + 1, np.uint32)
    _set_bits(emw, off[1:].astype(np.int64) - 1)
    nmw = np.zeros(nw + 1, np.uint32)
    if n_bases & 31:
        nmw[nw - 1] = np.uint32((0xFFFFFFFF << (n_bases & 31)) & 0xFFFFFFFF)
    nmw[nw] = 0xFFFFFFFF
    return Planes(planes["hi"], planes["lo"], nmw, emw, off, n_bases, None)


def concat_texts(a: Planes, b: Planes) -> Planes:
    """Concatenate two packed texts; a.n_bases must be a multiple of 32 (synth_genome guarantees it)."""
    if a.n_bases % 32:
        raise ValueError("first text must end on a word boundary")
    na = a.n_words
    hi = np.concatenate([a.hi[:na], b.hi])
    lo = np.concatenate([a.lo[:na], b.lo])
    nm = np.concatenate([a.nm[:na], b.nm])
    em = np.concatenate([a.em[:na], b.em])
    off = np.concatenate([a.offsets, b.offsets[1:] + np.uint64(a.n_bases)]).astype(np.uint64)
    return Planes(hi, lo, nm, em, off, a.n_bases + b.n_bases, None)


def unpack_codes(text, start: int = 0, n: int | None = None) -> np.ndarray:
    """Dna5 codes (0..3, 4 = N) of bases [start, start+n): what the CPU oracle consumes. start % 32 == 0."""
    if n is None:
        n = text.n_bases - start
    if start % 32:
        raise ValueError("start must be word aligned")
    w0, w1 = start // 32, (start + n + 31) // 32

    def bits(x):
        return np.unpackbits(np.ascontiguousarray(x[w0:w1]).view(np.uint8), bitorder="little")[:n]

    codes = (bits(text.hi) << 1) | bits(text.lo)
    codes[bits(text.nm) == 1] = 4
    return codes


def slice_offsets(text, start: int, n: int) -> np.ndarray:
    """Contig offsets of the sub-text [start, start+n) (contigs cut at the slice borders)."""
    off = text.offsets.astype(np.int64)
    inner = off[(off > start) & (off < start + n)] - start
    return np.concatenate([[0], inner, [n]]).astype(np.uint64)


def synth_guides(seed: int, n: int) -> np.ndarray:
    """20 random nt + random N-position base + GG."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, (n, 23)).astype(np.uint8)
    g[:, 21:] = 2
    return g


def build_workload(genome_bases: int, n_variants: int, scale: float = 1.0, seed_genome: int = 11, seed_variants: int = 12) -> Planes:
    """Text of a BASELINE config: genome (24 contigs, 5 % N) + the variant segments cut from it."""
    g = synth_genome(seed_genome, int(genome_bases * scale), 24, 0.05)
    nv = int(n_variants * scale)
    if nv > 0:
        return concat_texts(g, synth_variant_segments(g, seed_variants, nv))
    return g
