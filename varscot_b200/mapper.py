"""Host-side mirror of the reference's read_mapping interface, on top of the C ABI.

The reference exposes this path only as two executables (VARSCOT_pipeline/read_mapping/bidir_index.cpp,
bidir_mapping.cpp); `bidir_index()` and `bidir_mapping()` below take the same options with the same meaning
and exit codes, and call the very same entry points the executables call.  The lower-level classes
(PackedText, ScanContext) are what bench.py and the parity tests drive.  All compute happens in the CUDA
library; nothing here scores windows.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import GLEN, Hit, Record, ScanStats, VarscotError, check

WORD_DT = np.dtype([("hi", "<u4"), ("lo", "<u4"), ("nm", "<u4"), ("em", "<u4")])
HIT_DT = np.dtype([("pos", "<u4"), ("info", "<u4")])
REC_DT = np.dtype([("guide", "<u4"), ("contig", "<u4"), ("pos", "<u4"), ("flag", "<u2"), ("mm", "u1"), ("pad", "u1")])

MD_SEQAN, MD_SAMTOOLS = 0, 1

_GUIDE_LUT = np.zeros(256, dtype=np.uint8)
for _ch, _v in (("C", 1), ("G", 2), ("T", 3), ("U", 3)):
    _GUIDE_LUT[ord(_ch)] = _v
    _GUIDE_LUT[ord(_ch.lower())] = _v


def guide_codes(guides) -> np.ndarray:
    """23-nt guide strings -> (n, 23) Dna codes; non-ACGT -> A (bidir_mapping.cpp:194,256)."""
    out = np.zeros((len(guides), GLEN), dtype=np.uint8)
    for i, g in enumerate(guides):
        b = g.encode() if isinstance(g, str) else bytes(g)
        if len(b) != GLEN:
            raise ValueError(f"guide {i} is not {GLEN} nt")
        out[i] = _GUIDE_LUT[np.frombuffer(b, dtype=np.uint8)]
    return out


def pam_code(pam) -> int:
    """-P XY -> 4*x+y, or -1 when it can never match (bidir_mapping.cpp:240-247)."""
    if not pam:
        return -1
    lut = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}
    if len(pam) != 2 or pam[0].upper() not in lut or pam[1].upper() not in lut:
        return -1
    return 4 * lut[pam[0].upper()] + lut[pam[1].upper()]


@dataclass
class PackedText:
    """Bit-sliced text: words[(n_words + 1)] of {hi, lo, nm, em}, contig offsets, optional names."""
    words: np.ndarray
    offsets: np.ndarray
    n_bases: int
    names: list | None = None

    @property
    def n_words(self) -> int:
        return (self.n_bases + 31) // 32

    @property
    def n_contigs(self) -> int:
        return len(self.offsets) - 1

    @staticmethod
    def from_ascii(ascii_bytes, offsets, names=None) -> "PackedText":
        L = _lib.lib()
        a = np.frombuffer(ascii_bytes, dtype=np.uint8) if isinstance(ascii_bytes, (bytes, bytearray)) else np.ascontiguousarray(ascii_bytes, dtype=np.uint8)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = int(a.size)
        words = np.zeros((n + 31) // 32 + 1, dtype=WORD_DT)
        check(L.vs_pack_text(a.ctypes.data, n, off.ctypes.data, len(off) - 1, words.ctypes.data))
        return PackedText(words, off, n, names)

    @staticmethod
    def from_fasta(path: str) -> "PackedText":
        """Stream a FASTA through vs_packer_* exactly as bidir_index does."""
        L = _lib.lib()
        p = L.vs_packer_new()
        names = []
        try:
            have = False
            with open(path, "rb") as f:
                for line in f:
                    if line.startswith(b">"):
                        if have:
                            check(L.vs_packer_end_contig(p))
                        names.append(line[1:].rstrip(b"\r\n").decode())
                        have = True
                    elif have:
                        check(L.vs_packer_append(p, line, len(line)))
            if have:
                check(L.vs_packer_end_contig(p))
            n = int(L.vs_packer_num_bases(p))
            nc = int(L.vs_packer_num_contigs(p))
            nw = int(L.vs_packer_num_words(p))
            wp = L.vs_packer_words(p)
            words = np.ctypeslib.as_array(C.cast(wp, C.POINTER(C.c_uint32)), shape=((nw + 1) * 4,)).copy().view(WORD_DT)
            off = np.ctypeslib.as_array(C.cast(L.vs_packer_offsets(p), C.POINTER(C.c_uint64)), shape=(nc + 1,)).copy()
        finally:
            L.vs_packer_free(p)
        return PackedText(words, off, n, names)

    def save(self, prefix: str):
        check(_lib.lib().vs_text_save(prefix.encode(), self.words.ctypes.data, self.n_bases, self.offsets.ctypes.data, self.n_contigs))

    @staticmethod
    def load(prefix: str) -> "PackedText":
        L = _lib.lib()
        wp, op = C.c_void_p(), C.c_void_p()
        nb, nc = C.c_uint64(), C.c_uint32()
        check(L.vs_text_load(prefix.encode(), C.byref(wp), C.byref(nb), C.byref(op), C.byref(nc)))
        try:
            nw = (nb.value + 31) // 32
            words = np.ctypeslib.as_array(C.cast(wp, C.POINTER(C.c_uint32)), shape=((nw + 1) * 4,)).copy().view(WORD_DT)
            off = np.ctypeslib.as_array(C.cast(op, C.POINTER(C.c_uint64)), shape=(nc.value + 1,)).copy()
        finally:
            L.vs_free(wp); L.vs_free(op)
        return PackedText(words, off, int(nb.value))


class ScanContext:
    """One device context (vs_ctx): upload a shard of packed text, scan guides against it."""

    def __init__(self, device: int = 0):
        self._L = _lib.lib()
        self._ctx = C.c_void_p()
        check(self._L.vs_ctx_create(device, C.byref(self._ctx)))
        self.device = device
        self.last_stats = None

    def close(self):
        if self._ctx:
            self._L.vs_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, words: np.ndarray, first_word: int = 0, n_words: int | None = None, pinned_ptr: int | None = None):
        """Upload words[first_word : first_word + n_words] (+ the following halo/pad word)."""
        total = len(words) - 1
        if n_words is None:
            n_words = total - first_word
        if first_word < 0 or first_word + n_words > total:
            raise ValueError("shard outside the packed text")
        ptr = pinned_ptr if pinned_ptr is not None else words.ctypes.data
        check(self._L.vs_text_upload(self._ctx, ptr + first_word * 16, n_words, first_word * 32), self._ctx)

    def scan(self, guides: np.ndarray, k: int, pam=None, cap: int = 1 << 20, out: np.ndarray | None = None):
        """Returns (hits structured array, ScanStats). Hits are unordered."""
        g = np.ascontiguousarray(guides, dtype=np.uint8).reshape(-1, GLEN)
        pc = pam if isinstance(pam, int) else pam_code(pam)
        hits = out if out is not None else np.zeros(cap, dtype=HIT_DT)
        n = C.c_uint64()
        st = ScanStats()
        rc = self._L.vs_scan(self._ctx, g.ctypes.data, g.shape[0], k, pc, hits.ctypes.data, len(hits), C.byref(n), C.byref(st))
        if rc == _lib.VS_ERR_OVERFLOW:
            hits = np.zeros(n.value, dtype=HIT_DT)
            check(self._L.vs_scan_fetch(self._ctx, hits.ctypes.data, len(hits), C.byref(n)), self._ctx)
        else:
            check(rc, self._ctx)
        self.last_stats = st
        return hits[: n.value], st

    def measure_int_peaks(self):
        a, b = C.c_double(), C.c_double()
        check(self._L.vs_measure_int_peaks(self._ctx, C.byref(a), C.byref(b)), self._ctx)
        return a.value, b.value


def device_count() -> int:
    n = _lib.lib().vs_device_count()
    return max(n, 0)


def map_packed(text: PackedText, guides: np.ndarray, k: int, pam=None, devices=None):
    """vs_map_packed: shard over devices, scan, return (hits, stats)."""
    L = _lib.lib()
    g = np.ascontiguousarray(guides, dtype=np.uint8).reshape(-1, GLEN)
    dev = np.asarray(devices if devices else [0], dtype=np.int32)
    hp, n, st = C.c_void_p(), C.c_uint64(), ScanStats()
    check(L.vs_map_packed(text.words.ctypes.data, text.n_bases, g.ctypes.data, g.shape[0], k, pam_code(pam) if not isinstance(pam, int) else pam,
                          dev.ctypes.data, len(dev), C.byref(hp), C.byref(n), C.byref(st)))
    try:
        hits = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint32)), shape=(n.value * 2,)).copy().view(HIT_DT) if n.value else np.zeros(0, HIT_DT)
    finally:
        L.vs_free(hp)
    return hits, st


def shard_bounds(n_words: int, n_shards: int) -> np.ndarray:
    """vs_shard_bounds: word ranges owned by each of n_shards ranks / devices."""
    out = np.zeros(n_shards + 1, dtype=np.uint64)
    check(_lib.lib().vs_shard_bounds(n_words, n_shards, out.ctypes.data))
    return out


def resolve_hits(hits: np.ndarray, offsets: np.ndarray):
    """vs_resolve_hits: reference emission order + flags. Returns (records, key16_collisions)."""
    L = _lib.lib()
    h = np.ascontiguousarray(hits, dtype=HIT_DT)
    off = np.ascontiguousarray(offsets, dtype=np.uint64)
    rec = np.zeros(len(h), dtype=REC_DT)
    coll = C.c_uint64()
    check(L.vs_resolve_hits(h.ctypes.data, len(h), off.ctypes.data, len(off) - 1, rec.ctypes.data, C.byref(coll)))
    return rec, int(coll.value)


def md_string(text: PackedText, gpos: int, guide: np.ndarray, strand: int, md_style: int = MD_SEQAN) -> str:
    buf = C.create_string_buffer(64)
    g = np.ascontiguousarray(guide, dtype=np.uint8)
    check(_lib.lib().vs_md_string(text.words.ctypes.data, gpos, g.ctypes.data, strand, md_style, buf))
    return buf.value.decode()


def format_sam(rec, qname: str, rname: str, guide: np.ndarray, md: str) -> str:
    r = Record(int(rec["guide"]), int(rec["contig"]), int(rec["pos"]), int(rec["flag"]), int(rec["mm"]), 0)
    g = np.ascontiguousarray(guide, dtype=np.uint8)
    buf = C.create_string_buffer(len(qname) + len(rname) + 256)
    n = _lib.lib().vs_format_sam(C.byref(r), qname.encode(), rname.encode(), g.ctypes.data, md.encode(), buf, len(buf))
    if n < 0:
        raise VarscotError(_lib.VS_ERR_ARG, "vs_format_sam failed")
    return buf.raw[:n].decode()


def records_key_set(rec: np.ndarray):
    """Parity key: (guide, strand bit, contig, pos, NM)."""
    return set(zip(rec["guide"].tolist(), ((rec["flag"] & 16) >> 4).tolist(), rec["contig"].tolist(), rec["pos"].tolist(), rec["mm"].tolist()))


def _main(fn, prog: str, args: list) -> int:
    argv = [prog.encode()] + [str(a).encode() for a in args]
    arr = (C.c_char_p * (len(argv) + 1))(*argv, None)
    return int(fn(len(argv), arr))


def bidir_index(genome: str, index: str) -> int:
    """`bidir_index -G genome -I index` (bidir_index.cpp:19-24). Returns the exit code."""
    return _main(_lib.lib().vs_bidir_index_main, "bidir_index", ["-G", genome, "-I", index])


def bidir_mapping(genome: str, index: str, reads: str, mismatches: int, output: str, threads: int = 1, pam: str | None = None,
                  md_style: str | None = None) -> int:
    """`bidir_mapping -G -I -R -M -T -O [-P]` (bidir_mapping.cpp:196-216). Returns the exit code."""
    args = ["-G", genome, "-I", index, "-R", reads, "-M", mismatches, "-T", threads, "-O", output]
    if pam:
        args += ["-P", pam]
    if md_style:
        args += ["--md-style", md_style]
    return _main(_lib.lib().vs_bidir_mapping_main, "bidir_mapping", args)
