"""ctypes binding of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  See oracle/vo_core.h for what the oracle restates
(VARSCOT_pipeline/read_mapping/bidir_mapping.cpp) and why parity is "unpinned".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libvo_oracle.so")

MODE_LITERAL, MODE_SCAN = 0, 1
KEY_REF16, KEY_WIDE = 0, 1
MD_SEQAN, MD_SAMTOOLS = 0, 1
GLEN = 23


class _Rec(C.Structure):
    _fields_ = [("guide", C.c_uint32), ("contig", C.c_uint32), ("pos", C.c_uint32),
                ("flag", C.c_uint16), ("mm", C.c_uint8), ("pad", C.c_uint8),
                ("md", C.c_char * 64)]


class _Res(C.Structure):
    _fields_ = [("rec", C.POINTER(_Rec)), ("n", C.c_uint64), ("cap", C.c_uint64),
                ("key16_collisions", C.c_uint64)]


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile."""
    if force or not os.path.exists(_LIB) or \
            os.path.getmtime(_LIB) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("vo_core.c", "vo_core.h")):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.vo_map.restype = C.c_int
        L.vo_map.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                             C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_Res)]
        L.vo_result_free.argtypes = [C.POINTER(_Res)]
        L.vo_scan_count.restype = C.c_uint64
        L.vo_scan_count.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                    C.c_int, C.c_int, C.c_int]
        L.vo_md_string.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char_p]
        L.vo_num_procs.restype = C.c_int
        _lib = L
    return _lib


_TEXT_LUT = np.full(256, 4, dtype=np.uint8)
for _ch, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3)):
    _TEXT_LUT[ord(_ch)] = _v
    _TEXT_LUT[ord(_ch.lower())] = _v
_GUIDE_LUT = _TEXT_LUT.copy()
_GUIDE_LUT[_GUIDE_LUT > 3] = 0


def text_codes(ascii_bytes) -> np.ndarray:
    """R6: A,C,G,T(U) case-insensitive -> 0..3, everything else -> 4 (N)."""
    a = np.frombuffer(ascii_bytes, dtype=np.uint8) if isinstance(ascii_bytes, (bytes, bytearray)) else np.asarray(ascii_bytes, dtype=np.uint8)
    return _TEXT_LUT[a]


def guide_codes(guides) -> np.ndarray:
    """R5: list of 23-char strings -> (n, 23) uint8 codes, non-ACGT -> A."""
    out = np.zeros((len(guides), GLEN), dtype=np.uint8)
    for i, g in enumerate(guides):
        b = g.encode() if isinstance(g, str) else g
        if len(b) != GLEN:
            raise ValueError(f"guide {i} is not {GLEN} nt")
        out[i] = _GUIDE_LUT[np.frombuffer(b, dtype=np.uint8)]
    return out


def pam_code(pam: str | None) -> int:
    if not pam:
        return -1
    x, y = int(_TEXT_LUT[ord(pam[0])]), int(_TEXT_LUT[ord(pam[1])])
    return -1 if x > 3 or y > 3 else 4 * x + y


@dataclass
class Records:
    guide: np.ndarray
    contig: np.ndarray
    pos: np.ndarray
    flag: np.ndarray
    mm: np.ndarray
    md: list
    key16_collisions: int

    def __len__(self):
        return len(self.guide)

    def key_set(self):
        """Parity key of SURVEY.md section 8a: (guide, strand bit, contig, pos, NM)."""
        return set(zip(self.guide.tolist(), ((self.flag & 16) >> 4).tolist(), self.contig.tolist(),
                       self.pos.tolist(), self.mm.tolist()))

    def rows(self):
        return list(zip(self.guide.tolist(), self.flag.tolist(), self.contig.tolist(), self.pos.tolist(),
                        self.mm.tolist(), self.md))


def map_guides(codes: np.ndarray, offsets: np.ndarray, guides: np.ndarray, k: int, pam: str | None = None,
               mode: int = MODE_SCAN, key_mode: int = KEY_WIDE, md_style: int = MD_SEQAN, threads: int = 0) -> Records:
    """Run the oracle. codes: Dna5 codes (uint8), offsets: n_contigs+1 uint64, guides: (n,23) uint8 codes."""
    L = lib()
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    guides = np.ascontiguousarray(guides, dtype=np.uint8).reshape(-1, GLEN)
    if threads <= 0:
        threads = L.vo_num_procs()
    res = _Res()
    rc = L.vo_map(codes.ctypes.data, offsets.ctypes.data, len(offsets) - 1, guides.ctypes.data, guides.shape[0],
                  k, pam_code(pam), mode, key_mode, md_style, threads, C.byref(res))
    if rc:
        raise RuntimeError(f"vo_map failed with {rc}")
    n = res.n
    arr = np.ctypeslib.as_array(C.cast(res.rec, C.POINTER(C.c_uint8)), shape=(n * C.sizeof(_Rec),)) if n else np.zeros(0, np.uint8)
    dt = np.dtype([("guide", "<u4"), ("contig", "<u4"), ("pos", "<u4"), ("flag", "<u2"), ("mm", "u1"), ("pad", "u1"),
                   ("md", "S64")])
    rec = arr.view(dt).copy() if n else np.zeros(0, dt)
    out = Records(rec["guide"].copy(), rec["contig"].copy(), rec["pos"].copy(), rec["flag"].copy(), rec["mm"].copy(),
                  [m.decode() for m in rec["md"]], int(res.key16_collisions))
    L.vo_result_free(C.byref(res))
    return out


def scan_count(codes, offsets, guides, k, pam=None, threads=0) -> int:
    L = lib()
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    guides = np.ascontiguousarray(guides, dtype=np.uint8).reshape(-1, GLEN)
    if threads <= 0:
        threads = L.vo_num_procs()
    return int(L.vo_scan_count(codes.ctypes.data, offsets.ctypes.data, len(offsets) - 1, guides.ctypes.data,
                               guides.shape[0], k, pam_code(pam), threads))


def md_string(window_codes, pattern_codes, md_style=MD_SEQAN) -> str:
    L = lib()
    w = np.ascontiguousarray(window_codes, dtype=np.uint8)
    p = np.ascontiguousarray(pattern_codes, dtype=np.uint8)
    buf = C.create_string_buffer(64)
    L.vo_md_string(w.ctypes.data, p.ctypes.data, md_style, buf)
    return buf.value.decode()


def num_procs() -> int:
    return int(lib().vo_num_procs())
