"""The device code of the scan, run on the HOST (tests/cpu_scan_emulator.cpp: k_extract and k_score<K> compiled from
vs_kernels.cuh under tests/cuda_on_host.h, one OS thread per CUDA thread), against the oracle — the kernels' logic end to
end in the CPU suite.  Texts are packed and hits resolved by the library's host side, exactly as around a GPU scan.
This is test infrastructure: the product has no CPU path, and ptxas / launch geometry / streams / the upload path are what
the tests marked gpu add."""
import os
import struct
import subprocess

import numpy as np
import pytest

import varscot_b200 as V
from oracle import oracle as O
from tests.util import make_case, make_repeat_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_emulator(path, *defines):
    r = subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-Wno-unknown-pragmas", *defines, "-o", path,
                        os.path.join(ROOT, "tests", "cpu_scan_emulator.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return path


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    return build_emulator(str(tmp_path_factory.mktemp("emu") / "cpu_scan_emulator"))


def emulate(exe, tmp_path, text, guides, k, pam=None, tile_words=0, chunk_words=1 << 20):
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    g = np.ascontiguousarray(guides, dtype=np.uint8).reshape(-1, 23)
    with open(inp, "wb") as f:
        f.write(struct.pack("<QIiiIQ", text.n_words, len(g), k, V.pam_code(pam), tile_words, chunk_words))
        f.write(text.bases.tobytes()); f.write(text.masks.tobytes()); f.write(g.tobytes())
    r = subprocess.run([exe, inp, out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    n = struct.unpack("<Q", raw[:8])[0]
    return np.frombuffer(raw[8:8 + 8 * n], dtype=V.HIT_DT).copy()


def rows_of(text, hits, offsets, guides):
    rec, _ = V.resolve_hits(hits, offsets)
    return [(int(r["guide"]), int(r["flag"]), int(r["contig"]), int(r["pos"]), int(r["mm"]),
             V.md_string(text, int(offsets[r["contig"]]) + int(r["pos"]), guides[r["guide"]], (int(r["flag"]) >> 4) & 1)) for r in rec]


def check(exe, tmp_path, case, **kw):
    text = V.PackedText.from_ascii(case.ascii, case.offsets)
    hits = emulate(exe, tmp_path, text, case.guides, case.k, case.pam, **kw)
    exp = O.map_guides(O.text_codes(case.ascii), case.offsets, case.guides, case.k, pam=case.pam).rows()
    assert rows_of(text, hits, case.offsets, case.guides) == exp
    return len(exp)


@pytest.mark.parametrize("k", [0, 1, 3, 4, 5, 6, 7, 8])
def test_emulated_scan_equals_oracle_all_k(emulator, tmp_path, k):
    case = make_case(seed=300 + k, contig_lens=[9000, 45, 45, 23, 22, 0, 46, 4000], n_guides=5, k=k, pam=[None, "AG", "TC"][k % 3])
    assert check(emulator, tmp_path, case) > 0


def test_emulated_scan_chunks_tiles_and_guide_chunks(emulator, tmp_path):
    # chunk borders (halo word), odd tile sizes incl. the unaligned staging path, partial blocks, more guides than one launch
    case = make_case(seed=41, contig_lens=[6000] + [45] * 40 + [3000], n_guides=4, k=4, pam="AG")
    n = check(emulator, tmp_path, case, chunk_words=57, tile_words=16)
    assert n > 0
    assert check(emulator, tmp_path, case, chunk_words=100, tile_words=255) == n
    many = make_case(seed=42, contig_lens=[2500, 45, 45], n_guides=300, k=3)
    assert check(emulator, tmp_path, many) > 0


@pytest.mark.parametrize("n_guides", [29, 31, 32, 63, 157])      # (tests/cpu_kernel_units.cpp sweeps every count for k_score alone)
def test_emulated_scan_every_tail_of_a_warp_slice(emulator, tmp_path, n_guides):
    """A warp's slice of the guide list that holds 29..31 guides pads to a full 32-guide segment (round 2's first form split the
    tail into 16 + 8 + 4 and scored NOTHING for 29..31: found by tools/fuzz_device_code_on_host.py)."""
    case = make_case(seed=500 + n_guides, contig_lens=[1500, 45, 45, 23, 800], n_guides=n_guides, k=3, pam=[None, "AG"][n_guides % 2])
    assert check(emulator, tmp_path, case) > 0
    exp = O.map_guides(O.text_codes(case.ascii), case.offsets, case.guides, case.k, pam=case.pam).rows()
    assert any(r[0] >= (n_guides - 1) // 32 * 32 for r in exp)             # the last slice has hits to lose


@pytest.mark.parametrize("n_guides,guide_pass,ctas,rot", [(100, 37, 2, 0), (157, 64, 5, 3), (70, 128, 1, 40), (33, 4, 3, 1), (300, 129, 4, 2)])
def test_emulated_scan_in_guide_passes(emulator, tmp_path, monkeypatch, n_guides, guide_pass, ctas, rot):
    """scan_engine's guide super-chunks as the kernels see them: launches with guide_base > 0 over a pattern table wider than the
    launch (plain and bucketed), different numbers of persistent CTAs, different rotation periods of the warp roles."""
    monkeypatch.setenv("VS_EMU_GUIDE_PASS", str(guide_pass)); monkeypatch.setenv("VS_EMU_CTAS", str(ctas)); monkeypatch.setenv("VS_EMU_ROT", str(rot))
    case = make_case(seed=1000 + n_guides, contig_lens=[3000, 45, 45, 23, 900], n_guides=n_guides, k=4, pam="AG")
    assert check(emulator, tmp_path, case, chunk_words=40) > 0


@pytest.mark.parametrize("n_bases,n_guides,k,pam", [(150000, 6, 3, None), (90000, 40, 2, "AG"), (60000, 9, 5, None)])
def test_emulated_scan_low_complexity_text_buckets_of_several_batches(emulator, tmp_path, n_bases, n_guides, k, pam):
    """Tandem repeats: thousands of candidates in ONE bucket (several batches of the bucketed store, which random text of test size
    never fills), hits at every repeat (the hit buffer of the emulator regrows)."""
    case = make_repeat_case(seed=77 + n_guides, n_bases=n_bases, n_guides=n_guides, k=k, pam=pam)
    assert check(emulator, tmp_path, case, chunk_words=700) > 2000


def test_emulated_scan_n_runs_and_last_windows(emulator, tmp_path):
    # long N stretches (no candidates for many words: blocks span far), contigs ending on last windows (R4)
    rng = np.random.default_rng(9)
    case = make_case(seed=43, contig_lens=[20000, 45, 23, 45, 46], n_guides=6, k=6)
    asc = bytearray(case.ascii)
    asc[3000:3000 + 9000] = b"N" * 9000
    case.ascii = bytes(asc)
    assert check(emulator, tmp_path, case) > 0
