#!/usr/bin/env python
"""Differential fuzzing of the scan's device code WITHOUT a GPU: random texts (long contigs, swarms of 45-mers, contigs around the
window length, N densities), guides, k, PAMs, tile and chunk sizes run through tests/cpu_scan_emulator.cpp (k_extract + k_score
from their real source, on the host) and compared record by record with the oracle.

usage: tools/fuzz_device_code_on_host.py [SEED] [SECONDS] [-DMACRO ...]      e.g.  ... 7 300 -DVS_EX_HALF=1
Round 1: 1 329 cases with the default build and 564 with -DVS_EX_HALF=1, 74 k records, no mismatch.
Round 2 (plain, whole-store and bucketed scans; up to 170 guides per case, guide passes, CTA counts, tandem-repeat texts): the first
run with more than 6 guides found the 29..31-guide tail bug of k_score; after the fix 6 768 cases / 15.0 M records, no mismatch."""
import os
import pathlib
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import varscot_b200 as V                                                      # noqa: E402
from oracle import oracle as O                                                # noqa: E402
from tests.test_device_code_on_host import build_emulator, emulate, rows_of  # noqa: E402
from tests.util import make_case, make_repeat_case                                              # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("-D")]
    defines = [a for a in sys.argv[1:] if a.startswith("-D")]
    rng = np.random.default_rng(int(args[0]) if args else 0)
    t_end = time.time() + (float(args[1]) if len(args) > 1 else 600.0)
    tmp = pathlib.Path(tempfile.mkdtemp(prefix="vs_fuzz_"))
    exe = build_emulator(str(tmp / "emu"), *defines)
    n_cases = n_rows = 0
    while time.time() < t_end:
        lens = []
        for kd in rng.integers(0, 5, int(rng.integers(1, 8))):
            if kd == 0:
                lens.append(int(rng.integers(2000, 30000)))
            elif kd == 1:
                lens += [45] * int(rng.integers(1, 200))
            elif kd == 2:
                lens += [int(x) for x in rng.integers(0, 60, int(rng.integers(1, 30)))]
            elif kd == 3:
                lens.append(int(rng.integers(23, 600)))
            else:
                lens += [23, 22, 24, 0]
        k = int(rng.integers(0, 9))
        pam = [None, "AG", "TT", "CC", "GA", "CT"][int(rng.integers(0, 6))]
        seed = int(rng.integers(0, 1 << 30))
        # guide counts: mostly a handful; now and then enough for 32-guide segments, a folded tail, or more than one CTA's worth
        # (the bucketed kernel deals whole classes of a bucket in 32-guide and 4-guide segments)
        ng = int(rng.integers(1, 7))
        r = rng.random()
        if r < 0.25:
            ng = int(rng.integers(28, 72))
        elif r < 0.35:
            ng = int(rng.integers(120, 170))
        if ng > 8:                                            # keep the emulation (one OS thread per CUDA thread) short
            total = 0
            kept = []
            for L in lens:
                if total + L > (6000 if ng > 100 else 12000):
                    L = max(0, (6000 if ng > 100 else 12000) - total)
                kept.append(L)
                total += L
            lens = kept
        case_nfrac = float(rng.choice([0, 0.002, 0.02]))
        case_gpam = [None, "GG", "GG", "AG"][int(rng.integers(0, 4))]
        if rng.random() < 0.12:                               # low-complexity text: buckets of several batches, dense hits
            lens = ["repeat", int(rng.integers(20000, 120000))]
            ng = min(ng, 48)
            k = min(k, 5)
            case = make_repeat_case(seed, lens[1], ng, k, pam=pam if pam in (None, "AG", "TT") else None)
            pam = case.pam
        else:
            case = make_case(seed, lens, ng, k, pam=pam, n_frac=case_nfrac, guide_pam=case_gpam)
        text = V.PackedText.from_ascii(case.ascii, case.offsets)
        tile = int(rng.choice([0, 0, 8, 9, 16, 33, 100, 255, 256]))
        chunk = int(rng.choice([1 << 20, 1 << 20, 7, 64, 65, 1000]))
        # knobs of the emulator: guides per pass of the resident scans (guide_base > 0), persistent CTAs, rotation period of the warp roles
        knobs = dict(VS_EMU_GUIDE_PASS=str(int(rng.choice([0, 0, 4, 32, 37, 64, 128, 129]))), VS_EMU_CTAS=str(int(rng.choice([1, 2, 3, 3, 5, 5, 8, 8, 64, 300]))),
                     VS_EMU_ROT=str(int(rng.choice([0, 1, 3, 40]))))
        os.environ.update(knobs)
        try:
            hits = emulate(exe, tmp, text, case.guides, k, pam, tile_words=tile, chunk_words=chunk)
        except AssertionError as e:
            print("EMULATOR FAILED", dict(seed=seed, lens=lens, ng=ng, k=k, pam=pam, tile=tile, chunk=chunk, n_frac=case_nfrac, guide_pam=case_gpam, **knobs), str(e)[-300:], flush=True)
            return 1
        exp = O.map_guides(O.text_codes(case.ascii), case.offsets, case.guides, k, pam=pam).rows()
        got = rows_of(text, hits, case.offsets, case.guides)
        if got != exp:
            print("MISMATCH", dict(seed=seed, lens=lens, ng=ng, k=k, pam=pam, tile=tile, chunk=chunk, n_frac=case_nfrac, guide_pam=case_gpam, got=len(got), exp=len(exp), **knobs), flush=True)
            return 1
        n_cases += 1
        n_rows += len(exp)
    print("ok", n_cases, "cases", n_rows, "records")
    return 0


if __name__ == "__main__":
    sys.exit(main())
