#!/usr/bin/env python
"""End-to-end timing of the drop-in executables on a synthetic FASTA (run on the GPU box).
usage: tools/cli_bench.py [genome_bases] [n_variants] [n_guides] [k]"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import varscot_b200 as V                      # noqa: E402
from varscot_b200 import synth                # noqa: E402

BIN = os.path.join(ROOT, "build", "read_mapping_build")


def write_fasta(path, text, names, width=70):
    codes = synth.unpack_codes(text)
    asc = np.frombuffer(b"ACGTN", dtype=np.uint8)[codes]
    with open(path, "wb") as f:
        for i, nm in enumerate(names):
            s, e = int(text.offsets[i]), int(text.offsets[i + 1])
            f.write(b">" + nm.encode() + b"\n")
            body = asc[s:e]
            full = (len(body) // width) * width
            if full:
                lines = np.empty((full // width, width + 1), dtype=np.uint8)
                lines[:, :width] = body[:full].reshape(-1, width)
                lines[:, width] = 10
                f.write(lines.tobytes())
            if len(body) > full:
                f.write(body[full:].tobytes() + b"\n")


def main():
    gb = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000_000
    nv = int(float(sys.argv[2])) if len(sys.argv) > 2 else 500_000
    ng = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    g = synth.synth_genome(11, gb, 24, 0.05)
    s = synth.synth_variant_segments(g, 12, nv)
    guides = synth.synth_guides(13, ng)
    d = tempfile.mkdtemp(prefix="vscli_")
    out = {}
    for label, text, names in (("ref", g, g.names), ("snp", s, [f"chr1_{i}_REF" for i in range(s.n_contigs)])):
        fa = os.path.join(d, label + ".fa")
        t = time.time(); write_fasta(fa, text, names); out[label + "_fasta_write_s"] = round(time.time() - t, 2)
        rf = os.path.join(d, "guides.fa")
        with open(rf, "w") as f:
            for i, gd in enumerate(guides):
                f.write(f">g{i}\n{''.join('ACGT'[b] for b in gd)}\n")
        t = time.time(); r = subprocess.run([os.path.join(BIN, "bidir_index"), "-G", fa, "-I", os.path.join(d, label)], capture_output=True, text=True)
        out[label + "_index_s"] = round(time.time() - t, 2); assert r.returncode == 0, r.stderr
        env = dict(os.environ, VARSCOT_VERBOSE="1")
        for ngpu in sorted({1, V.device_count()}):
            env["VARSCOT_GPUS"] = str(ngpu)
            sam = os.path.join(d, f"{label}_{ngpu}.sam")
            t = time.time()
            r = subprocess.run([os.path.join(BIN, "bidir_mapping"), "-G", fa, "-I", os.path.join(d, label), "-R", rf, "-M", str(k), "-T", "8", "-O", sam],
                               capture_output=True, text=True, env=env)
            out[f"{label}_mapping_{ngpu}gpu_s"] = round(time.time() - t, 2); assert r.returncode == 0, r.stderr
            out[f"{label}_mapping_{ngpu}gpu_note"] = " | ".join(r.stderr.strip().splitlines()[-2:]) if r.stderr.strip() else ""
            out[f"{label}_sam_lines_{ngpu}gpu"] = sum(1 for _ in open(sam))
        sams = [open(os.path.join(d, f"{label}_{n}.sam"), "rb").read() for n in sorted({1, V.device_count()})]
        out[label + "_multi_gpu_sam_identical"] = all(x == sams[0] for x in sams)
        # library result on the same text for the record count
        with V.ScanContext(0) as ctx:
            hits, st = ctx.scan_text(text, guides, k)
        out[label + "_library_hits"] = len(hits)
        out[label + "_bases"] = text.n_bases
    print(out)


if __name__ == "__main__":
    main()
