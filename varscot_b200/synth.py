"""Synthetic workloads as PackedText objects (the bit planes come from synth_core.py, numpy only).

Host-side tooling for bench.py and the tests; no search arithmetic.
"""
from __future__ import annotations

import numpy as np

from . import synth_core as core
from .mapper import PackedText
from .synth_core import Planes, contig_ladder, slice_offsets, synth_guides, _gather  # noqa: F401


class SynthText(PackedText):
    """A PackedText that keeps its N and contig-end planes (unpack_codes needs the N plane)."""
    hi = property(lambda self: self.bases["hi"])
    lo = property(lambda self: self.bases["lo"])
    nm = property(lambda self: self._nm)
    em = property(lambda self: self._em)


def from_planes(p: Planes) -> PackedText:
    t = PackedText.from_planes(p.hi, p.lo, p.nm, p.em, p.offsets, p.n_bases, p.names)
    t.__class__ = SynthText
    t._nm, t._em = p.nm, p.em
    return t


def _planes_of(t) -> Planes:
    return Planes(t.bases["hi"], t.bases["lo"], t._nm, t._em, t.offsets, t.n_bases, t.names)


def synth_genome(seed: int, total_bases: int, n_contigs: int = 24, n_frac: float = 0.05) -> PackedText:
    return from_planes(core.synth_genome(seed, total_bases, n_contigs, n_frac))


def synth_variant_segments(genome, seed: int, n_variants: int, **kw) -> PackedText:
    return from_planes(core.synth_variant_segments(_planes_of(genome), seed, n_variants, **kw))


def concat_texts(a, b) -> PackedText:
    return from_planes(core.concat_texts(_planes_of(a), _planes_of(b)))


def build_workload(genome_bases: int, n_variants: int, scale: float = 1.0, seed_genome: int = 11, seed_variants: int = 12) -> PackedText:
    return from_planes(core.build_workload(genome_bases, n_variants, scale, seed_genome, seed_variants))


def unpack_codes(text, start: int = 0, n: int | None = None) -> np.ndarray:
    """Dna5 codes (0..3, 4 = N) of bases [start, start+n): what the CPU oracle consumes. start % 32 == 0."""
    return core.unpack_codes(_planes_of(text), start, n)
