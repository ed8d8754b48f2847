"""GPU tests written AFTER round 2's GPU budget was spent: they have passed on the host emulation of the device code
(tests/test_device_code_on_host.py has their twins) but have never run on hardware.  The file sorts last so that, under `pytest -x`,
everything that HAS run on a B200 runs first.

* guide counts whose last warp slice holds 29..31 guides: `k_score` lost the hits of that slice until the fuzzer found it;
* low-complexity (tandem repeat) texts: buckets of the bucketed index that span tens of batches, a million hits per scan;
* random sequences of uploads / scans / option changes on one context (the context as a state machine), results only."""
import numpy as np
import pytest

from tests.test_gpu_parity import oracle_rows, rows_from_records
from tests.util import make_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_guides", [29, 30, 31, 61, 63, 95, 127, 157])
def test_guide_counts_whose_last_warp_slice_holds_29_to_31_guides(n_guides):
    """k_score pads a slice of 29..31 guides to a full 32-guide segment (the first form of round 2 split tails into 16 + 8 + 4 and
    scored nothing for these counts): streamed scan, resident plain index and bucketed index against the oracle."""
    import varscot_b200 as V
    from varscot_b200 import _lib
    case = make_case(seed=900 + n_guides, contig_lens=[30000, 45, 45, 23, 8000], n_guides=n_guides, k=4, pam=[None, "AG"][n_guides % 2])
    text = V.PackedText.from_ascii(case.ascii, case.offsets)
    exp = oracle_rows(case.ascii, case.offsets, case.guides, 4, case.pam)
    with V.ScanContext(0) as ctx:
        ctx.set_chunk_words(500)
        ctx.set_option(_lib.VS_OPT_BUCKET_INDEX, 0)
        h1, st1 = ctx.scan_resolved(case.guides, 4, pam=case.pam, text=text)        # streamed: k_score chunk by chunk
        h2, st2 = ctx.scan_resolved(case.guides, 4, pam=case.pam)                   # resident plain index: one k_score launch
        ctx.set_option(_lib.VS_OPT_BUCKET_INDEX, 2)
        h3, st3 = ctx.scan_resolved(case.guides, 4, pam=case.pam)                   # bucketed index
        assert (st1.index_reused, st2.index_reused, st3.index_reused) == (0, 1, 2)
        for h in (h1, h2, h3):
            rec, _ = V.merge_resolved([h.copy()])
            assert rows_from_records(text, rec, case.offsets, case.guides) == exp
    assert any(r[0] >= (n_guides - 1) // 32 * 32 for r in exp)             # the last slice has hits to lose


@pytest.mark.parametrize("n_bases,n_guides,k,pam", [(2_000_000, 40, 3, None), (600_000, 130, 4, "AG")])
def test_low_complexity_text_buckets_of_many_batches(n_bases, n_guides, k, pam):
    """Tandem repeats (tests/util.py: make_repeat_case): one bucket of the bucketed index holds tens of batches and every repeat is a
    hit for many guides — what uniform random text only produces at genome size.  Streamed, resident plain and bucketed scans must
    deliver the same sorted list, and the records must equal the oracle's."""
    import varscot_b200 as V
    from varscot_b200 import _lib
    from tests.util import make_repeat_case
    case = make_repeat_case(seed=31 + n_guides, n_bases=n_bases, n_guides=n_guides, k=k, pam=pam)
    from oracle import oracle as O
    text = V.PackedText.from_ascii(case.ascii, case.offsets)
    exp = O.map_guides(O.text_codes(case.ascii), case.offsets, case.guides, k, pam=pam)
    n = len(exp.guide)
    assert n > 1_000_000
    with V.ScanContext(0) as ctx:
        ctx.set_option(_lib.VS_OPT_BUCKET_INDEX, 0)
        h1, st1 = ctx.scan_resolved(case.guides, k, pam=pam, text=text, cap=1 << 22)
        h2, st2 = ctx.scan_resolved(case.guides, k, pam=pam, cap=1 << 22)
        ctx.set_option(_lib.VS_OPT_BUCKET_INDEX, 2)
        h3, st3 = ctx.scan_resolved(case.guides, k, pam=pam, cap=1 << 22)
        h4, st4 = ctx.scan_resolved(case.guides, k, pam=pam, cap=1 << 22)
        assert (st1.index_reused, st2.index_reused, st3.index_reused, st4.index_reused) == (0, 1, 2, 2)
        assert len(h1) == n and h1.tobytes() == h2.tobytes() == h3.tobytes() == h4.tobytes()
        rec, _ = V.merge_resolved([h4.copy()], threads=4)
    # a million records: compared as arrays (order, FLAGs and counts included), the MD strings on a sample
    for name in ("guide", "flag", "contig", "pos", "mm"):
        assert np.array_equal(rec[name].astype(np.int64), np.asarray(getattr(exp, name)).astype(np.int64)), name
    sample = rec[:: max(1, n // 1500)]
    rows = exp.rows()
    assert rows_from_records(text, sample, case.offsets, case.guides) == rows[:: max(1, n // 1500)]


@pytest.mark.parametrize("seed", range(6))
def test_random_sequences_of_scans_on_one_context(seed):
    """A context is a state machine (resident shard, resident plain / bucketed index per PAM set, buffers that grow): random
    sequences of uploads of different shards of different texts, streamed and resident scans with changing guide sets (counts that
    exercise every segment shape), k, PAM sets and options.  After EVERY scan the records of the shard must be the oracle's (as a set) —
    only results are asserted, no statistics."""
    import varscot_b200 as V
    from varscot_b200 import _lib
    from oracle import oracle as O
    from tests.util import make_repeat_case
    rng = np.random.default_rng(4000 + seed)
    cases = [make_case(seed=4100 + seed, contig_lens=[40000, 45, 45, 23, 0, 46, 12000] + [45] * 150, n_guides=300, k=5),
             make_case(seed=4200 + seed, contig_lens=[9000] * 3, n_guides=300, k=5, pam="AG", guide_pam="AG"),
             make_repeat_case(seed=4300 + seed, n_bases=120000, n_guides=300, k=4)]
    texts = [V.PackedText.from_ascii(c.ascii, c.offsets) for c in cases]
    codes = [O.text_codes(c.ascii) for c in cases]
    counts = [1, 4, 5, 29, 31, 33, 61, 64, 100, 127, 130, 157, 300]

    def expected(ti, g0, ng, k, pam, first, words):
        """oracle records of the text whose window starts in words [first, first + words), as (guide, strand, contig, pos, mm)"""
        c = cases[ti]
        r = O.map_guides(codes[ti], c.offsets, c.guides[g0:g0 + ng], k, pam=pam)
        gpos = c.offsets[r.contig].astype(np.int64) + r.pos.astype(np.int64)
        own = (gpos >= first * 32) & (gpos < (first + words) * 32)
        return list(zip(r.guide[own].tolist(), ((r.flag[own] & 16) >> 4).tolist(), r.contig[own].tolist(), r.pos[own].tolist(), r.mm[own].tolist()))

    def got_rows(h):
        rec, _ = V.merge_resolved([h.copy()])
        return list(zip(rec["guide"].tolist(), ((rec["flag"] & 16) >> 4).tolist(), rec["contig"].tolist(), rec["pos"].tolist(), rec["mm"].tolist()))

    with V.ScanContext(0) as ctx:
        resident = None                                       # (text index, first word, words)
        n_scans = 0
        for step in range(30):
            op = rng.choice(["upload", "resident", "resident", "resident", "streamed", "options", "drop"])
            if op == "options":
                ctx.set_option(_lib.VS_OPT_KEEP_INDEX, int(rng.integers(0, 2)))
                ctx.set_option(_lib.VS_OPT_BUCKET_INDEX, int(rng.integers(0, 3)))
                ctx.set_option(_lib.VS_OPT_HIT_CAPACITY, int(rng.choice([0, 0, 64, 5000])))
                ctx.set_chunk_words(int(rng.choice([500, 2000, 1 << 20])))
                continue
            if op == "drop":
                ctx.drop_index()
                continue
            if op in ("upload", "streamed") or resident is None:
                ti = int(rng.integers(0, len(cases)))
                nw = texts[ti].n_words
                shards = int(rng.integers(1, 4))
                i = int(rng.integers(0, shards))
                first, words = nw * i // shards, nw * (i + 1) // shards - nw * i // shards
                if op != "streamed":
                    ctx.upload(texts[ti], first, words)
                    resident = (ti, first, words)
                    if op == "upload":
                        continue
            ng = int(rng.choice(counts))
            k = int(rng.integers(0, 6))
            if op == "streamed":
                pam = rng.choice([None, "AG", "TT"])
                g0 = int(rng.integers(0, 300 - ng + 1))
                h, _ = ctx.scan_resolved(cases[ti].guides[g0:g0 + ng], k, pam=pam, text=texts[ti], first_word=first, n_words=words, cap=1 << 22)
                resident = (ti, first, words)
            else:
                ti, first, words = resident
                pam = rng.choice([None, None, "AG", "TT"])
                g0 = int(rng.integers(0, 300 - ng + 1))
                h, _ = ctx.scan_resolved(cases[ti].guides[g0:g0 + ng], k, pam=pam, cap=1 << 22)
            # (as sets: the emission order and the FLAGs of a pass depend on ALL its records — the best one is written last — so a shard's
            # list is not a sub-sequence of the whole text's; order and FLAGs are what the tests on whole texts check)
            assert sorted(got_rows(h)) == sorted(expected(ti, g0, ng, k, pam, first, words)), (seed, step, op, ti, first, words, g0, ng, k, pam)
            n_scans += 1
        assert n_scans >= 8
