"""GPU tests written AFTER round 2's GPU budget was spent: they have passed on the host emulation of the device code
(tests/test_device_code_on_host.py has their twins) but have never run on hardware.  The file sorts last so that, under `pytest -x`,
everything that HAS run on a B200 runs first.

* guide counts whose last warp slice holds 29..31 guides: `k_score` lost the hits of that slice until the fuzzer found it;
* low-complexity (tandem repeat) texts: buckets of the bucketed index that span tens of batches, a million hits per scan."""
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
