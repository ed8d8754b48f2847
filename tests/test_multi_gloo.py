"""N > 1 host logic on CPU: two gloo ranks each own one shard of the packed text (vs_shard_bounds), produce the hits
of their own window starts (here with the oracle standing in for the device, reading the same halo the device reads),
gather them on rank 0 and resolve; the merged records must equal the unpartitioned result.  No data-path collective
exists in the product; gloo only carries the test's gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import varscot_b200 as V
    from oracle import oracle as O
    from tests.util import make_case
    case = make_case(77, [30000, 45, 45, 45, 8211, 0, 23, 9000], 6, 6)
    text = V.PackedText.from_ascii(case.ascii, case.offsets)
    codes = O.text_codes(case.ascii)
    b = V.shard_bounds(text.n_words, world)
    s0, s1 = int(b[rank]) * 32, min(int(b[rank + 1]) * 32, text.n_bases)
    hits = np.zeros(0, dtype=V.HIT_DT)
    if s1 > s0:
        # the shard reads its owned words plus a halo; one extra base keeps owned windows away from the artificial end
        e = min(text.n_bases, s1 + 23)
        off = np.concatenate([[0], case.offsets[(case.offsets > s0) & (case.offsets < e)].astype(np.int64) - s0, [e - s0]]).astype(np.uint64)
        r = O.map_guides(codes[s0:e], off, case.guides, case.k)
        gpos = off[r.contig].astype(np.int64) + r.pos.astype(np.int64) + s0
        own = gpos < s1
        hits = np.zeros(int(own.sum()), dtype=V.HIT_DT)
        hits["pos"] = gpos[own]
        hits["info"] = (r.guide[own].astype(np.uint32) << 8) | (((r.flag[own] & 16) >> 4).astype(np.uint32) << 7) | r.mm[own]
    gathered = [None] * world
    dist.all_gather_object(gathered, hits.tobytes())
    t = torch.tensor([float(len(hits))])
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        allh = np.concatenate([np.frombuffer(x, dtype=V.HIT_DT) for x in gathered])
        assert len(allh) == int(t.item())
        rec, _ = V.resolve_hits(allh, case.offsets)
        whole = O.map_guides(codes, case.offsets, case.guides, case.k)
        got = [(int(x["guide"]), int(x["flag"]), int(x["contig"]), int(x["pos"]), int(x["mm"])) for x in rec]
        q.put((got == [x[:5] for x in whole.rows()], len(got), [len(np.frombuffer(x, dtype=V.HIT_DT)) for x in gathered]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_hits_merge_to_unpartitioned_result(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, n, per = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and n > 10
    assert sum(1 for x in per if x > 0) >= 2        # more than one rank contributed hits


def _worker_exchange(rank, world, port, q):
    """The N > 1 host path of bench.py / the executables: every rank hands its SORTED, resolved hit list (here built from the
    oracle's hits; on the GPU box vs_scan_resolved delivers it) to rank 0 through shared memory (bench.HostExchange), rank 0
    merges with vs_merge_resolved."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import varscot_b200 as V
    import bench
    from oracle import oracle as O
    from tests.util import make_case
    case = make_case(78, [30000] + [45] * 400 + [8211, 0, 23, 9000], 6, 6)
    text = V.PackedText.from_ascii(case.ascii, case.offsets)
    codes = O.text_codes(case.ascii)
    whole = O.map_guides(codes, case.offsets, case.guides, case.k)
    gpos = case.offsets[whole.contig].astype(np.int64) + whole.pos.astype(np.int64)
    b = V.shard_bounds(text.n_words, world)
    own = (gpos >= int(b[rank]) * 32) & (gpos < int(b[rank + 1]) * 32)          # the window starts this rank's shard owns
    loc = np.zeros(int(own.sum()), dtype=V.LOC_DT)
    strand = ((whole.flag[own] & 16) >> 4).astype(np.uint64)
    loc["key"] = (whole.guide[own].astype(np.uint64) << np.uint64(49)) | (strand << np.uint64(48)) | \
                 ((whole.contig[own].astype(np.uint64) & np.uint64(0xFFFF)) << np.uint64(32)) | whole.pos[own].astype(np.uint64)
    loc["contig"] = whole.contig[own]
    loc["info"] = (whole.guide[own].astype(np.uint32) << 8) | (strand.astype(np.uint32) << 7) | whole.mm[own]
    loc = loc[np.argsort(loc["key"], kind="stable")]
    ex = bench.HostExchange(V, world, rank, 1 << 14, pin=False, tag=f"test_{port}")
    dist.barrier()
    ex.attach()
    import time
    results = []
    half = 1 << 13
    # several "phases" of steps as bench.py runs them (end to end, resident genome with two lists per rank, the config-4 block):
    # later steps publish LATE and with fewer hits, so a rank 0 that mistook an old step for the current one would merge stale lists
    for step in range(1, 8):
        ex.begin()
        sub = loc if step % 3 else loc[: len(loc) // 2]          # the list of this step
        if rank and step >= 3:
            time.sleep(0.05 * rank)
        if step % 2:
            ex.mine[: len(sub)] = sub
            ex.publish(len(sub))
        else:                                                    # two lists per rank: even / odd entries
            a, b2 = sub[0::2], sub[1::2]
            ex.mine[: len(a)] = a
            ex.mine[half: half + len(b2)] = b2
            ex.publish(len(a), len(b2))
        if rank == 0:
            rec, coll = V.merge_resolved(ex.collect(split=0 if step % 2 else half), threads=2)
            ex.done()
            results.append((step, len(rec), [(int(x["guide"]), int(x["flag"]), int(x["contig"]), int(x["pos"]), int(x["mm"])) for x in rec]))
        else:
            ex.wait_done()
    counts = [None] * world
    dist.all_gather_object(counts, (len(loc), len(loc) // 2))
    if rank == 0:
        full = [x[:5] for x in whole.rows()]
        ok = True
        for step, n, got in results:
            if step % 3:
                ok = ok and got == full
            else:
                ok = ok and n == sum(c[1] for c in counts) and n < len(full)
        q.put((ok, len(full), [c[0] for c in counts]))
    dist.barrier()
    ex.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shared_memory_exchange_and_merge(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker_exchange, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, n, per = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and n > 10
    assert sum(1 for x in per if x > 0) >= 2
