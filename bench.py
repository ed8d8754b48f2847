#!/usr/bin/env python
"""bench.py — guide·Gbp/s of the off-target scan (BASELINE.json metric) on 1..8 B200.

A step = one full scan of all guides (both strands, <= k mismatches, PAM on the genome) over this
rank's synthetic text: reference genome + variant haplotype segments ("SNP genome").
  value : whole-job guide·Gbp/s with the packed text resident in HBM; timed with CUDA events on the
          library's stream from the first kernel to the last hit in host memory, max over ranks
  e2e   : same metric through the C-ABI call with HOST buffers: H2D of the packed text + guides, scan,
          D2H of the hits, every step
  roofline : the scoring kernel against the MEASURED alu-pipe LOP3 rate, in the yardstick of SURVEY.md 8d
             (4.0 LOP3 per guide·bp for a dense bit-sliced scan) and in executed instructions
  cpu_baseline : the CPU oracle (a port: linear XOR/popcount scan, not SeqAn's FM index) on a bounded sample
`--impl reference` times that CPU oracle alone (the reference binary needs SeqAn, absent here).
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C_ALG = 4.0            # yardstick LOP3 per guide·bp: 64 per 32 starts per guide per strand


def csa_ops(n, init=False):
    """LOP3-class ops of popcount_planes<n, init> in vs_kernels.cuh: column compression, FA = 2 ops, HA = 2 ops."""
    ops, cols = 0, n
    for w in range(5):
        left = cols + (1 if init else 0)
        nxt = 0
        while left >= 2:
            left -= 2 if left >= 3 else 1
            ops += 2
            nxt += 1
        cols = nxt
    return ops


def score_ops(k):
    """(LOP3, LDS) k_score executes per (32-candidate block, guide) in stage A, and the extra of stage B."""
    pa = 23 if k >= 8 else 7 + 2 * k
    a = (csa_ops(pa) + 2, pa)
    # stage B scores its slots from raw planes: 2 LDS + 2 LOP3 per slot, then the adders onto stage A's count
    b = (csa_ops(23 - pa, True) + 2 + 2 * (23 - pa), 2 * (23 - pa)) if pa < 23 else (0, 0)
    return a, b

CONFIGS = {
    # id: (description, genome bases, variants, guides, k, extra PAM)
    1: ("cfg1: 10 guides vs 50 Mbp + 10k SNVs, <=4 mm", 50_000_000, 10_000, 10, 4, None),
    2: ("cfg2: 100 guides vs 3.1 Gbp, reference only, <=4 mm", 3_100_000_000, 0, 100, 4, None),
    3: ("cfg3: 100 guides vs 3.1 Gbp + 5M variants, <=6 mm", 3_100_000_000, 5_000_000, 100, 6, None),
    4: ("cfg4: 1000 guides vs 3.1 Gbp + 5M variants, <=6 mm, NGG+NAG", 3_100_000_000, 5_000_000, 1000, 6, "AG"),
    5: ("cfg5: 10000 guides vs 3.1 Gbp + 5M variants, <=8 mm", 3_100_000_000, 5_000_000, 10000, 8, None),
}


def build_text(cfg, rank, scale=1.0):
    from varscot_b200 import synth
    _, gbases, nvar, _, _, _ = cfg
    gbases = int(gbases * scale)
    nvar = int(nvar * scale)
    g = synth.synth_genome(11 + 1000 * rank, gbases, 24, 0.05)
    if nvar > 0:
        s = synth.synth_variant_segments(g, 12 + 1000 * rank, nvar)
        return synth.concat_texts(g, s)
    return g


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled in the background (B200_PROFILING.md recipe)."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.rows = []
        self.device = device
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(device)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                self.rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[4], f[5], f[6], f[7]))
            except Exception:
                pass

    def stop(self, t0=None, t1=None):
        if not self.p:
            return None
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0 - 0.05) and (t1 is None or r[0] <= t1 + 0.05)]
        window = "timed"
        if not rows:
            rows, window = self.rows, "whole_run"
        if not rows:
            # the looping sampler produced nothing (pipe buffering / slow start): take one synchronous sample now
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.device)],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
                f = [x.strip() for x in out.split(",")]
                rows, window = [(time.time(), float(f[1]), float(f[2]), float(f[3]), f[4], f[5], f[6], f[7])], "after_timed"
            except Exception:
                return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        sm = sorted(r[1] for r in rows)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": rows[0][2], "power_w_max": max(r[3] for r in rows),
                "reasons": reasons, "samples": len(rows), "window": window}


def bind_to_gpu_numa(local):
    """Pin this process to the CPUs next to its GPU (NVML affinity) BEFORE allocating page-locked buffers, so that the
    H2D source memory is first-touched on the GPU's NUMA node; matters when 8 ranks upload at once."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def barrier(world, local):
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()


def all_max(x, world, local):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(x, world, local):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def cpu_sample(text, guides, k, pam, target_s=12.0, max_bases=1 << 30):
    """Time the oracle (all host threads) on a bounded, word-aligned prefix of the text; return its records too."""
    from oracle import oracle as O
    from varscot_b200 import synth
    probe = min(text.n_bases // 32 * 32, 16 << 20)
    codes = synth.unpack_codes(text, 0, probe)
    off = synth.slice_offsets(text, 0, probe)
    t = time.perf_counter()
    O.scan_count(codes, off, guides, k, pam)
    dt = max(time.perf_counter() - t, 1e-4)
    n = int(min(max_bases, text.n_bases, max(probe, probe * target_s / dt))) // 32 * 32
    codes = synth.unpack_codes(text, 0, n)
    off = synth.slice_offsets(text, 0, n)
    t = time.perf_counter()
    rec = O.map_guides(codes, off, guides, k, pam=pam)
    dt = time.perf_counter() - t
    return n, dt, rec, off, O.num_procs()


def run_reference(args, cfg, world, rank):
    """--impl reference: the CPU oracle port on the host cores (SeqAn's bidir_mapping cannot be built here)."""
    if rank != 0:
        return
    from oracle import oracle as O
    from varscot_b200 import synth
    desc, gbases, nvar, ng, k, pam = cfg
    guides = synth.synth_guides(13, ng)
    # bounded sample with the composition of the full workload; sized so that the whole run takes a few minutes
    budget_s = float(os.environ.get("VARSCOT_BENCH_BUDGET_S", "150")) / max(1, args.steps + args.warmup)
    scale = min(1.0, (512 << 20) / gbases)
    text = build_text(cfg, 0, scale)
    n = text.n_bases // 32 * 32
    codes = synth.unpack_codes(text, 0, n)
    off = synth.slice_offsets(text, 0, n)
    t = time.perf_counter(); O.scan_count(codes[: 8 << 20], synth.slice_offsets(text, 0, 8 << 20), guides, k, pam); dt = max(time.perf_counter() - t, 1e-4)
    want = int(min(n, max(8 << 20, (8 << 20) * budget_s / dt))) // 32 * 32
    codes, off = codes[:want], synth.slice_offsets(text, 0, want)
    for _ in range(args.warmup):
        O.scan_count(codes, off, guides, k, pam)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.scan_count(codes, off, guides, k, pam)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    val = ng * want / dt / 1e9
    cores = O.num_procs()
    sample = f"first {want} bases of a {scale:.4f}-scale copy of the workload, all {ng} guides, per step"
    print(json.dumps({
        "impl": "reference", "metric": "guide_Gbp_per_s", "value": val, "unit": "guide*Gbp/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 xor+popcount", "data": "synthetic", "config": {"workload": desc, "k": k, "guides": ng, "extra_pam": pam},
        "cpu_baseline": {"value": val, "unit": "guide*Gbp/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "guide*Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle port (linear 2-bit XOR/popcount scan, OpenMP); the reference's SeqAn FM-index binary cannot be built: SeqAn 2.4.0rc2 is not vendored",
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the text (quick runs only; invalid as a bench number)")
    ap.add_argument("--guides", type=int, default=0, help="override the number of guides")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the contract): one text of the configured size per rank; strong: ONE text sharded over the ranks by vs_shard_bounds")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    cfg = list(CONFIGS[args.config])
    if args.guides:
        cfg[3] = args.guides
    cfg = tuple(cfg)
    world, rank, local = dist_setup(args.gpus)
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = bind_to_gpu_numa(local) if args.impl == "ours" else 0
    if args.impl == "reference":
        run_reference(args, cfg, world, rank)
        return

    import varscot_b200 as V
    from varscot_b200 import synth
    desc, gbases, nvar, ng, k, pam = cfg
    guides = synth.synth_guides(13, ng)
    strong = args.scaling == "strong" and world > 1
    t_gen = time.perf_counter()
    text = build_text(cfg, 0 if strong else rank, args.scale)
    t_gen = time.perf_counter() - t_gen
    B = text.n_bases
    nw = text.n_words
    # strong scaling: every rank holds the same text and owns one shard of its window starts (no collective anywhere)
    shard_first, shard_words = 0, nw
    if strong:
        sb = V.shard_bounds(nw, world)
        shard_first, shard_words = int(sb[rank]), int(sb[rank + 1] - sb[rank])
    text.pin()                                  # page-locked host buffers: what a caller of the C ABI would hand in
    ctx = V.ScanContext(local)
    peak_lop3 = peak_lds = None
    if rank == 0:
        peak_lop3, peak_lds = ctx.measure_int_peaks()

    # ---- resident-text scan ----------------------------------------------------------------------
    ctx.upload(text, shard_first, shard_words)
    import ctypes as C
    from varscot_b200 import _lib
    pinned = []

    def pinned_hits(n):                          # page-locked result buffer, as a caller expecting many hits would use
        p = _lib.lib().vs_host_alloc(n * 8)
        if not p:
            raise RuntimeError("vs_host_alloc failed")
        pinned.append(p)
        return np.frombuffer((C.c_uint8 * (n * 8)).from_address(p), dtype=V.HIT_DT, count=n)

    cap = 1 << 22
    hits_buf = pinned_hits(cap)
    for _ in range(args.warmup):
        hits, st = ctx.scan(guides, k, pam=pam, out=hits_buf)
        if len(hits) > cap:
            cap = int(len(hits) * 1.2); hits_buf = pinned_hits(cap)
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.15)
    barrier(world, local)
    t0_wall = time.time(); t0 = time.perf_counter()
    dev_ms = score_ms = extract_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        hits, st = ctx.scan(guides, k, pam=pam, out=hits_buf)
        dev_ms += st.total_ms; score_ms += st.score_ms; extract_ms += st.extract_ms
        launches += st.launches
    barrier(world, local)
    wall_ms = (time.perf_counter() - t0) * 1e3
    t1_wall = time.time()
    n_hits = len(hits)
    ms_step = all_max(dev_ms / args.steps, world, local)
    wall_step = all_max(wall_ms / args.steps, world, local)
    # guide·bp per step over all ranks: the shards of one text (strong) or one whole text per rank (weak)
    units = float(ng) * B if strong else all_sum(float(ng) * B, world, local)
    value = units / (ms_step * 1e-3) / 1e9

    # ---- end to end: host buffers in, hits out, every step ------------------------------------------
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            ctx.scan_text(text, guides, k, pam=pam, out=hits_buf, first_word=shard_first, n_words=shard_words)
        barrier(world, local)
        t0 = time.perf_counter()
        e_dev = 0.0
        for _ in range(args.steps):
            h2, st2 = ctx.scan_text(text, guides, k, pam=pam, out=hits_buf, first_word=shard_first, n_words=shard_words)
            e_dev += st2.total_ms
        barrier(world, local)
        e_ms = all_max((time.perf_counter() - t0) * 1e3 / args.steps, world, local)
        e2e = {"value": units / (e_ms * 1e-3) / 1e9, "unit": "guide*Gbp/s", "ms_per_step": e_ms, "device_ms_per_step": e_dev / args.steps,
               "h2d_bytes_per_step": int(st2.h2d_bytes), "d2h_bytes_per_step": int(st2.d2h_bytes),
               "note": "vs_scan_text: pinned host text -> H2D (chunked, overlapped with extract+score) -> hits D2H, wall clock"}
    clocks = sampler.stop(t0_wall, time.time()) if sampler else None

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (k_score) -----------------------------------------------------
    blocks = st.n_blocks_fwd + st.n_blocks_rev
    score_s = score_ms / args.steps * 1e-3
    n_score_launch = st.score_launches
    B_local = min(B - shard_first * 32, shard_words * 32)     # bases whose window starts this rank owns
    yard = C_ALG * ng * B_local / score_s                 # yardstick LOP3/s of the scoring launches of one step
    (lop_a, lds_a), (lop_b, lds_b) = score_ops(k)
    executed = lop_a * blocks * ng / score_s             # stage A only: a lower bound (stage B runs for the few warps that pass)
    lds = lds_a * blocks * ng / score_s
    # DRAM traffic of one full-chunk k_score launch from the committed ncu capture (dram__bytes_read + dram__bytes_write)
    traffic = None
    try:
        rd = wr = None
        for line in open(os.path.join(ROOT, "profiles", "r1_score_final_summary.txt")):
            f = line.split()
            if len(f) >= 3 and f[0] == "dram__bytes_read.sum":
                rd = float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[2]]
            if len(f) >= 3 and f[0] == "dram__bytes_write.sum":
                wr = float(f[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[2]]
        if rd is not None and wr is not None:
            traffic = rd + wr
    except Exception:
        pass
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    roof = {"bound": "int_alu", "kernel": "k_score", "achieved": yard / 1e12, "peak": peak_lop3 / 1e12, "unit": "Tlop3/s",
            "frac": yard / peak_lop3, "traffic": traffic,
            "traffic_note": "bytes per full-chunk launch (8 Mi words), ncu capture in profiles/r1_score_final_summary.txt; algorithmic = 192 B x blocks of the chunk",
            "hbm": {"achieved_gbs": (blocks * 192.0 * ((ng + 255) // 256)) / score_s / 1e9, "peak_gbs": hbm_peak,
                    "frac": ((blocks * 192.0 * ((ng + 255) // 256)) / score_s / 1e9 / hbm_peak) if hbm_peak else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if hbm_peak else "MEASURED_PEAKS.json absent"},
            "peak_source": "measured in this run by vs_measure_int_peaks (k_peak_lop3); MEASURED_PEAKS.json has no integer peak",
            "avg_launch_ms": score_ms / args.steps / max(1, n_score_launch), "launches_per_step": n_score_launch,
            "executed_lop3_tlops": executed / 1e12, "frac_executed": executed / peak_lop3,
            "ops_per_block_guide": {"stage_a_lop3": lop_a, "stage_a_lds": lds_a, "stage_b_lop3": lop_b, "stage_b_lds": lds_b},
            "lds_words_per_s_T": lds / 1e12, "lds_peak_T": peak_lds / 1e12, "frac_lds": lds / peak_lds,
            "hbm_gbs_algorithmic": (blocks * 192.0 * ((ng + 255) // 256)) / score_s / 1e9,
            "note": "yardstick = 4.0 LOP3 per guide*bp (dense scan, SURVEY.md 8d); PAM-first compaction scores ~1/8 of the windows per strand, so frac may exceed 1; frac_executed is the real alu-pipe load"}
    out = {
        "metric": "guide_Gbp_per_s", "value": value, "unit": "guide*Gbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u32 bit-sliced (LOP3)",
        "data": "synthetic",
        "config": {"workload": desc + (f" (scale {args.scale})" if args.scale != 1.0 else ""), "guides": ng, "k": k, "extra_pam": pam,
                   "text_bases_per_gpu": B, "contigs_per_gpu": text.n_contigs, "chunks": int(st.n_chunks),
                   "l2": "inputs larger than L2 (packed text %.2f GB resident, candidate planes %.2f GB written+read per step)" % ((nw * 16) / 1e9, blocks * 192 / 1e9),
                   "sharding": ("ONE text cut into %d word ranges by vs_shard_bounds, one per rank" % world) if strong else
                               "one text shard per rank, no collective; hits merged on the host", "cpus_bound_per_rank": numa_cpus},
        "wall_ms_per_step": wall_step, "phase_ms": {"extract": extract_ms / args.steps, "score": score_ms / args.steps},
        "hits_per_step": n_hits, "candidates": int(st.n_cand_fwd + st.n_cand_rev), "gpu_launches": launches,
        "roofline": roof, "e2e": e2e, "clocks": clocks, "gen_s": t_gen,
    }
    # ---- CPU baseline + hit-set diff on a bounded sample ------------------------------------------------
    if not args.no_cpu and world == 1:
        os.sched_setaffinity(0, all_cpus)             # the CPU baseline gets every host core back
        n, dt, rec, off, cores = cpu_sample(text, guides, k, pam)
        out["cpu_baseline"] = {"value": ng * n / dt / 1e9, "unit": "guide*Gbp/s", "cores": cores, "kind": "port",
                               "sample": f"first {n} bases of the same text, all {ng} guides, one pass ({dt:.1f} s); linear-scan oracle, not SeqAn"}
        # parity on the sample: windows starting before n-23 (the slice end is an artificial contig end)
        g_rec, _ = V.resolve_hits(hits, text.offsets)
        gpos = text.offsets[g_rec["contig"]].astype(np.int64) + g_rec["pos"].astype(np.int64)
        sel = gpos < n - 23
        gk = set(zip(g_rec["guide"][sel].tolist(), ((g_rec["flag"][sel] & 16) >> 4).tolist(), gpos[sel].tolist(), g_rec["mm"][sel].tolist()))
        opos = off[rec.contig].astype(np.int64) + rec.pos.astype(np.int64)
        osel = opos < n - 23
        ok = set(zip(rec.guide[osel].tolist(), ((rec.flag[osel] & 16) >> 4).tolist(), opos[osel].tolist(), rec.mm[osel].tolist()))
        out["parity"] = {"sample_bases": n, "hits_cpu": len(ok), "hits_gpu": len(gk), "diff": len(ok ^ gk)}
    print(json.dumps(out))
    ctx.close()
    text.unpin()
    del hits, hits_buf
    for p in pinned:
        _lib.lib().vs_host_free(p)


def _shutdown():
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    try:
        main()
    finally:
        _shutdown()
