#!/usr/bin/env python
"""bench.py -- one "step" = one pass of the hot path (vclust prefilter + vclust align) over ONE synthetic genome set.

Workload (config.workload): BASELINE.json configs[2] = "c3", the largest single-GPU configuration: 10 000 synthetic
~40 kb phage genomes (500 families x 20), prefilter k=25 --min-kmers 20 --min-ident 0.7 all-vs-all, then LZ-ANI alignment
of every candidate pair in both directions.  Metric: candidate genome pairs ANI-aligned per second (1 candidate pair =
2 directed parses).  `--workload c2` selects configs[1] (1 000 genomes).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm   (libvclust_b200.so through the C ABI)
  python bench.py --impl reference [...]                         the reference's own CPU binaries (oracle/_ref)

N > 1 (torchrun, one rank per GPU): STRONG scaling -- the same genome set, block-partitioned over the ranks and run
through the sharded pipeline (vclust_b200/distributed.py: tuple all-to-all, owner merge, owner-local parses, gather on
rank 0); value = candidate pairs of the set / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

PRE = dict(k=25, min_kmers=20, min_ident=0.7, kmers_fraction=1.0)
# the prefilter's kernels and their launches per step (one GPU, dense pair counters): for roofline.traffic
PREFILTER_KERNELS = {"collect_kernel": 1, "part_kernel<1>": 1, "part_kernel<2>": 1, "bucket_chain_kernel": 1, "dense_emit_kernel": 2}
CONFIG_NAME = "c3"
DTYPE = "u8/int32 (2-bit bases, integer counts; f64 only for the final ratios)"


def workload_text():
    from vclust_b200 import synth
    c = synth.CONFIGS[CONFIG_NAME]
    return "%s: %d synthetic %d kb phage genomes (%d families x %d), prefilter k=25 min-kmers 20 min-ident 0.7 + LZ-ANI align of all candidate pairs" % (
        CONFIG_NAME, c["n"], c["length"] // 1000, c["n"] // c["family"], c["family"])


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled every 50 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_set():
    from vclust_b200 import synth
    return synth.make_genomes(**synth.CONFIGS[CONFIG_NAME])


def alg_bytes_align(lens, ref_ids, qry_ids):
    lens = np.asarray(lens, dtype=np.int64)
    return int((lens[qry_ids] // 4 + 2 * (lens[ref_ids] // 4) + 12).sum())


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own binaries on the host cores
# ----------------------------------------------------------------------------------------------------------------
REF_NOTE = ("unmodified kmer-db 2.3.1 + lz-ani 1.2.3 sources built by oracle/build_ref.sh with plain g++ -O3 -march=x86-64-v3, system zlib "
            "and no mimalloc (the reference's own makefile would use -march=native, zlib-ng/isa-l and mimalloc); FASTA file in, "
            "filter + ani.tsv files out, 4 processes")


def run_reference_once(fa: Path, wd: Path, threads: int):
    from oracle import oracle
    t = {}
    flt = wd / "fltr.txt"
    t0 = time.perf_counter()
    oracle.ref_prefilter([fa], flt, wd / "p", k=PRE["k"], fraction=PRE["kmers_fraction"], min_kmers=PRE["min_kmers"],
                         min_ident=PRE["min_ident"], threads=threads, timings=t)
    t_pre = time.perf_counter() - t0
    n_pairs = sum(ln.count(":") for ln in flt.read_text().splitlines()[1:])
    ta = {}
    t0 = time.perf_counter()
    oracle.ref_align([fa], wd / "ani.tsv", wd / "a", filter_path=flt, threads=threads, timings=ta)
    t_al = time.perf_counter() - t0
    return n_pairs, t_pre, t_al, ta.get("lz_matching")


def cpu_baseline(names, seqs, n_sample: int, threads: int):
    """Reference binaries on a bounded sample (the first n_sample genomes = whole families) of the workload."""
    from oracle import oracle
    from vclust_b200 import synth
    if not oracle.ref_available():
        return None
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        fa = td / "sample.fna"
        synth.write_fasta(fa, names[:n_sample], seqs[:n_sample])
        n_pairs, t_pre, t_al, lz = run_reference_once(fa, td, threads)
    return {"value": n_pairs / (t_pre + t_al), "unit": "candidate pairs/s", "cores": threads, "kind": "reference",
            "sample": "first %d of %d genomes (whole families): %d candidate pairs; kmer-db build+all2all-sp+distance %.2f s, "
                      "lz-ani %.2f s (LZ matching %.2f s), -t %d" % (n_sample, len(names), n_pairs, t_pre, t_al, lz or -1, threads),
            "prefilter_s": t_pre, "align_s": t_al, "build": REF_NOTE}


def ncu_value(kernel, metrics):
    """Sum of `metrics` for `kernel` (a name prefix, or a dict {prefix: launches per step}) from the committed
    `ncu --set full` summary of this same bench command (profiles/r02_kernels_ncu_full.txt, written by
    profiles/ncu_summary.py; per-launch values); None when absent."""
    if isinstance(kernel, dict):
        vals = [(ncu_value(k, metrics), n) for k, n in kernel.items()]
        if any(v is None for v, _ in vals):
            return None
        return float(sum(v * n for v, n in vals))
    f = ROOT / "profiles" / ("r02_kernels_ncu_full_%s.txt" % CONFIG_NAME)
    if not f.exists():
        return None
    total, inside, seen = 0.0, False, False
    for ln in f.read_text().splitlines():
        if ln.startswith("## "):
            if inside:
                break
            inside = ln[3:].startswith(kernel)
        elif inside and any(m in ln for m in metrics):
            parts = ln.split()
            try:
                total += float(parts[1]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(parts[2] if len(parts) > 2 else "", 1)
                seen = True
            except (ValueError, IndexError):
                pass
    return total if seen else None


def port_once(names, seqs, n_sample: int):
    """The plain-C oracle port (one thread) on the first n_sample genomes: used only when oracle/_ref did not travel."""
    from oracle import oracle
    raw = [s.tobytes() for s in seqs[:n_sample]]
    t0 = time.perf_counter()
    sets = oracle.kmer_sets([[s] for s in raw], PRE["k"], PRE["kmers_fraction"])
    pairs = oracle.prefilter_pairs(sets, PRE["k"], PRE["min_kmers"], PRE["min_ident"])
    t_pre = time.perf_counter() - t0
    ref = [r for r, c, *_ in pairs] + [c for r, c, *_ in pairs]
    qry = [c for r, c, *_ in pairs] + [r for r, c, *_ in pairs]
    t0 = time.perf_counter()
    oracle.run_pairs([oracle.lz_codes(s) for s in raw], ref, qry)
    t_al = time.perf_counter() - t0
    return len(pairs), t_pre, t_al


def reference_line(args, value, ms, sample, cores, kind):
    return {
        "impl": "reference", "metric": "genome_pairs_ani_per_sec", "value": value, "unit": "candidate pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic", "config": {"workload": workload_text(), "sample": sample, "build": REF_NOTE},
        "cpu_baseline": {"value": value, "unit": "candidate pairs/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "candidate pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from oracle import oracle
    from vclust_b200 import synth
    names, seqs = make_set()
    if not oracle.ref_available():
        # the reference binaries did not travel: the oracle port, 1 thread, 100 genomes (5 whole families) per step
        n_sample = 100
        for _ in range(args.warmup):
            port_once(names, seqs, n_sample)
        times, pairs = [], 0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            pairs, _, _ = port_once(names, seqs, n_sample)
            times.append(time.perf_counter() - t0)
        ms = 1000 * sum(times) / len(times)
        sample = "oracle port (plain C, 1 thread) on the first %d of %d genomes: %d candidate pairs per step" % (n_sample, len(names), pairs)
        emit(json.dumps(reference_line(args, pairs / (ms / 1000), ms, sample, 1, "port")))
        return
    threads = os.cpu_count() or 1
    total_passes = args.steps + args.warmup
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        fa = td / "sample.fna"
        # bounded run: a first pass on 1 000 genomes (whole families) sizes the sample so that all passes take ~2 minutes
        n_sample = min(1000, len(names))
        synth.write_fasta(fa, names[:n_sample], seqs[:n_sample])
        t0 = time.perf_counter()
        run_reference_once(fa, td, threads)
        first = time.perf_counter() - t0
        per_genome = first / n_sample
        want = int(120.0 / max(total_passes, 1) / per_genome)
        n_sample = max(40, min(len(names), want) // 20 * 20)
        synth.write_fasta(fa, names[:n_sample], seqs[:n_sample])
        for _ in range(max(args.warmup - 1, 0)):
            run_reference_once(fa, td, threads)
        times, pairs = [], 0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            n_pairs, t_pre, t_al, _ = run_reference_once(fa, td, threads)
            times.append(time.perf_counter() - t0)
            pairs = n_pairs
    ms = 1000 * sum(times) / len(times)
    sample = "first %d of %d genomes (whole families): %d candidate pairs per step, -t %d, file to file" % (n_sample, len(names), pairs, threads)
    emit(json.dumps(reference_line(args, pairs / (ms / 1000), ms, sample, threads, "reference")))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def files_leg(names, seqs):
    """FASTA file -> filter file -> ani.tsv + ids.tsv through the file-level API, each stage with a fresh context (what
    `vclust prefilter` + `vclust align` do): the like-for-like counterpart of the reference arm."""
    from vclust_b200 import api, synth
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        fa = td / "in.fna"
        synth.write_fasta(fa, names, seqs)
        best = None
        for _ in range(2):                       # second pass: page cache and CUDA module warm
            t0 = time.perf_counter()
            ip = api.prefilter([fa], td / "fltr.txt", True, kmer_size=PRE["k"], min_kmers=PRE["min_kmers"], min_ident=PRE["min_ident"])
            t1 = time.perf_counter()
            ia = api.align([fa], td / "ani.tsv", True, filter_file=td / "fltr.txt")
            t2 = time.perf_counter()
            best = (t1 - t0, t2 - t1, ip["wall_s"], ia["wall_s"])
        pairs = sum(ln.count(":") for ln in (td / "fltr.txt").read_text().splitlines()[1:])
    return {"value": pairs / (best[0] + best[1]), "unit": "candidate pairs/s", "prefilter_s": best[0], "align_s": best[1],
            "prefilter_phases_s": best[2], "align_phases_s": best[3],
            "note": "FASTA -> filter -> ani.tsv/ids.tsv, context creation, FASTA parsing, upload and text output included (second of two passes)"}


def stage_means(infos, key):
    return {k: float(np.mean([i[key][k] for i in infos])) for k in infos[0][key]}


def build_line(args, world, step_ms, e2e_ms, pairs_total, directed, lens, infos, launches, clocks, region_s, h2d, d2h, extra):
    peak, sm_mhz, peak_src = peaks()
    pre, aln = stage_means(infos, "pre"), stage_means(infos, "aln")
    total_bases = int(sum(lens))
    # algorithmic bytes of the prefilter = 24.25 B/base + 12 B/pair + 4 B/genome (whole set; per GPU: / N)
    pre_bytes = 24.25 * total_bases * PRE["kmers_fraction"] + 12 * pairs_total + 4 * len(lens)
    pre_kernel_ms = pre["extract_ms"] + pre["sort_ms"] + pre["segment_ms"] + pre["emit_ms"] + pre.get("exchange_ms", 0.0)
    pre_rate = pre_bytes / world / (pre_kernel_ms / 1000) / 1e9
    out = {
        "metric": "genome_pairs_ani_per_sec", "value": pairs_total / (step_ms / 1000), "unit": "candidate pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": dict({"workload": workload_text(), "genomes": len(lens), "bases": total_bases,
                        "candidate_pairs": int(pairs_total), "directed_parses": int(directed),
                        "l2": "256 MiB written between steps (outside the per-step brackets)",
                        "timing": "per step: CUDA events on the stream every kernel and collective of the step runs on, max over ranks; "
                                  "cross-checked by the wall clock of the synchronous calls (ms_per_step_wall)"}, **extra.pop("config", {})),
        "directed_parses_per_sec": 2 * pairs_total / (step_ms / 1000),
        "e2e": {"value": pairs_total / (e2e_ms / 1000), "unit": "candidate pairs/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "stages_ms": {"prefilter": pre, "align": aln},
        # the path's HBM-bound stage and the north-star metric: the whole prefilter (its kernels + collectives), per GPU
        "roofline": {"kernel": "prefilter: collect + partition x2 + bucket grouping + emit (all kernels of the stage)", "bound": "hbm",
                     "achieved": pre_rate, "peak": peak, "unit": "GB/s", "frac": pre_rate / peak,
                     "traffic": (lambda t: None if t is None else t / world)(ncu_value(dict(PREFILTER_KERNELS, **({"screen_kernel": 1} if CONFIG_NAME == "c2" else {})), ["dram__bytes_read.sum", "dram__bytes_write.sum"])),
                     "algorithmic_bytes": pre_bytes / world, "stage_ms": pre_kernel_ms, "peak_source": peak_src,
                     "note": "algorithmic bytes = 24.25 B/base + 12 B/pair + 4 B/genome (SURVEY 8(d)), per GPU; duration = CUDA events around the stage's kernels"},
        "timed_region_s": region_s,
    }
    # the parse is bound by the SM issue rate, not by HBM: warp instructions (deterministic for a workload; from the
    # committed ncu capture of this command) / live duration / (SMs x 4 schedulers x clock)
    inst = ncu_value("parse_sched_kernel", ["smsp__inst_executed.sum"])
    if inst and aln.get("parse_ms"):
        slots = 148 * 4 * sm_mhz * 1e6 * (aln["parse_ms"] / 1000)
        out["roofline_issue"] = {"kernel": "parse_sched_kernel (align)", "bound": "issue", "achieved": inst / world / (aln["parse_ms"] / 1000) / 1e9,
                                 "peak": 148 * 4 * sm_mhz * 1e6 / 1e9, "unit": "G warp-instructions/s", "frac": inst / world / slots,
                                 "note": "smsp__inst_executed.sum of the committed ncu capture (same workload) / live CUDA-event duration"}
    out.update(extra)
    return out


def main_single(args, local_rank: int):
    import torch
    from vclust_b200 import api

    torch.cuda.set_device(local_rank)
    names, seqs = make_set()
    lens = [int(s.size) for s in seqs]
    ctx = api.Context(local_rank)
    g = api.Genomes.from_memory(names, seqs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        t0 = time.perf_counter()
        pairs = api.prefilter_genomes(ctx, g, **PRE)
        t1 = time.perf_counter()
        res = api.align_genomes(ctx, g, pairs)
        t2 = time.perf_counter()
        return pairs, res, dict(pairs=pairs.n_pairs, directed=res.n, wall_prefilter_ms=1000 * (t1 - t0), wall_align_ms=1000 * (t2 - t1))

    def timed(n_steps: int, resident: bool = True):
        """K steps; every step bracketed by CUDA events on the library's stream; L2 flushed between steps."""
        ev_ms, wall_ms, infos = [], [], []
        for _ in range(n_steps):
            flush.zero_()
            torch.cuda.synchronize()
            if not resident:
                ctx.evict()
            ctx.mark(0)
            t0 = time.perf_counter()
            pairs, res, info = step()
            ctx.mark(1)
            wall_ms.append(1000 * (time.perf_counter() - t0))
            ev_ms.append(ctx.elapsed_ms(0, 1))
            info.update(pre=ctx.timings("prefilter"), aln=ctx.timings("align"))
            if not infos:
                info.update(ref=res.ref, qry=res.qry, order=res.order)
            pairs.close(); res.close()
            infos.append(info)
        return ev_ms, wall_ms, infos

    def warm():
        pairs, res, _ = step()
        pairs.close(); res.close()

    # ---- device-resident measurement ("value")
    ctx.make_resident(g, api.FASTA_KMERDB)
    ctx.make_resident(g, api.FASTA_LZANI, 40)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        warm()
    torch.cuda.synchronize()
    l0 = ctx.launches
    t_region0 = time.perf_counter()
    ev_ms, wall_ms, infos = timed(args.steps)
    torch.cuda.synchronize()
    region_s = time.perf_counter() - t_region0
    launches = ctx.launches - l0
    # ---- end-to-end measurement: host buffers, H2D inside every step
    ctx.evict()
    if args.quick:
        e_wall = wall_ms
    else:
        warm()
        e_ev, e_wall, e_infos = timed(args.steps, resident=False)
    clocks = sampler.stop()

    step_ms = float(np.mean(wall_ms))          # wall of the synchronous calls == device events + host glue
    e2e_ms = float(np.mean(e_wall))
    info = infos[0]
    pairs_step = info["pairs"]
    total_bases = int(sum(lens))
    # one upload per step serves both stages: ASCII bases + record offsets + store offsets/lengths
    h2d = total_bases + 8 * (len(lens) + 1) + 12 * len(lens)
    d2h = 12 * pairs_step + 4 * len(lens) + 20 * info["directed"]
    order = info["order"].astype(np.int64)
    lens_lz = np.asarray(lens, dtype=np.int64)[order]
    a_bytes = alg_bytes_align(lens_lz, info["ref"].astype(np.int64), info["qry"].astype(np.int64))
    parse_ms = float(np.mean([i["aln"]["parse_ms"] for i in infos]))
    extra = {"ms_per_step_events": float(np.mean(ev_ms)), "ms_per_step_wall": step_ms,
             "align_hbm": {"kernel": "parse_sched_kernel", "algorithmic_bytes": a_bytes, "achieved_GBs": a_bytes / (parse_ms / 1000) / 1e9,
                           "note": "sum over directed pairs of Lq/4 + 2*Lr/4 + 12 bytes; the parse is issue/latency bound (roofline_issue), this is for reference only",
                           "traffic": ncu_value("parse_sched_kernel", ["dram__bytes_read.sum", "dram__bytes_write.sum"])},
             "wall_ms": {"prefilter_call": float(np.mean([i["wall_prefilter_ms"] for i in infos])),
                         "align_call": float(np.mean([i["wall_align_ms"] for i in infos]))}}
    out = build_line(args, 1, step_ms, e2e_ms, float(pairs_step), info["directed"], lens, infos, launches, clocks, region_s, h2d, d2h, extra)
    if not args.quick and not args.no_files:
        ctx.evict()
        out["e2e_files"] = files_leg(names, seqs)
    if not args.no_cpu_baseline:
        n_sample = min(len(names), 2000)
        cb = cpu_baseline(names, seqs, n_sample, os.cpu_count() or 1)
        if not cb:                                   # the reference binaries did not travel: the oracle port, 1 thread
            n_pairs, t_pre, t_al = port_once(names, seqs, 100)
            cb = {"value": n_pairs / (t_pre + t_al), "unit": "candidate pairs/s", "cores": 1, "kind": "port",
                  "sample": "oracle port (plain C, 1 thread), first 100 of %d genomes: %d candidate pairs; prefilter %.2f s, "
                            "parse %.2f s" % (len(names), n_pairs, t_pre, t_al), "prefilter_s": t_pre, "align_s": t_al}
        out["cpu_baseline"] = cb
    emit(json.dumps(out))
    ctx.evict()
    g.close()
    ctx.close()


def main_sharded(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    from vclust_b200 import distributed

    torch.cuda.set_device(local_rank)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    names, seqs = make_set()                     # every rank generates the same set and keeps the sequences of its block
    lens = [int(s.size) for s in seqs]
    first, count = distributed.block_partition(lens, world)[rank]
    run = distributed.ShardedRun(dist, local_rank, names, lens, seqs[first:first + count])
    del seqs
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ctx = run.ctx

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def step():
        pairs = run.prefilter(**PRE)
        pre = ctx.timings("prefilter")
        res = run.align()
        return pairs, res, dict(pairs=pairs.n_pairs, directed=res.n, pre=pre, aln=ctx.timings("align"))

    def timed(n_steps: int, reload: bool):
        ev_ms, wall_ms, infos = [], [], []
        for _ in range(n_steps):
            flush.zero_()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(run.stream)
            t0 = time.perf_counter()
            if reload:
                run.load()
            pairs, res, info = step()
            e1.record(run.stream)
            run.stream.synchronize()
            wall_ms.append(1000 * (time.perf_counter() - t0))
            ev_ms.append(e0.elapsed_time(e1))
            pairs.close(); res.close()
            infos.append(info)
        return ev_ms, wall_ms, infos

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        p_, r_, _ = step()
        p_.close(); r_.close()
    barrier()
    l0, b0, c0 = ctx.launches, run.comm.bytes_sent, run.comm.calls
    t_region0 = time.perf_counter()
    ev_ms, wall_ms, infos = timed(args.steps, reload=False)
    barrier()
    region_s = time.perf_counter() - t_region0
    launches = ctx.launches - l0
    sent = (run.comm.bytes_sent - b0) / args.steps
    calls = (run.comm.calls - c0) / args.steps
    if args.quick:
        e_wall = wall_ms
    else:
        run.load()
        p_, r_, _ = step()
        p_.close(); r_.close()
        barrier()
        _, e_wall, _ = timed(args.steps, reload=True)
        barrier()
    clocks = sampler.stop()

    # max over ranks, per step (every rank's step i is bracketed by the same barriers)
    t = torch.tensor([ev_ms, wall_ms, e_wall], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    peer = float(np.mean([i["pre"].get("peer_bytes", 0.0) for i in infos]))
    agg = torch.tensor([float(launches), float(sent), peer], dtype=torch.float64, device="cuda")
    dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    # per-rank stage times (for the report: slowest rank per stage)
    keys_p = ["extract_ms", "sort_ms", "segment_ms", "exchange_ms", "emit_ms", "total_ms"]
    keys_a = ["list_ms", "index_ms", "parse_ms", "gather_ms", "total_ms", "pairs"]
    st = torch.tensor([[float(np.mean([i["pre"].get(k, 0.0) for i in infos])) for k in keys_p] +
                       [float(np.mean([i["aln"].get(k, 0.0) for i in infos])) for k in keys_a]], dtype=torch.float64, device="cuda")
    st_max = st.clone(); dist.all_reduce(st_max, op=dist.ReduceOp.MAX)
    st_min = st.clone(); dist.all_reduce(st_min, op=dist.ReduceOp.MIN)
    if rank == 0:
        step_ms = float(t[0].mean())               # device time of the step, slowest rank
        e2e_ms = float(t[2].mean())
        info = infos[-1]
        for i in infos:                            # the stage report of build_line: slowest rank per stage
            i["pre"] = dict(i["pre"]); i["aln"] = dict(i["aln"])
        mx, mn = st_max[0].tolist(), st_min[0].tolist()
        for j, k in enumerate(keys_p):
            for i in infos:
                i["pre"][k] = mx[j]
        for j, k in enumerate(keys_a):
            for i in infos:
                i["aln"][k] = mx[len(keys_p) + j]
        total_bases = int(sum(lens))
        h2d = total_bases + 8 * (len(lens) + 1) * world + 12 * len(lens) * world
        d2h = 12 * info["pairs"] + 4 * len(lens) + 20 * info["directed"]
        extra = {"ms_per_step_events": step_ms, "ms_per_step_wall": float(t[1].mean()),
                 "config": {"parallelism": "genomes block-partitioned over %d GPUs; k-mer tuples all-to-all by hash range; partial pair counts "
                                           "all-to-all to the genome owners; owner-local parses; results gathered on rank 0" % world,
                            "e2e_note": "every step: each rank uploads and packs its block (ASCII from host memory), the packed records are all-gathered, then the step"},
                 "comm": {"backend": "nccl + CUDA-IPC peer copies (tuple exchange)" if float(agg[2]) > 0 else "nccl", "ranks": world,
                          "collectives_per_step_per_rank": calls, "nccl_bytes_sent_per_step_all_ranks": float(agg[1]),
                          "nvlink_peer_copy_bytes_per_step_all_ranks": float(agg[2]), "exchange_ms_prefilter_max_rank": mx[keys_p.index("exchange_ms")],
                          "gather_ms_align_max_rank": mx[len(keys_p) + keys_a.index("gather_ms")]},
                 "balance": {"parse_ms_min_rank": mn[len(keys_p) + keys_a.index("parse_ms")],
                             "parse_ms_max_rank": mx[len(keys_p) + keys_a.index("parse_ms")],
                             "directed_pairs_min_rank": mn[len(keys_p) + keys_a.index("pairs")],
                             "directed_pairs_max_rank": mx[len(keys_p) + keys_a.index("pairs")]}}
        out = build_line(args, world, step_ms, e2e_ms, float(info["pairs"]), info["directed"], lens, infos, float(agg[0]), clocks, region_s,
                         h2d, d2h, extra)
        emit(json.dumps(out))
    run.close()
    dist.destroy_process_group()


_RESULT_FD = None


def emit(line: str) -> None:
    """The one JSON line of the contract, written to the process' ORIGINAL stdout."""
    data = (line + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line + "\n")
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # stdout carries the JSON line and nothing else: whatever libraries print there (NCCL's version banner, under
    # NCCL_DEBUG=VERSION, goes to stdout) is sent to stderr instead, and the line itself to a duplicate of the original fd
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip the file-to-file leg (e2e_files)")
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c3_s200"],
                    help="c3 = the configuration the metric is quoted on (default: the largest single-GPU one); c2: 1 000 genomes")
    ap.add_argument("--quick", action="store_true", help="profiling run: 1 warm-up, no e2e leg, no CPU baseline (never a bench value)")
    args = ap.parse_args()
    if args.quick:
        args.warmup, args.no_cpu_baseline = 1, True
    elif args.impl == "ours":
        args.warmup = max(args.warmup, 3)
    global CONFIG_NAME
    CONFIG_NAME = args.workload
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
    elif world > 1:
        main_sharded(args, rank, world, local_rank)
    else:
        main_single(args, local_rank)


if __name__ == "__main__":
    main()
