#!/usr/bin/env python
"""bench.py -- one "step" = one pass of the hot path (vclust prefilter + vclust align) over one synthetic genome set.

Workload (config.workload): BASELINE.json configs[1] = "c2": 1 000 synthetic ~40 kb phage genomes (50 families x 20,
seed 20261019), prefilter k=25 --min-kmers 20 --min-ident 0.7 all-vs-all, then LZ-ANI alignment of every candidate
pair in both directions.  Metric: candidate genome pairs ANI-aligned per second (1 candidate pair = 2 directed parses).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm   (libvclust_b200.so through the C ABI)
  python bench.py --impl reference [...]                         the reference's own CPU binaries (oracle/_ref)

N > 1 (torchrun, one rank per GPU): weak scaling -- every rank runs the step on its own genome set (seed + rank), no
data-path collective; value = total pairs of all ranks / max-over-ranks time.
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
WORKLOAD = "c2: 1000 synthetic 40 kb phage genomes (50 families x 20), prefilter k=25 min-kmers 20 min-ident 0.7 + LZ-ANI align of all candidate pairs"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


CONFIG_NAME = "c2"     # --workload c3 / c3_s200: the other single-GPU BASELINE configurations (not the headline line)


def make_set(rank: int):
    from vclust_b200 import synth
    cfg = dict(synth.CONFIGS[CONFIG_NAME])
    cfg["seed"] += 1000 * rank
    return synth.make_genomes(**cfg)


def alg_bytes_align(lens, ref_ids, qry_ids):
    lens = np.asarray(lens, dtype=np.int64)
    return int((lens[qry_ids] // 4 + 2 * (lens[ref_ids] // 4) + 12).sum())


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own binaries on the host cores
# ----------------------------------------------------------------------------------------------------------------
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
            "prefilter_s": t_pre, "align_s": t_al}


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` summary (profiles/r01_kernels_ncu_full.txt,
    written by profiles/ncu_summary.py from the capture of this same bench command); None when absent."""
    f = ROOT / "profiles" / "r01_kernels_ncu_full.txt"
    if not f.exists():
        return None
    total, inside = 0.0, False
    for ln in f.read_text().splitlines():
        if ln.startswith("## "):
            if inside:
                break
            inside = ln[3:].startswith(kernel)
        elif inside and ("dram__bytes_read.sum" in ln or "dram__bytes_write.sum" in ln):
            parts = ln.split()
            total += float(parts[1]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(parts[2], 1)
    return total if inside and total else None


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


def main_reference_port(args):
    """--impl reference without the reference binaries: the oracle port, 1 thread, 100 genomes (5 whole families) per step."""
    names, seqs = make_set(0)
    n_sample = 100
    for _ in range(args.warmup):
        port_once(names, seqs, n_sample)
    times, pairs = [], 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        pairs, _, _ = port_once(names, seqs, n_sample)
        times.append(time.perf_counter() - t0)
    ms = 1000 * sum(times) / len(times)
    value = pairs / (ms / 1000)
    sample = "oracle port (plain C, 1 thread) on the first %d of %d genomes: %d candidate pairs per step" % (n_sample, len(names), pairs)
    print(json.dumps({
        "impl": "reference", "metric": "genome_pairs_ani_per_sec", "value": value, "unit": "candidate pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (2-bit bases, integer counts; f64 only for the final ratios)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "candidate pairs/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "candidate pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from oracle import oracle
    from vclust_b200 import synth
    if not oracle.ref_available():
        main_reference_port(args)
        return
    threads = os.cpu_count() or 1
    names, seqs = make_set(0)
    n_sample = len(names)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        fa = td / "c2.fna"
        synth.write_fasta(fa, names, seqs)
        # bounded run: if the first (warm-up) pass predicts more than ~150 s for all passes, shrink the sample
        t0 = time.perf_counter()
        n_pairs, t_pre, t_al, _ = run_reference_once(fa, td, threads)
        first = time.perf_counter() - t0
        total_passes = args.steps + args.warmup
        if first * total_passes > 150 and total_passes > 1:
            frac = max(0.05, 150.0 / (first * total_passes))
            n_sample = max(40, int(len(names) * frac) // 20 * 20)
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
    value = pairs / (ms / 1000)
    sample = "first %d of 1000 genomes: %d candidate pairs per step, -t %d" % (n_sample, pairs, threads)
    print(json.dumps({
        "impl": "reference", "metric": "genome_pairs_ani_per_sec", "value": value, "unit": "candidate pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (2-bit bases, integer counts; f64 only for the final ratios)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "candidate pairs/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "candidate pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def main_ours(args, rank: int, world: int, local_rank: int):
    import torch
    from vclust_b200 import api

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)

    names, seqs = make_set(rank)
    lens = [int(s.size) for s in seqs]
    ctx = api.Context(local_rank)
    g = api.Genomes.from_memory(names, seqs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step(detail=True):
        t0 = time.perf_counter()
        pairs = api.prefilter_genomes(ctx, g, **PRE)
        t1 = time.perf_counter()
        res = api.align_genomes(ctx, g, pairs)
        t2 = time.perf_counter()
        n_pairs, n_dir = pairs.n_pairs, res.n
        info = dict(pairs=n_pairs, directed=n_dir, wall_prefilter_ms=1000 * (t1 - t0), wall_align_ms=1000 * (t2 - t1))
        if detail:      # read-back of timers and pair lists for the report: outside the timed region
            info.update(pre=ctx.timings("prefilter"), aln=ctx.timings("align"), ref=res.ref, qry=res.qry, order=res.order)
        info["_objs"] = (pairs, res)
        return info

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

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
            info = step(detail=False)
            ctx.mark(1)
            wall_ms.append(1000 * (time.perf_counter() - t0))
            ev_ms.append(ctx.elapsed_ms(0, 1))
            pairs, res = info.pop("_objs")
            info.update(pre=ctx.timings("prefilter"), aln=ctx.timings("align"), ref=res.ref, qry=res.qry, order=res.order)
            pairs.close(); res.close()
            infos.append(info)
        return ev_ms, wall_ms, infos

    # ---- device-resident measurement ("value")
    ctx.make_resident(g, api.FASTA_KMERDB)
    ctx.make_resident(g, api.FASTA_LZANI, 40)
    def warm():
        i = step(detail=False)
        for o in i.pop("_objs"):
            o.close()

    sampler = ClockSampler(local_rank)       # runs from the warm-up to the end of the e2e leg (the timed region alone
    sampler.start()                          # lasts tens of milliseconds: too short for 200 ms samples)
    for _ in range(args.warmup):
        warm()
    barrier()
    l0 = ctx.launches
    t_region0 = time.perf_counter()
    ev_ms, wall_ms, infos = timed(args.steps)
    barrier()
    region_s = time.perf_counter() - t_region0
    launches = ctx.launches - l0
    # ---- end-to-end measurement: host buffers, H2D inside every step
    ctx.evict()
    if args.quick:
        e_wall = wall_ms
    else:
        warm()
        barrier()
        e_ev, e_wall, e_infos = timed(args.steps, resident=False)
        barrier()
    clocks = sampler.stop()

    step_ms = float(np.mean(wall_ms))          # wall of the synchronous calls == device events + host glue
    e2e_ms = float(np.mean(e_wall))
    info = infos[-1]
    pairs_step = info["pairs"]
    t = torch.tensor([step_ms, e2e_ms, float(pairs_step)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        step_ms, e2e_ms, pairs_total = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        pairs_total = float(pairs_step)

    if rank == 0:
        peak, peak_src = peaks()
        order = info["order"].astype(np.int64)
        lens_lz = np.asarray(lens, dtype=np.int64)[order]
        a_bytes = alg_bytes_align(lens_lz, info["ref"].astype(np.int64), info["qry"].astype(np.int64))
        parse_ms = float(np.mean([i["aln"]["parse_ms"] for i in infos]))
        pre = {k: float(np.mean([i["pre"][k] for i in infos])) for k in infos[0]["pre"]}
        aln = {k: float(np.mean([i["aln"][k] for i in infos])) for k in infos[0]["aln"]}
        total_bases = int(sum(lens))
        pre_bytes = 24.25 * total_bases * PRE["kmers_fraction"] + 12 * pairs_step + 4 * len(lens)
        pre_kernel_ms = pre["extract_ms"] + pre["sort_ms"] + pre["segment_ms"] + pre["emit_ms"]
        # one upload per step serves both stages: ASCII bases + record offsets + store offsets/lengths
        h2d = total_bases + 8 * (len(lens) + 1) + 12 * len(lens)
        d2h = 12 * pairs_step + 4 * len(lens) + 12 * info["directed"]
        out = {
            "metric": "genome_pairs_ani_per_sec", "value": pairs_total / (step_ms / 1000), "unit": "candidate pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32 (2-bit bases, integer counts; f64 only for the final ratios)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "genomes_per_gpu": len(lens), "bases_per_gpu": total_bases,
                       "candidate_pairs_per_gpu": pairs_step, "directed_parses_per_gpu": info["directed"],
                       "l2": "256 MiB written between steps (outside the per-step brackets)",
                       "timing": "per step: wall clock of the synchronous C-ABI calls, cross-checked by CUDA events on the library stream (ms_per_step_events)"},
            "ms_per_step_events": float(np.mean(ev_ms)),
            "directed_parses_per_sec": 2 * pairs_total / (step_ms / 1000),
            "e2e": {"value": pairs_total / (e2e_ms / 1000), "unit": "candidate pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stages_ms": {"prefilter": pre, "align": aln,
                          "wall_prefilter_call": float(np.mean([i["wall_prefilter_ms"] for i in infos])),
                          "wall_align_call": float(np.mean([i["wall_align_ms"] for i in infos]))},
            "roofline": {"kernel": "parse_kernel (align)", "bound": "hbm", "achieved": a_bytes / (parse_ms / 1000) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": a_bytes / (parse_ms / 1000) / 1e9 / peak, "traffic": ncu_traffic("parse_kernel"),
                         "algorithmic_bytes": a_bytes,
                         "peak_source": peak_src,
                         "note": "algorithmic bytes = sum over directed pairs of Lq/4 + 2*Lr/4 + 12; the parse is issue/latency bound, not HBM bound (DESIGN.md)"},
            "roofline_prefilter": {"kernels": "screen+collect+partition+bucket+emit (whole prefilter device time)", "bound": "hbm",
                                   "achieved": pre_bytes / (pre_kernel_ms / 1000) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": pre_bytes / (pre_kernel_ms / 1000) / 1e9 / peak,
                                   "note": "algorithmic bytes = 24.25 B/base + 12 B/pair + 4 B/genome (SURVEY 8(d))"},
            "timed_region_s": region_s,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(names, seqs, len(names), os.cpu_count() or 1)
            if not cb:                                   # the reference binaries did not travel: the oracle port, 1 thread
                n_pairs, t_pre, t_al = port_once(names, seqs, 100)
                cb = {"value": n_pairs / (t_pre + t_al), "unit": "candidate pairs/s", "cores": 1, "kind": "port",
                      "sample": "oracle port (plain C, 1 thread), first 100 of %d genomes: %d candidate pairs; prefilter %.2f s, "
                                "parse %.2f s" % (len(names), n_pairs, t_pre, t_al), "prefilter_s": t_pre, "align_s": t_al}
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    ctx.evict()
    g.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c3_s200"],
                    help="c2 = the configuration the metric is quoted on (default); c3: 10 000 x 40 kb genomes")
    ap.add_argument("--quick", action="store_true", help="profiling run: 1 warm-up, no e2e leg, no CPU baseline (never a bench value)")
    args = ap.parse_args()
    if args.quick:
        args.warmup, args.no_cpu_baseline = 1, True
    elif args.impl == "ours":
        args.warmup = max(args.warmup, 3)
    global CONFIG_NAME, WORKLOAD
    if args.workload != "c2":
        from vclust_b200 import synth
        CONFIG_NAME = args.workload
        c = synth.CONFIGS[CONFIG_NAME]
        WORKLOAD = "%s: %d synthetic %d kb genomes (families of %d), prefilter k=25 min-kmers 20 min-ident 0.7 + LZ-ANI align of all candidate pairs" % (
            CONFIG_NAME, c["n"], c["length"] // 1000, c["family"])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
    else:
        main_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
