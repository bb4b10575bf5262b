#!/usr/bin/env python
"""bench.py -- headline benchmark of the khoice exp-type-1 k-mer path on B200.

Metric (BASELINE.json): Gbases/s to final k-mer occurrence tables (k=31).  One "step" = the whole job on
the workload below: every group's step_4 histogram plus the step_8 across-group histogram.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N=1 workload: BASELINE config 2 -- 10 groups x 50 synthetic 5 Mbp genomes (2.5 Gbp), k=31, single B200.
N>1 (torchrun, one rank per GPU): weak scaling -- every rank owns 10 such groups; steps 1-6 need no
collective; for the across-group stage every group's distinct k-mers are pushed to their hash-range owner over peer memory
(csrc/peer.cu) -- or one NCCL all-to-all (KHB_EXCHANGE=nccl) -- then a local count and a histogram all-reduce (khoice_b200/dist.py).

value   Gbases/s with the FASTA text already staged in HBM when the clock starts (khb_group_from_staged)
e2e     the same job through khb_group_from_fasta with the text in pinned HOST memory: H2D copies of all
        text and D2H of every histogram inside the timed region
roofline  dominant kernel = onesweep_kernel (one radix digit pass): algorithmic bytes 2*W per key per launch,
        timed with CUDA events on the library's stream during the timed steps (khb_profile_*)
cpu_baseline  the CPU oracle (oracle/, OpenMP) on a bounded sample of the same workload, rank 0, N=1 only

Environment overrides for quick runs: KHB_BENCH_GROUPS, KHB_BENCH_GENOMES, KHB_BENCH_LEN, KHB_BENCH_K.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gbases/s to final k-mer occurrence tables (k=31)"
UNIT = "Gbases/s"


def env_int(name, default):
    return int(os.environ.get(name, default))


def workload():
    return {"groups_per_gpu": env_int("KHB_BENCH_GROUPS", 10), "genomes": env_int("KHB_BENCH_GENOMES", 50),
            "genome_len": env_int("KHB_BENCH_LEN", 5_000_000), "k": env_int("KHB_BENCH_K", 31)}


def _gen_group(args):
    from khoice_b200 import synth
    cfg, g = args
    return g, [synth.make_genome(cfg, g, i) for i in range(1, cfg.genomes_per_group + 1)]


def generate_groups(cfg, group_numbers, procs):
    """{group: [fasta bytes]} generated with a fork pool (must run before CUDA is initialised)."""
    if procs <= 1 or len(group_numbers) <= 1:
        return dict(_gen_group((cfg, g)) for g in group_numbers)
    with mp.get_context("fork").Pool(min(procs, len(group_numbers))) as pool:
        return dict(pool.imap_unordered(_gen_group, [(cfg, g) for g in group_numbers]))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms.  The process is started BEFORE the warm-up (its
    start-up alone takes a few hundred ms, longer than a short timed region); samples carry nvidia-smi's own
    timestamp and only those inside [begin(), end()] -- the timed region -- are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.t0 = self.t1 = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        import datetime
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[1]), float(c[2]), float(c[3]), c[4:8]))
            except ValueError:
                continue
        self.f.close()
        os.unlink(self.f.name)
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.05 <= r[0] <= (self.t1 or r[0]) + 0.05]
        where = "timed region"
        if not inside and rows and self.t0 is not None:
            # region shorter than the sampling period: the sample closest to it (taken under the same load, in the warm-up)
            inside = [min(rows, key=lambda r: abs(r[0] - self.t0))]
            where = "nearest sample (region shorter than the 100 ms period)"
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm = [r[1] for r in inside]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(r[2] for r in inside) if inside else None,
                "power_w_max": max(r[3] for r in inside) if inside else None, "samples": len(inside), "samples_total": len(rows),
                "window": where, "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic_ratio():
    """dram bytes / algorithmic bytes of onesweep_kernel from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "onesweep_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["dram_bytes_per_launch"]) / float(d["algorithmic_bytes_per_launch"])
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's own implementation of this path is KMC 3.2.1 on the host cores; it is
    not installable here (no source under /root/reference, no network), so the arm times the CPU oracle port
    (oracle/kmer_oracle.c, OpenMP over genomes and groups like `snakemake --cores`) on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from khoice_b200 import synth
    from oracle import oracle as O
    wl = workload()
    O.build()
    cores = O.num_threads()
    k = wl["k"]
    cfg1 = synth.SynthConfig(n_groups=1, genomes_per_group=1, genome_len=wl["genome_len"])
    probe = synth.make_genome(cfg1, 1, 1)
    t0 = time.time()
    O.exp1([probe], [0], 1, k)
    per_genome = max(time.time() - t0, 1e-3)
    total_steps = args.steps + args.warmup
    budget = 150.0 / max(total_steps, 1)                         # seconds per step
    n_groups, genomes = sample_shape(wl, per_genome, cores, budget)
    cfg = synth.SynthConfig(n_groups=n_groups, genomes_per_group=genomes, genome_len=wl["genome_len"])
    groups = generate_groups(cfg, list(range(1, n_groups + 1)), min(cores, n_groups))
    flat = [f for g in range(1, n_groups + 1) for f in groups[g]]
    gid = [g for g in range(n_groups) for _ in range(genomes)]
    bases = sum(synth.count_bases(f) for f in flat)
    for _ in range(args.warmup):
        O.exp1(flat, gid, n_groups, k)
    t0 = time.time()
    for _ in range(args.steps):
        O.exp1(flat, gid, n_groups, k)
    dt = (time.time() - t0) / max(args.steps, 1)
    val = bases / dt / 1e9
    sample = f"{n_groups} groups x {genomes} genomes x {wl['genome_len']} bp ({bases} bases) per step, k={k}"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": "bounded sample of config 2 (10 groups x 50 synthetic 5 Mbp genomes, k=31): " + sample,
                       "note": "KMC 3.2.1 (the reference's engine) is not in /root/reference and not installed: CPU oracle port timed"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def sample_shape(wl, per_genome, cores, seconds):
    """(groups, genomes per group) of the bounded CPU sample: the workload's own group size if the time allows, at most 4
    groups (about 20 GB of host memory at 50 x 5 Mbp).  per_genome = seconds of the whole oracle chain for ONE genome on one
    thread; the per-genome stage parallelises over genomes and every union over key ranges (oracle/ko_body.inc), measured
    at ~0.65 x per_genome x genomes / cores on 16 cores -- 0.8 keeps a margin."""
    est = lambda ng, g: per_genome * max(ng * g / cores, 1.0) * 0.8
    n_groups, genomes = 2, 2
    while genomes < wl["genomes"] and est(n_groups, genomes + 1) < seconds:
        genomes += 1
    while genomes == wl["genomes"] and n_groups < min(4, wl["groups_per_gpu"]) and est(n_groups + 1, genomes) < seconds:
        n_groups += 1
    return n_groups, genomes


def cpu_baseline(wl, seconds=20.0):
    from khoice_b200 import synth
    from oracle import oracle as O
    O.build()
    cores = O.num_threads()
    k = wl["k"]
    cfg1 = synth.SynthConfig(n_groups=1, genomes_per_group=1, genome_len=wl["genome_len"])
    probe = synth.make_genome(cfg1, 1, 1)
    t0 = time.time()
    O.exp1([probe], [0], 1, k)
    per_genome = max(time.time() - t0, 1e-3)
    n_groups, genomes = sample_shape(wl, per_genome, cores, seconds)
    cfg = synth.SynthConfig(n_groups=n_groups, genomes_per_group=genomes, genome_len=wl["genome_len"])
    groups = generate_groups(cfg, list(range(1, n_groups + 1)), min(cores, n_groups))
    flat = [f for g in range(1, n_groups + 1) for f in groups[g]]
    bases = sum(synth.count_bases(f) for f in flat)
    t0 = time.time()
    O.exp1(flat, [g for g in range(n_groups) for _ in range(genomes)], n_groups, k)
    dt = time.time() - t0
    return {"value": bases / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"CPU oracle (oracle/kmer_oracle.c, OpenMP) on {n_groups} groups x {genomes} genomes x {wl['genome_len']} bp, k={k}: "
                      f"{bases} bases in {dt:.1f} s; KMC3 itself is unavailable (not in /root/reference, not installed)"}


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="khoice_b200", choices=["khoice_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    from khoice_b200 import synth
    wl = workload()
    k = wl["k"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    n_groups_total = wl["groups_per_gpu"] * world
    cfg = synth.SynthConfig(n_groups=n_groups_total, genomes_per_group=wl["genomes"], genome_len=wl["genome_len"])

    # 1. synthetic data for this rank's groups (fork pool: before any CUDA initialisation)
    from khoice_b200.dist import groups_of_rank
    mine = groups_of_rank(n_groups_total, rank, world)
    procs = max(1, (os.cpu_count() or 8) // max(world, 1))
    t0 = time.time()
    groups = generate_groups(cfg, mine, min(procs, 16))
    gen_s = time.time() - t0

    import torch
    import torch.distributed as dist
    from khoice_b200 import dist as kd
    from khoice_b200.engine import Engine
    if world > 1:
        kd.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    eng = Engine(local)
    ext_stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    adapter = kd.CudaAdapter(eng, dev)

    # 2. pinned host copies (e2e leg) and HBM-resident staged copies (device leg)
    total_bytes = sum(len(f) for g in mine for f in groups[g])
    pinned = torch.empty(total_bytes, dtype=torch.uint8, pin_memory=True)
    pview = pinned.numpy()
    host_views, off = {}, 0
    for g in mine:
        host_views[g] = []
        for f in groups[g]:
            pview[off:off + len(f)] = np.frombuffer(f, dtype=np.uint8)
            host_views[g].append(pview[off:off + len(f)])
            off += len(f)
    staged = {g: eng.stage_fasta(host_views[g]) for g in mine}
    groups = None  # the bytes now live in pinned memory and in HBM

    def barrier():
        eng.sync()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # across-group stage: local (N=1); N>1: the k-mer space is hash-range partitioned and every group's new keys are stored
    # straight into their owner's receive buffer over NVLink by one kernel behind the group's K5 (csrc/peer.cu); the first
    # round (a warm-up step) runs over NCCL (partition + all-to-all) and sizes the regions.  KHB_EXCHANGE=nccl keeps NCCL.
    ex = kd.AcrossExchanger(adapter, k, n_groups_total, mode=os.environ.get("KHB_EXCHANGE", "peer"))

    def step_device():
        eng.group_sets_reset()
        ex.begin()
        hs, nb = {}, 0
        for g in mine:
            hs[g], st = eng.group_from_staged(staged[g], k)
            ex.after_group()
            nb += st["bases"]
        ha, _ = ex.finish()
        return hs, ha, nb

    pipelined = {"next": False}

    def step_e2e():
        # a stream of jobs: while the last group of a job and its across-group stage run, the first group of the NEXT job is
        # already being copied (every step still copies all of its text inside the timed region; only the very first step's
        # first group is copied ahead of its compute instead of behind the previous step's tail)
        eng.group_sets_reset()
        ex.begin()
        hs, nb = {}, 0
        if not pipelined["next"]:
            eng.prefetch_fasta(host_views[mine[0]])
        for i, g in enumerate(mine):
            nxt = mine[i + 1] if i + 1 < len(mine) else mine[0]
            eng.prefetch_fasta(host_views[nxt])                  # H2D of the next group overlaps this group's kernels
            hs[g], st = eng.group_from_fasta(host_views[g], k)   # uses the prefetched copy, waits for it on the device
            ex.after_group()
            nb += st["bases"]
        pipelined["next"] = True
        ha, _ = ex.finish()
        return hs, ha, nb

    def timed(fn, steps, profile=False, sampler=None):
        barrier()
        if profile:
            eng.profile_enable(True)
        if sampler:
            sampler.begin()
        launches0 = eng.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext_stream)
        t0 = time.time()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(ext_stream)
        barrier()
        wall = time.time() - t0
        ms = e0.elapsed_time(e1)
        prof = eng.profile_read() if profile else None
        if profile:
            eng.profile_enable(False)
        if sampler:
            sampler.end()
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t.item()), wall, eng.launch_count - launches0, prof, clocks

    # 3. device-resident leg
    sampler = ClockSampler(local) if rank == 0 else None  # started before the warm-up, filtered to the timed region
    for _ in range(args.warmup):
        ref_out = step_device()
    (hs, ha, nb), ms_dev, wall_dev, launches, prof, clocks = timed(step_device, args.steps, profile=True, sampler=sampler)
    # 4. end-to-end leg (host buffers)
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    (hs2, ha2, nb2), ms_e2e, wall_e2e, _, _, _ = timed(step_e2e, args.steps)
    same = all(np.array_equal(hs[g], hs2[g]) for g in mine) and np.array_equal(ha, ha2) and np.array_equal(ha, ref_out[1])
    if not same:
        print("FATAL: device-resident and end-to-end legs disagree", file=sys.stderr)
        return 2

    tb = torch.tensor([float(nb), float(total_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
    bases_all, bytes_all = float(tb[0].item()), float(tb[1].item())
    value = bases_all * args.steps / (ms_dev * 1e-3) / 1e9
    e2e_value = bases_all * args.steps / (ms_e2e * 1e-3) / 1e9

    if rank == 0:
        peak, peak_src = measured_peak()
        osw = prof["onesweep"]
        achieved = osw["alg_bytes"] / (osw["ms"] * 1e-3) / 1e9 if osw["ms"] > 0 else 0.0
        ratio = ncu_traffic_ratio()
        per_launch_alg = osw["alg_bytes"] / max(osw["launches"], 1)
        kernels = {name: {"launches": v["launches"], "ms": round(v["ms"], 3),
                          "alg_GBps": round(v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None,
                          "share_of_step": round(v["ms"] / ms_dev, 4)} for name, v in prof.items()}
        # whole step against the roofline: algorithmic bytes of every kernel launched in the timed steps (this design's own
        # per-kernel contracts) / device time of the steps.  Rank 0's kernels x world (weak scaling).
        total_alg = float(sum(v["alg_bytes"] for v in prof.values())) * world
        pipeline = {"algorithmic_bytes_per_base": total_alg / max(bases_all * args.steps, 1.0),
                    "achieved": total_alg / (ms_dev * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                    "frac": total_alg / (ms_dev * 1e-3) / 1e9 / (peak * world) if peak else None,
                    "note": "single-sort contract (~110 B/base); the KMC-shaped chain of SURVEY.md 8d (sort, unique, sort, count: ~329 B/base) "
                            "would need 3x these bytes for the same tables"}
        hist_bytes = (len(mine) + 1) * 5001 * 8
        default_shape = (wl["groups_per_gpu"], wl["genomes"], wl["genome_len"], k) == (10, 50, 5_000_000, 31)
        shape_name = "config 2" if default_shape else "custom shape (KHB_BENCH_* overrides)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{shape_name}: {wl['groups_per_gpu']} groups x {wl['genomes']} synthetic {wl['genome_len']} bp genomes per GPU, k={k}"
                                   + (f"; {world} GPUs, {n_groups_total} groups, hash-range all-to-all for the across-group stage" if world > 1 else ", single B200"),
                       "k": k, "groups_total": n_groups_total, "genomes_per_group": wl["genomes"], "bases_per_step": bases_all,
                       "l2": "inputs exceed L2: every sort streams >= 2 GB of keys through a 126 MB L2; no explicit flush",
                       "parallelism": f"groups dealt round-robin to {world} rank(s)", "data_gen_s": round(gen_s, 1),
                       "exchange": (f"peer-memory push (CUDA IPC over NVLink), {ex.rounds_peer} rounds; NCCL all-to-all, {ex.rounds_nccl} rounds (sizing / fallback)"
                                    if world > 1 else "none (one GPU)")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_all, "d2h_bytes_per_step": float(hist_bytes * world),
                    "ms_per_step": ms_e2e / args.steps,
                    "note": "steps are pipelined like a stream of jobs: the H2D copy of a step's first group overlaps the previous step's last group and "
                            "across-group stage; every step's text is copied inside the timed region"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "onesweep_kernel<Key64,14> (one 8-bit radix pass)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": (ratio * per_launch_alg) if ratio else None, "algorithmic_bytes_per_launch": per_launch_alg,
                         "launches": osw["launches"], "avg_launch_ms": osw["ms"] / max(osw["launches"], 1), "peak_source": peak_src},
            "pipeline_roofline": pipeline,
            "kernels": kernels,
            "clocks": clocks,
            "wall_ms_per_step": wall_dev * 1e3 / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line))
    ex.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
