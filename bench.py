#!/usr/bin/env python
"""bench.py -- headline benchmark of the khoice exp-type-1 k-mer path on B200.

Metric (BASELINE.json): Gbases/s to final k-mer occurrence tables (k=31).  One "step" = the whole job on
the workload below: every group's step_4 histogram plus the step_8 across-group histogram.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workloads (KHB_BENCH_CONFIG, default 2 -- the configuration BASELINE.json's metric is quoted on):
  2  10 groups x 50 synthetic 5 Mbp genomes PER GPU (2.5 Gbp per GPU), k=31; weak scaling
  3  the k sweep 7..31 (odd) over the config-2 set, 2-bit stream packed once (value = bases x k values / s)
  4  20 groups x 100 genomes in total (10 Gbp), k=47 (KHB_BENCH_K=63 for the other half), 128-bit words; strong scaling
  5  100 groups x 200 genomes in total (~100 Gbp), k=31, groups dealt over the ranks; strong scaling -- the north-star target

N>1 (torchrun, one rank per GPU): steps 1-6 need no collective; for the across-group stage every group's distinct k-mers are
pushed to their hash-range owner over peer memory (csrc/peer.cu) -- or one NCCL all-to-all (KHB_EXCHANGE=nccl) -- then a
local count and a histogram all-reduce (khoice_b200/dist.py).

value   Gbases/s with the FASTA text already staged in HBM when the clock starts (khb_group_from_staged)
e2e     the same job through khb_group_from_fasta with the text in pinned HOST memory: H2D copies of all
        text and D2H of every histogram inside the timed region
roofline  the kernel with the largest share of the step: algorithmic bytes per launch over its CUDA-event
        time on the library's stream during the timed steps (khb_profile_*)
pipeline_roofline  the whole step, on this design's own byte contract and on SURVEY.md 8(d)'s KMC-shaped contract
cpu_baseline  the CPU oracle (oracle/, OpenMP) on a bounded sample of the same workload, rank 0, N=1 only
parity_in_run  the GPU histograms of that sample == the oracle's (N=1); N>1: conservation laws of the all-reduced tables and
        one group of another rank recomputed on rank 0.  A false value fails the run (exit 2).

Overrides for quick runs: KHB_BENCH_GROUPS (per GPU), KHB_BENCH_GROUPS_TOTAL, KHB_BENCH_GENOMES, KHB_BENCH_LEN, KHB_BENCH_K,
KHB_BENCH_E2E=0 (skip the host-buffer leg), KHB_BENCH_CSV_DIR (write step_5 / step_9 CSVs of the run there).
"""
from __future__ import annotations

import argparse
import glob
import json
import multiprocessing as mp
import os
import shutil
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

CONFIGS = {
    "2": {"name": "config 2", "groups_per_gpu": 10, "genomes": 50, "k": 31, "scaling": "weak"},
    "3": {"name": "config 3 (k sweep 7..31 odd over the config-2 set)", "groups_per_gpu": 10, "genomes": 50, "k": 31, "scaling": "weak",
          "ks": list(range(7, 32, 2))},
    "4": {"name": "config 4", "groups_total": 20, "genomes": 100, "k": 47, "scaling": "strong"},
    "5": {"name": "config 5", "groups_total": 100, "genomes": 200, "k": 31, "scaling": "strong"},
}


def env_int(name, default):
    return int(os.environ.get(name, default))


def workload(world: int = 1):
    key = os.environ.get("KHB_BENCH_CONFIG", "2")
    if key not in CONFIGS:
        raise SystemExit(f"KHB_BENCH_CONFIG={key}: expected one of {sorted(CONFIGS)}")
    c = dict(CONFIGS[key])
    wl = {"config": key, "name": c["name"], "scaling": c["scaling"], "genomes": env_int("KHB_BENCH_GENOMES", c["genomes"]),
          "genome_len": env_int("KHB_BENCH_LEN", 5_000_000), "k": env_int("KHB_BENCH_K", c["k"]), "ks": c.get("ks")}
    if "KHB_BENCH_GROUPS_TOTAL" in os.environ:
        wl["groups_total"], wl["scaling"] = env_int("KHB_BENCH_GROUPS_TOTAL", 0), "strong"
    elif "KHB_BENCH_GROUPS" in os.environ or "groups_per_gpu" in c:
        wl["groups_per_gpu"] = env_int("KHB_BENCH_GROUPS", c.get("groups_per_gpu", 10))
        wl["groups_total"], wl["scaling"] = wl["groups_per_gpu"] * world, "weak"
    else:
        wl["groups_total"] = c["groups_total"]
    wl["groups_per_gpu"] = wl.get("groups_per_gpu", -(-wl["groups_total"] // world))
    overridden = any(v in os.environ for v in ("KHB_BENCH_GROUPS", "KHB_BENCH_GROUPS_TOTAL", "KHB_BENCH_GENOMES", "KHB_BENCH_LEN")) or \
        (key != "4" and "KHB_BENCH_K" in os.environ)
    wl["default_shape"] = not overridden
    return wl


def _gen_genome(args):
    from khoice_b200 import synth
    cfg, g, i = args
    return g, i, synth.make_genome(cfg, g, i)


def generate_groups(cfg, group_numbers, procs, genomes=None, genome_range=None):
    """{group: [fasta bytes]} generated with a fork pool (must run before CUDA is initialised).  genome_range = (lo, hi): only the
    genomes lo + 1 .. hi of every group (a team member's slice)."""
    n = genomes or cfg.genomes_per_group
    lo, hi = genome_range or (0, n)
    if procs <= 1 or len(group_numbers) * (hi - lo) <= 2:
        return {g: [synth_genome(cfg, g, i) for i in range(lo + 1, hi + 1)] for g in group_numbers}
    out = {g: [None] * (hi - lo) for g in group_numbers}
    jobs = [(cfg, g, i) for g in group_numbers for i in range(lo + 1, hi + 1)]   # per genome: few big groups still use every worker
    with mp.get_context("fork").Pool(min(procs, len(jobs))) as pool:
        for g, i, text in pool.imap_unordered(_gen_genome, jobs, chunksize=max(1, (hi - lo) // 8)):
            out[g][i - 1 - lo] = text
    return out


def synth_genome(cfg, g, i):
    from khoice_b200 import synth
    return synth.make_genome(cfg, g, i)


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


def ncu_traffic(kernel: str):
    """(dram bytes / algorithmic bytes, label) of `kernel` from the committed ncu capture (profiles/<kernel>_traffic.json), or
    (None, why).  The ratio is a property of the kernel variant named in the file, not measured in this run: ncu cannot run
    inside the timed region."""
    p = os.path.join(ROOT, "profiles", f"{kernel}_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["dram_bytes_per_launch"]) / float(d["algorithmic_bytes_per_launch"]), f"ncu --set full capture {d.get('captured', '?')} of {d.get('variant', kernel)}"
        except Exception as e:
            return None, f"unreadable {p}: {e}"
    return None, "no ncu capture committed for this kernel"


# Dominant-kernel descriptions for the roofline object: profile id -> (kernel, what bounds it)
KERNEL_INFO = {
    "onesweep": ("onesweep_kernel (one 8-bit radix pass over (key, genome id) records)", "hbm"),
    "pack": ("fasta_summary/scan/pack_kernel (K1)", "hbm"),
    "extract": ("extract64/128_kernel (K2)", "hbm"),
    "radix_hist": ("radix_hist_kernel", "hbm"),
    "unique": ("pairs/runs kernels (K4/K5/K6)", "hbm"),
    "rle_hist": ("rle_hist_kernel", "hbm"),
    "partition": ("partition / push kernels (K7)", "hbm"),
    "hash_insert": ("hash_insert_kernel", "hbm"),
    "hash_count": ("hash_count_kernel", "hbm"),
    "bin_partition": ("mb_partition_kernel (minimizer bins: packed symbol stream -> super-k-mer records, csrc/bins.cu)", "hbm"),
    "bin_count": ("mb_count_kernel (persistent CTAs, one bin at a time: distinct records, then a shared-memory (k-mer, genome bits) table, csrc/bins.cu)", "hbm"),
    "bin_across": ("mb_across_kernel (across-group stage bin by bin over the groups' key segments, csrc/bins.cu)", "hbm"),
}

KERNEL_NOTE = {
    "bin_count": "the minimizer-bin path moves ~2.4 bytes per window through HBM (24-byte super-k-mer records written and read once, 8 bytes per "
                 "DISTINCT k-mer out) instead of ~100 for the prefix sort it replaced, so its HBM fraction is small by construction: the kernel is "
                 "bound by shared-memory table work and instruction issue (profiles/r2_bin_count_top.txt); pipeline_roofline.survey_8d prices the "
                 "same tables on SURVEY.md 8(d)'s byte contract",
    "bin_partition": "bound by instruction issue (hashing 13-mers, sliding minimum, record emission), not HBM: it reads 3/8 byte and writes ~2.4 bytes "
                     "per window (profiles/r2_bin_partition_top.txt)",
}

# bytes per input base of the KMC-shaped chain (sort, unique, sort, count per genome / group / across) at
# W = 8 (k <= 32) with P = ceil(2k/8) passes:  K1 1.39 + K2 (3/8 + W) + K3 W(2P+1) + K4 2W + K3' W(2P+1) + K5 W(1+rho) + K3'' W(2P+1) rho + K6 W rho
def survey_bytes_per_base(k: int, rho: float) -> float:
    W = 8 if k <= 32 else 16
    P = -(-2 * k // 8)
    return 1.39 + (0.375 + W) + W * (2 * P + 1) + 2 * W + W * (2 * P + 1) + W * (1 + rho) + W * (2 * P + 1) * rho + W * rho


# ---------------------------------------------------------------------------------------------------------
def find_kmc():
    """SURVEY.md 8(c), last row: a real KMC 3 on PATH or under baseline/_ref/ (never shipped, never installable here)."""
    found = {}
    for exe in ("kmc", "kmc_tools"):
        p = shutil.which(exe)
        if not p:
            hits = [h for h in glob.glob(os.path.join(ROOT, "baseline", "_ref", "**", exe), recursive=True) if os.access(h, os.X_OK)]
            p = hits[0] if hits else None
        if p:
            found[exe] = p
    return found if len(found) == 2 else None


def run_kmc_chain(kmc, groups, k, work):
    """The UNMODIFIED shell strings of /root/reference/workflow/rules/exp_type_1.smk:163,173,182,191,241,250,259 on `groups`
    ({group: [fasta bytes]}) with a real KMC; returns (within hists, across hist, seconds).  Used for timing AND as the
    parity pin: the caller diffs these histograms with the oracle's and the GPU's."""
    import gzip
    from khoice_b200 import pipeline, tables
    env = dict(os.environ, PATH=os.path.dirname(kmc["kmc"]) + os.pathsep + os.path.dirname(kmc["kmc_tools"]) + os.pathsep + os.environ.get("PATH", ""))
    nums = sorted(groups)
    for n in nums:
        os.makedirs(os.path.join(work, f"data/dataset_{n}"), exist_ok=True)
        for i, text in enumerate(groups[n], 1):
            with gzip.open(os.path.join(work, f"data/dataset_{n}/g{i}.fna.gz"), "wb", compresslevel=1) as fd:
                fd.write(bytes(text))
    pipeline.write_complex_ops(work, [str(k)], len(nums))
    os.makedirs(os.path.join(work, "tmp"), exist_ok=True)
    t0 = time.time()

    def sh(cmd):
        subprocess.run(cmd, shell=True, check=True, cwd=work, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for n in nums:
        for g in pipeline.genomes_of(work, n):
            for d in (pipeline.p_step1(k, n, g), pipeline.p_step2(k, n, g)):
                os.makedirs(os.path.dirname(os.path.join(work, d)), exist_ok=True)
            sh(f"kmc -fm -m64 -k{k} -ci1 {pipeline.p_genome(n, g)} {pipeline.p_step1(k, n, g)} tmp/")
            sh(f"kmc_tools transform {pipeline.p_step1(k, n, g)} set_counts 1 {pipeline.p_step2(k, n, g)}")
        for d in (pipeline.p_step3(k, n), pipeline.p_step4(k, n), pipeline.p_step6(k, n)):
            os.makedirs(os.path.dirname(os.path.join(work, d)), exist_ok=True)
        sh(f"kmc_tools complex {pipeline.p_ops_within(k, n)}")
        sh(f"kmc_tools transform {pipeline.p_step3(k, n)} histogram {pipeline.p_step4(k, n)}")
        sh(f"kmc_tools transform {pipeline.p_step3(k, n)} set_counts 1 {pipeline.p_step6(k, n)}")
    for d in (pipeline.p_step7(k), pipeline.p_step8(k)):
        os.makedirs(os.path.dirname(os.path.join(work, d)), exist_ok=True)
    sh(f"kmc_tools complex {pipeline.p_ops_across(k)}")
    sh(f"kmc_tools transform {pipeline.p_step7(k)} histogram {pipeline.p_step8(k)}")
    dt = time.time() - t0
    # file row i is occurrence i + 1; the in-memory histograms are indexed by occurrence (index 0 unused)
    within = [[0] + tables.read_histogram_file(os.path.join(work, pipeline.p_step4(k, n))) for n in nums]
    across = [0] + tables.read_histogram_file(os.path.join(work, pipeline.p_step8(k)))
    return within, across, dt


def _same_hist(a, b):
    """Histograms equal up to trailing zero rows (KMC's own row count is its default, ours is 5000)."""
    a, b = np.asarray(a, dtype=np.uint64), np.asarray(b, dtype=np.uint64)
    n = max(a.size, b.size)
    return np.array_equal(np.pad(a, (0, n - a.size)), np.pad(b, (0, n - b.size)))


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


def cpu_sample(wl, seconds, steps=1, warmup=0):
    """The CPU arm on a bounded sample of the workload, with ALL host cores (explicitly: torchrun exports OMP_NUM_THREADS=1).
    A real KMC 3 (PATH or baseline/_ref) runs the unmodified rule chain -> kind "reference"; otherwise the oracle port.
    Returns (cpu_baseline dict, sample dict with the texts and the histograms for the in-run parity check)."""
    from khoice_b200 import synth
    from oracle import oracle as O
    O.build()
    cores = O.set_num_threads(O.host_cores())
    k = wl["k"]
    cfg1 = synth.SynthConfig(n_groups=1, genomes_per_group=1, genome_len=wl["genome_len"])
    probe = synth.make_genome(cfg1, 1, 1)
    t0 = time.time()
    O.exp1([probe], [0], 1, k)
    per_genome = max(time.time() - t0, 1e-3)
    n_groups, genomes = sample_shape(wl, per_genome, cores, seconds)
    cfg = synth.SynthConfig(n_groups=n_groups, genomes_per_group=genomes, genome_len=wl["genome_len"])
    groups = generate_groups(cfg, list(range(1, n_groups + 1)), min(cores, 16), genomes=genomes)
    flat = [f for g in range(1, n_groups + 1) for f in groups[g]]
    gid = [g for g in range(n_groups) for _ in range(genomes)]
    bases = sum(synth.count_bases(f) for f in flat)
    for _ in range(warmup):
        O.exp1(flat, gid, n_groups, k)
    t0 = time.time()
    for _ in range(max(steps, 1)):
        within, across, _ = O.exp1(flat, gid, n_groups, k)
    dt = (time.time() - t0) / max(steps, 1)
    shape = f"{n_groups} groups x {genomes} genomes x {wl['genome_len']} bp ({bases} bases), k={k}"
    base = {"value": bases / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"CPU oracle (oracle/kmer_oracle.c, OpenMP, {cores} threads) on {shape}: {dt:.1f} s per pass; "
                      "KMC3 itself is unavailable (not in /root/reference, not on PATH, nothing under baseline/_ref)"}
    kmc = find_kmc()
    kmc_diff = None
    if kmc:
        work = tempfile.mkdtemp(prefix="khb_kmc_")
        try:
            kw, ka, kdt = run_kmc_chain(kmc, groups, k, work)
            kmc_diff = {"within_equal_oracle": all(_same_hist(kw[i], within[i]) for i in range(n_groups)), "across_equal_oracle": _same_hist(ka, across)}
            base = {"value": bases / kdt / 1e9, "unit": UNIT, "cores": cores, "kind": "reference",
                    "sample": f"KMC3 ({kmc['kmc']}) + kmc_tools, the unmodified exp_type_1.smk shell strings, on {shape}: {kdt:.1f} s; "
                              f"histograms vs the oracle: {kmc_diff}"}
        except Exception as e:  # a broken binary must not take the bench down: report and keep the port
            base["sample"] += f"; a KMC binary was found ({kmc['kmc']}) but its rule chain failed: {e}"
        finally:
            shutil.rmtree(work, ignore_errors=True)
    return base, {"groups": groups, "n_groups": n_groups, "genomes": genomes, "within": within, "across": across, "bases": bases,
                  "shape": shape, "kmc_diff": kmc_diff, "seconds": dt}


def run_reference(args):
    """Reference arm: the reference's own implementation of this path is KMC 3.2.1 on the host cores.  If a KMC binary is on
    PATH or under baseline/_ref it runs the unmodified rule chain; otherwise (this image: no source under /root/reference,
    no network) the arm times the CPU oracle port (oracle/kmer_oracle.c, OpenMP over genomes and key ranges, like
    `snakemake --cores`) -- on a bounded sample, on rank 0 only, with every host core whatever OMP_NUM_THREADS says."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    wl = workload(int(os.environ.get("WORLD_SIZE", "1")))
    total_steps = args.steps + args.warmup
    base, smp = cpu_sample(wl, 150.0 / max(total_steps, 1), steps=args.steps, warmup=args.warmup)
    val = base["value"]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": smp["bases"] / val / 1e6, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "u64" if wl["k"] <= 32 else "u128", "data": "synthetic",
            "config": {"workload": f"bounded sample of {wl['name']}: " + smp["shape"] + " per step",
                       "note": "host CPU arm; its value is a throughput on the sample and does not depend on the GPU count"},
            "cpu_baseline": base,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------
def write_csvs(out_dir, wl, within, across_by_k, n_groups_total):
    """step_4 / step_8 histogram files and the step_5 / step_9 CSVs of this run through the product's own writers
    (khoice_b200.tables, the rules within_group_union_analysis / across_group_union_analysis)."""
    from khoice_b200 import pipeline, tables
    os.makedirs(out_dir, exist_ok=True)
    work = tempfile.mkdtemp(prefix="khb_bench_root_")
    try:
        for n in range(1, n_groups_total + 1):   # group membership = the *.fna.gz listing (exp_type_1.smk:107-113)
            d = os.path.join(work, f"data/dataset_{n}")
            os.makedirs(d)
            for i in range(wl["genomes"]):
                open(os.path.join(d, f"g{i + 1}.fna.gz"), "wb").close()
        ks = sorted(across_by_k)
        for k in ks:
            for n in range(1, n_groups_total + 1):
                tables.write_histogram_file(os.path.join(work, pipeline.p_step4(str(k), n)), within[k][n - 1])
            tables.write_histogram_file(os.path.join(work, pipeline.p_step8(str(k))), across_by_k[k])
        pipeline.build_tables(work, [str(k) for k in ks], n_groups_total)
        for p in (pipeline.P_STEP5, pipeline.P_STEP9):
            shutil.copyfile(os.path.join(work, p), os.path.join(out_dir, os.path.basename(p)))
    finally:
        shutil.rmtree(work, ignore_errors=True)


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
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = workload(world)
    k = wl["k"]
    ks = wl["ks"] or [k]
    if world != args.gpus and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    n_groups_total = wl["groups_total"]
    cfg = synth.SynthConfig(n_groups=n_groups_total, genomes_per_group=wl["genomes"], genome_len=wl["genome_len"])

    # 1. synthetic data for this rank's groups (fork pool: before any CUDA initialisation)
    from khoice_b200.dist import genome_slices, groups_of_rank, team_shape
    # Teams: where whole groups do not divide evenly over the GPUs of a fixed-size job (config 4 on 8 GPUs: 20 groups), every group is
    # sharded over the T members of a team -- genomes over members, minimizer bins over owners (dist.TeamSharder) -- and the groups are
    # dealt to the world / T teams.  KHB_BENCH_TEAM=T forces a team size (1: whole groups).
    T = env_int("KHB_BENCH_TEAM", 0) or (team_shape(n_groups_total, world) if wl["scaling"] == "strong" and not wl["ks"] else 1)
    if T < 1 or world % T or T > 8 or wl["genomes"] < T:
        raise SystemExit(f"KHB_BENCH_TEAM={T}: the team size must divide the {world} ranks, be at most 8 and at most the genomes of a group")
    n_teams, team_idx, member = world // T, rank // T, rank % T
    mine = groups_of_rank(n_groups_total, team_idx, n_teams)
    slices = genome_slices(wl["genomes"], T)
    slice_sizes = [hi - lo for lo, hi in slices]
    procs = max(1, (os.cpu_count() or 8) // max(world, 1))
    t0 = time.time()
    groups = generate_groups(cfg, mine, min(procs, 32), genome_range=slices[member] if T > 1 else None)
    # N>1 parity: rank 0 recomputes one group owned by another rank (teams: one group of another team -- or, with a single team, its
    # own first group -- unsharded on rank 0)
    foreign = None
    if world > 1 and rank == 0:
        fg = groups_of_rank(n_groups_total, 1, n_teams) if n_teams > 1 else mine[:1]
        if fg:
            foreign = (fg[0], generate_groups(cfg, [fg[0]], min(procs, 32))[fg[0]])
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
    do_e2e = env_int("KHB_BENCH_E2E", 1) != 0

    # 2. pinned host copies (e2e leg) and HBM-resident staged copies (device leg)
    total_bytes = sum(len(f) for g in mine for f in groups[g])
    pinned = torch.empty(max(total_bytes, 1), dtype=torch.uint8, pin_memory=do_e2e)
    pview = pinned.numpy()
    host_views, off = {}, 0
    for g in mine:
        host_views[g] = []
        for f in groups[g]:
            pview[off:off + len(f)] = np.frombuffer(f, dtype=np.uint8)
            host_views[g].append(pview[off:off + len(f)])
            off += len(f)
    groups = None  # the bytes now live in pinned memory (and, below, in HBM)
    sweep = wl["ks"] is not None
    staged, packed = {}, {}
    if sweep:
        packed = {g: eng.pack_group(host_views[g]) for g in mine}      # config 3: K1 once, the 2-bit stream stays in HBM
    else:
        staged = {g: eng.stage_fasta(host_views[g]) for g in mine}

    def barrier():
        eng.sync()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # across-group stage: local (N=1); N>1: the k-mer space is hash-range partitioned and every group's new keys are stored
    # straight into their owner's receive buffer over NVLink by one kernel behind the group's K5 (csrc/peer.cu); the first
    # round (a warm-up step) runs over NCCL (partition + all-to-all) and sizes the regions.  KHB_EXCHANGE=nccl keeps NCCL.
    ex = kd.AcrossExchanger(adapter, k, n_groups_total, mode=os.environ.get("KHB_EXCHANGE", "peer"))
    ts, group_syms, team_pg = None, {}, None
    if T > 1:
        for t in range(n_teams):                # every rank creates every team's process group, in the same order
            pg = dist.new_group(ranks=list(range(t * T, (t + 1) * T)))
            if t == team_idx:
                team_pg = pg
        ts = kd.TeamSharder(eng, T, member, group=team_pg)
        # symbols of every whole group, the same number on every member: the text bytes of all slices (an estimate is all the planner needs)
        sz = torch.tensor([sum(len(v) for v in host_views[g]) for g in mine], dtype=torch.int64, device=dev)
        dist.all_reduce(sz, group=team_pg)
        group_syms = {g: int(x) for g, x in zip(mine, sz.tolist())}

    def team_group(src, g, kk):
        return ts.run_group(src, kk, wl["genomes"], slice_sizes, group_syms[g])

    def run_k(kk, group_fn):
        eng.group_sets_reset()
        ex.set_k(kk)
        ex.begin()
        hs, nb, dsum = {}, 0, 0
        for i, g in enumerate(mine):
            hs[g], st = group_fn(i, g, kk)
            ex.after_group()
            nb += st["bases"]
            dsum += st["distinct"]
        ha, _ = ex.finish()
        if ts and mine:
            # every member holds the rows of ITS bins' k-mers; which bins those are moves with the planner's hints, the sum does not:
            # one all-reduce in the team per k gives every member the groups' step_4 histograms
            t = torch.from_numpy(np.stack([hs[g] for g in mine]).astype(np.int64)).to(dev)
            dist.all_reduce(t, group=team_pg)
            t = t.cpu().numpy().astype(np.uint64)
            hs = {g: t[i] for i, g in enumerate(mine)}
        return hs, ha, nb, dsum

    def step_device():
        out = {}
        for kk in ks:
            if ts:
                out[kk] = run_k(kk, lambda i, g, kk: team_group(packed[g] if sweep else staged[g], g, kk))
            elif sweep:
                out[kk] = run_k(kk, lambda i, g, kk: eng.group_from_packed(packed[g], kk))
            else:
                out[kk] = run_k(kk, lambda i, g, kk: eng.group_from_staged(staged[g], kk))
        return out

    pipelined = {"next": False}

    def e2e_group(i, g, kk):
        nxt = mine[i + 1] if i + 1 < len(mine) else mine[0]
        eng.prefetch_fasta(host_views[nxt])                  # H2D of the next group overlaps this group's kernels
        if ts:
            return team_group(host_views[g], g, kk)
        return eng.group_from_fasta(host_views[g], kk)       # uses the prefetched copy, waits for it on the device

    def step_e2e():
        # a stream of jobs: while the last group of a job and its across-group stage run, the first group of the NEXT job is
        # already being copied (every step still copies all of its text inside the timed region; only the very first step's
        # first group is copied ahead of its compute instead of behind the previous step's tail)
        if sweep:
            # config 3: H2D + K1 once per step, then the sweep on the resident 2-bit stream
            pk = {g: eng.pack_group(host_views[g]) for g in mine}
            try:
                if ts:
                    return {kk: run_k(kk, lambda i, g, kk: team_group(pk[g], g, kk)) for kk in ks}
                return {kk: run_k(kk, lambda i, g, kk: eng.group_from_packed(pk[g], kk)) for kk in ks}
            finally:
                for p in pk.values():
                    p.free()
        if not pipelined["next"] and mine:
            eng.prefetch_fasta(host_views[mine[0]])
        pipelined["next"] = True
        return {k: run_k(k, e2e_group)}

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

    def equal_runs(a, b):
        return all(all(np.array_equal(a[kk][0][g], b[kk][0][g]) for g in mine) and np.array_equal(a[kk][1], b[kk][1]) for kk in ks)

    # 3. device-resident leg
    sampler = ClockSampler(local) if rank == 0 else None  # started before the warm-up, filtered to the timed region
    ref_out = None
    for _ in range(args.warmup):
        ref_out = step_device()
    out_dev, ms_dev, wall_dev, launches, prof, clocks = timed(step_device, args.steps, profile=True, sampler=sampler)
    same = ref_out is None or equal_runs(out_dev, ref_out)
    # 4. end-to-end leg (host buffers)
    ms_e2e = None
    if do_e2e:
        for _ in range(min(args.warmup, 2)):
            step_e2e()
        out_e2e, ms_e2e, wall_e2e, _, _, _ = timed(step_e2e, args.steps)
        same = same and equal_runs(out_dev, out_e2e)
    if not same:
        print("FATAL: device-resident, warm-up and end-to-end legs disagree", file=sys.stderr)
        return 2

    nb = sum(out_dev[kk][2] for kk in ks)          # bases processed per step (config 3: every k counts the set again)
    tb = torch.tensor([float(nb), float(total_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
    bases_all, bytes_all = float(tb[0].item()), float(tb[1].item())
    value = bases_all * args.steps / (ms_dev * 1e-3) / 1e9
    e2e_value = bases_all * args.steps / (ms_e2e * 1e-3) / 1e9 if ms_e2e else None

    # 5. every rank's within-group histograms -> one table (what run_fused_distributed writes as step_4), conservation laws
    nbins1 = 5001
    within_all, dsum_all = {}, {}
    for kk in ks:
        w = np.zeros((n_groups_total, nbins1 + 1), dtype=np.int64)   # last column: distinct k-mers of the group
        for g in (mine if member == 0 else []):      # teams: every member holds the team's sum, one of them contributes it
            w[g - 1, :nbins1] = out_dev[kk][0][g].astype(np.int64)
            w[g - 1, nbins1] = int(out_dev[kk][0][g][1:].sum())
        wt = torch.from_numpy(w).to(dev)
        if world > 1:
            dist.all_reduce(wt, op=dist.ReduceOp.SUM)
        within_all[kk] = wt.cpu().numpy()
        dsum_all[kk] = int(within_all[kk][:, nbins1].sum())
    parity = {"checks": []}
    ok = True
    for kk in ks:
        ha = out_dev[kk][1].astype(np.int64)
        c = np.arange(ha.size, dtype=np.int64)
        # every (k-mer, group) pair is counted exactly once: sum_c c * H[c] = sum_G D_G  (no counter reaches the 5000 cap here)
        law = int((c * ha).sum()) == dsum_all[kk] and int(ha[0]) == 0
        parity["checks"].append({"k": kk, "sum_c_times_H_equals_sum_D_G": law, "sum_D_G": dsum_all[kk], "distinct_overall": int(ha.sum())})
        ok = ok and law
    if foreign is not None:
        fg, texts = foreign
        eng.group_sets_reset()
        h1, _ = eng.group_from_fasta(texts, k, keep_set=False)
        eq = bool(np.array_equal(h1.astype(np.int64), within_all[k][fg - 1, :nbins1]))
        parity["checks"].append({"group_of_rank_1_recomputed_on_rank_0": fg, "equal": eq})
        ok = ok and eq
        # N>1: the CPU oracle on two sampled groups at full size -- one of rank 0's own and the one of rank 1 -- against the table all ranks built
        if env_int("KHB_BENCH_ORACLE_GROUPS", 1) and not sweep:
            from oracle import oracle as O
            O.build()
            O.set_num_threads(O.host_cores())
            t_or = time.time()
            samples = [("of_rank_1" if n_teams > 1 or T == 1 else "own", fg, texts)]
            if T == 1:
                samples.insert(0, ("own", mine[0], list(host_views[mine[0]])))
            for label, gnum, tx in samples:
                w_or, _, _ = O.exp1(tx, [0] * len(tx), 1, k)
                eq = bool(np.array_equal(np.asarray(w_or[0][:nbins1], dtype=np.int64), within_all[k][gnum - 1, :nbins1]))
                parity["checks"].append({"oracle_group": gnum, "which": label, "genomes": len(tx), "within_equal_oracle": eq})
                ok = ok and eq
            parity["oracle_seconds"] = round(time.time() - t_or, 1)
        foreign = None

    rc = 0
    if rank == 0:
        peak, peak_src = measured_peak()
        live = {name: v for name, v in prof.items() if v["launches"]}
        dom = max(live, key=lambda n: live[n]["ms"])
        dk = live[dom]
        achieved = dk["alg_bytes"] / (dk["ms"] * 1e-3) / 1e9 if dk["ms"] > 0 else 0.0
        per_launch_alg = dk["alg_bytes"] / max(dk["launches"], 1)
        ratio, ratio_src = ncu_traffic(dom)
        kernels = {name: {"launches": v["launches"], "ms": round(v["ms"], 3),
                          "alg_GBps": round(v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None,
                          "frac_of_peak": round(v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9 / peak, 3) if v["ms"] > 0 else None,
                          "share_of_step": round(v["ms"] / ms_dev, 4)} for name, v in live.items()}
        # whole step against the roofline: algorithmic bytes of every kernel launched in the timed steps (this design's own
        # per-kernel contracts) / device time of the steps.  Rank 0's kernels x world.
        total_alg = float(sum(v["alg_bytes"] for v in prof.values())) * world
        rho = dsum_all[k] / max(sum(out_dev[k][0][g][1:].astype(np.float64) @ np.arange(1, nbins1) for g in mine) * n_teams, 1.0) if mine else 0.0
        sv = survey_bytes_per_base(k, rho)
        pipeline = {"algorithmic_bytes_per_base": total_alg / max(bases_all * args.steps, 1.0),
                    "achieved": total_alg / (ms_dev * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                    "frac": total_alg / (ms_dev * 1e-3) / 1e9 / (peak * world) if peak else None,
                    "note": "this design's own byte contract: what its kernels must read and write once (DESIGN.md section 4)",
                    "survey_8d": {"algorithmic_bytes_per_base": sv, "rho": rho, "achieved": sv * value, "frac": sv * value / (peak * world),
                                  "note": "SURVEY.md 8(d)'s KMC-shaped chain (sort, unique, sort, count; P = ceil(2k/8) passes) priced at this run's "
                                          "Gbases/s: above 1.0 means the same tables are produced faster than that chain could run at the HBM peak"}}
        hist_bytes = (len(mine) + 1) * 5001 * 8 * len(ks)
        shape_name = wl["name"] if wl["default_shape"] else wl["name"] + " shape with KHB_BENCH_* overrides"
        per = f"{wl['groups_per_gpu']} groups x {wl['genomes']} synthetic {wl['genome_len']} bp genomes per GPU" if wl["scaling"] == "weak" else \
            f"{n_groups_total} groups x {wl['genomes']} synthetic {wl['genome_len']} bp genomes in total"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "u64" if max(ks) <= 32 else "u128", "data": "synthetic",
            "config": {"workload": f"{shape_name}: {per}, k={k if not sweep else ','.join(map(str, ks))}"
                                   + ((f"; {world} GPUs, groups dealt round-robin" + (f" to {n_teams} team(s) of {T} GPUs, every group sharded inside its team "
                                                                                              f"(genomes over members, minimizer bins over owners, records over NVLink)" if T > 1 else "")
                                       + ", hash-range exchange for the across-group stage") if world > 1 else ", single B200"),
                       "k": k if not sweep else ks, "groups_total": n_groups_total, "genomes_per_group": wl["genomes"], "bases_per_step": bases_all,
                       "l2": "inputs exceed L2: every group streams >= 1 GB through a 126 MB L2; no explicit flush",
                       "parallelism": (f"groups dealt round-robin to {world} rank(s)" if T == 1 else
                                       f"groups dealt round-robin to {n_teams} team(s) of {T} ranks; genome slices {slice_sizes}; "
                                       f"{ts.retries} repartitions, {ts.setups} buffer set-ups"), "team_size": T, "data_gen_s": round(gen_s, 1),
                       "group_mode": os.environ.get("KHB_GROUP_MODE", "auto"), "bins_counters": eng.bins_counters,
                       "exchange": (f"peer-memory push (CUDA IPC over NVLink), {ex.rounds_peer} rounds; NCCL all-to-all, {ex.rounds_nccl} rounds (sizing / fallback)"
                                    if world > 1 else "none (one GPU)")},
            "e2e": ({"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_all, "d2h_bytes_per_step": float(hist_bytes * world),
                     "ms_per_step": ms_e2e / args.steps,
                     "note": "steps are pipelined like a stream of jobs: the H2D copy of a step's first group overlaps the previous step's last group and "
                             "across-group stage; every step's text is copied inside the timed region"} if ms_e2e else None),
            "gpu_launches": launches,
            "roofline": {"bound": KERNEL_INFO.get(dom, (dom, "hbm"))[1], "kernel": KERNEL_INFO.get(dom, (dom, "hbm"))[0], "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": (ratio * per_launch_alg) if ratio else None, "traffic_source": ratio_src,
                         "algorithmic_bytes_per_launch": per_launch_alg,
                         "launches": dk["launches"], "avg_launch_ms": dk["ms"] / max(dk["launches"], 1), "peak_source": peak_src,
                         "share_of_step": dk["ms"] / ms_dev,
                         "note": KERNEL_NOTE.get(dom)},
            "pipeline_roofline": pipeline,
            "kernels": kernels,
            "clocks": clocks,
            "wall_ms_per_step": wall_dev * 1e3 / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            base, smp = cpu_sample(wl, 20.0)
            line["cpu_baseline"] = base
            # in-run parity: the GPU on exactly the sample's texts against the oracle's histograms
            eng.group_sets_reset()
            eqw = []
            for g in range(1, smp["n_groups"] + 1):
                hg, _ = eng.group_from_fasta(smp["groups"][g], k)
                eqw.append(bool(np.array_equal(hg, smp["within"][g - 1])))
            hga, _ = eng.across_groups()
            eqa = bool(np.array_equal(hga, smp["across"]))
            full = smp["genomes"] == wl["genomes"]
            eqt = all(np.array_equal(out_dev[k][0][g], smp["within"][g - 1]) for g in range(1, smp["n_groups"] + 1) if g in out_dev[k][0]) if full else None
            parity["checks"].append({"oracle_sample": smp["shape"], "within_equal": eqw, "across_equal": eqa,
                                     "timed_run_groups_equal_oracle": eqt, "kmc": smp["kmc_diff"]})
            ok = ok and all(eqw) and eqa and (eqt is not False)
            if smp["kmc_diff"]:
                ok = ok and all(smp["kmc_diff"].values())
        parity["ok"] = bool(ok)
        line["parity_in_run"] = bool(ok)
        line["parity"] = parity
        csv_dir = os.environ.get("KHB_BENCH_CSV_DIR")
        if csv_dir:
            write_csvs(csv_dir, wl, {kk: within_all[kk][:, :nbins1].astype(np.uint64) for kk in ks}, {kk: out_dev[kk][1] for kk in ks}, n_groups_total)
            line["config"]["csv_dir"] = csv_dir
        print(json.dumps(line))
        if not ok:
            print("FATAL: in-run parity check failed: " + json.dumps(parity), file=sys.stderr)
            rc = 2
    ex.close()
    if ts:
        ts.close()
    for p in packed.values():
        p.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return rc


if __name__ == "__main__":
    sys.exit(main())
