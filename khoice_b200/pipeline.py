"""Host driver of experiment type 1 on the B200 engine: the rule chain of
/root/reference/workflow/rules/exp_type_1.smk with the same rule names, inputs and outputs.

Two execution modes over the same work-root layout (``data/dataset_{n}/{genome}.fna.gz`` in,
``step_4`` / ``step_8`` histograms, ``step_5`` / ``step_9`` CSVs and ``final_results_type1/`` out):

* ``fused``  -- one pass per (k, group): the four per-genome / per-group rules collapse into
  ``Engine.group_from_fasta`` and the two across-group rules into ``Engine.across_groups``; the
  intermediate databases of steps 1,2,3,6,7 are written as header-only stubs so the file DAG and
  Snakemake's resume semantics still hold.  This is the fast path.
* ``rules``  -- every rule instance runs separately through the kmc / kmc_tools shims
  (khoice_b200/cli.py), exchanging real intermediate databases, exactly like the reference does with
  KMC.  Either in-process (one CUDA context) or as sub-processes with khoice_b200/bin first on PATH
  (the literal drop-in: the shell strings are the reference's own).

Snakemake is not installed in this image, so ``run_rules`` is a minimal stand-in for its scheduler for
these ten rules: it walks the same DAG, skips rule instances whose outputs exist (resume), and removes
the outputs of a failed job.
"""
from __future__ import annotations

import json
import os
import shlex
import shutil
import subprocess
import time
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import cli, ingest, kmcdb, tables
from .engine import COUNTER_MAX, Engine

# /root/reference/workflow/Snakefile:36 -- a Python literal there; a config key (K_VALUES) here.
DEFAULT_K_VALUES = [str(x) for x in range(7, 31, 1)] + [str(x) for x in range(34, 50, 3)]
BIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin")


# ---- layout helpers (paths exactly as in the rules) ---------------------------------------------------
def genomes_of(work_root: str, num: int) -> List[str]:
    """Genome names of data/dataset_{num}: text before ``.fna.gz`` (exp_type_1.smk:44-47), sorted."""
    d = os.path.join(work_root, "data", f"dataset_{num}")
    return sorted(f.split(".fna.gz")[0] for f in os.listdir(d) if f.endswith(".fna.gz"))


def p_genome(num, g): return f"data/dataset_{num}/{g}.fna.gz"
def p_step1(k, num, g): return f"step_1/k_{k}/dataset_{num}/{g}"
def p_step2(k, num, g): return f"step_2/k_{k}/dataset_{num}/{g}.transformed"
def p_step3(k, num): return f"step_3/k_{k}/dataset_{num}/dataset_{num}.transformed.combined"
def p_step4(k, num): return f"step_4/k_{k}/dataset_{num}/dataset_{num}_k{k}_hist.txt"
def p_step6(k, num): return f"step_6/k_{k}/dataset_{num}/dataset_{num}.transformed.combined.transformed"
def p_step7(k): return f"step_7/k_{k}/all_datasets.transformed.combined.transformed.combined"
def p_step8(k): return f"step_8/k_{k}/all_datasets_k{k}_hist.txt"
P_STEP5 = "step_5/within_datasets_analysis.csv"
P_STEP9 = "step_9/across_datasets_analysis.csv"
P_FINAL = ("final_results_type1/within_datasets_analysis.csv", "final_results_type1/across_datasets_analysis.csv")
def p_ops_within(k, num): return f"complex_ops/within_groups/k_{k}/dataset_{num}/within_dataset_{num}.txt"
def p_ops_across(k): return f"complex_ops/across_groups/k_{k}/across_all_datasets.txt"


def write_complex_ops(work_root: str, k_values: Sequence[str], num_datasets: int) -> None:
    """The parse-time block of exp_type_1.smk:26-84: tmp/, and one `kmc_tools complex` operation file
    per (k, group) and per k."""
    os.makedirs(os.path.join(work_root, "tmp"), exist_ok=True)

    def emit(path, inputs, output):
        full = os.path.join(work_root, path)
        os.makedirs(os.path.dirname(full), exist_ok=True)
        lines = ["INPUT:"] + [f"set{i + 1} = {p}" for i, p in enumerate(inputs)]
        expr = "(" + " + ".join(f"set{i + 1}" for i in range(len(inputs))) + ")"
        lines += ["OUTPUT:", f"{output} = {expr}", "OUTPUT_PARAMS:", "-cs5000"]
        with open(full, "w") as fd:
            fd.write("\n".join(lines) + "\n")

    for k in k_values:
        for num in range(1, num_datasets + 1):
            emit(p_ops_within(k, num), [p_step2(k, num, g) for g in genomes_of(work_root, num)], p_step3(k, num))
        emit(p_ops_across(k), [p_step6(k, num) for num in range(1, num_datasets + 1)], p_step7(k))


def _members_of(work_root: str):
    return lambda num: tables.get_num_of_dataset_members(num, os.path.join(work_root, "data"))


def build_tables(work_root: str, k_values: Sequence[str], num_datasets: int) -> None:
    """Rules within_group_union_analysis, across_group_union_analysis, copy_final_results_type1."""
    nums = range(1, num_datasets + 1)
    tables.within_group_union_analysis([os.path.join(work_root, p_step4(k, n)) for k in k_values for n in nums],
                                       os.path.join(work_root, P_STEP5), num_datasets, _members_of(work_root))
    tables.across_group_union_analysis([os.path.join(work_root, p_step8(k)) for k in k_values],
                                       os.path.join(work_root, P_STEP9), num_datasets)
    for src, dst in zip((P_STEP5, P_STEP9), P_FINAL):
        os.makedirs(os.path.dirname(os.path.join(work_root, dst)), exist_ok=True)
        shutil.copyfile(os.path.join(work_root, src), os.path.join(work_root, dst))


# ---- fused mode ----------------------------------------------------------------------------------------
def run_fused(work_root: str, num_datasets: int, k_values: Optional[Sequence] = None, engine: Optional[Engine] = None,
              stubs: bool = True, report_path: Optional[str] = None) -> Dict:
    """All of exp type 1 for ``work_root``; returns (and optionally writes) a JSON-able run report."""
    k_values = [str(k) for k in (k_values or DEFAULT_K_VALUES)]
    own = engine is None
    eng = engine or Engine(int(os.environ.get("KHB_DEVICE", "0")))
    report = {"mode": "fused", "work_root": work_root, "num_datasets": num_datasets, "k_values": k_values, "stages": []}
    t_start = time.time()
    packed: Dict[int, object] = {}
    try:
        write_complex_ops(work_root, k_values, num_datasets)
        names = {n: genomes_of(work_root, n) for n in range(1, num_datasets + 1)}
        zero = np.zeros(tables.HIST_ROWS + 1, dtype=np.uint64)
        # parallel inflate, one group ahead of the GPU (khoice_b200/ingest.py)
        reader = ingest.GroupReader({n: [os.path.join(work_root, p_genome(n, g)) for g in names[n]] for n in names}, sorted(names))
        for k in k_values:
            ki = int(k)
            eng.group_sets_reset()
            for num in range(1, num_datasets + 1):
                if num not in packed:  # inflate + pack ONCE; the 2-bit stream (3/8 byte per base) stays in HBM for the k sweep
                    texts = reader.get(num)
                    packed[num] = eng.pack_group(texts)
                    del texts
                hist, st = eng.group_from_packed(packed[num], ki, nbins=tables.HIST_ROWS, keep_set=True)
                tables.write_histogram_file(os.path.join(work_root, p_step4(k, num)), hist)
                if stubs:
                    one = zero.copy()
                    for g in names[num]:
                        kmcdb.write_db(os.path.join(work_root, p_step1(k, num, g)), ki, None, None, zero, cli.KMC_DEFAULT_CS)
                        kmcdb.write_db(os.path.join(work_root, p_step2(k, num, g)), ki, None, None, zero, cli.KMC_DEFAULT_CS)
                    kmcdb.write_db(os.path.join(work_root, p_step3(k, num)), ki, None, None, hist, COUNTER_MAX, st["distinct"])
                    one[1] = st["distinct"]
                    kmcdb.write_db(os.path.join(work_root, p_step6(k, num)), ki, None, None, one, COUNTER_MAX, st["distinct"])
                report["stages"].append({"k": ki, "group": num, **st})
            hist, st = eng.across_groups(nbins=tables.HIST_ROWS)
            tables.write_histogram_file(os.path.join(work_root, p_step8(k)), hist)
            if stubs:
                kmcdb.write_db(os.path.join(work_root, p_step7(k)), ki, None, None, hist, COUNTER_MAX, st["distinct"])
            report["stages"].append({"k": ki, "group": "across", **st})
        build_tables(work_root, k_values, num_datasets)
    finally:
        if "reader" in locals():
            reader.close()
        for pk in packed.values():
            pk.free()
        if own:
            eng.close()
    report["seconds"] = time.time() - t_start
    if report_path:
        with open(report_path, "w") as fd:
            json.dump(report, fd, indent=1)
    return report


# ---- fused mode on several GPUs ------------------------------------------------------------------------
def run_fused_distributed(work_root: str, num_datasets: int, k_values: Optional[Sequence] = None, adapter=None, exchange: str = "peer",
                          stubs: bool = True, report_path: Optional[str] = None, team: int = 0) -> Dict:
    """``run_fused`` under ``torchrun`` (one process per GPU, ``torch.distributed`` already initialised or initialisable from
    the environment): groups are dealt to the ranks round-robin (khoice_b200/dist.py), every rank inflates and packs only
    its own groups once and sweeps k over them; per k the across-group stage goes through ``dist.AcrossExchanger`` (peer-
    memory push, or ``exchange="nccl"``), the per-group histograms are all-reduced, and rank 0 writes the step_4 / step_8
    files, the stubs and the step_5 / step_9 CSVs -- byte-identical to a single-GPU run.  ``adapter``: tests pass a
    stand-in; the product adapter is dist.CudaAdapter on this rank's GPU.

    ``team``: GPUs per group.  1 = whole groups dealt to ranks; T > 1 = the ranks form world / T teams, the groups are dealt to the teams
    and every group is sharded inside its team (dist.TeamSharder: genomes over members, minimizer bins over owners) -- for work roots
    with fewer, or unevenly many, groups than GPUs; 0 = the smallest T that deals the groups evenly (dist.team_shape), provided every
    k of the sweep can take the minimizer-bin path (17 <= k <= 63, k != 32) and every group has at least T genomes."""
    import torch
    import torch.distributed as tdist
    from . import dist as kd
    k_values = [str(k) for k in (k_values or DEFAULT_K_VALUES)]
    own_eng = None
    if not tdist.is_initialized():
        kd.init_from_env()
    rank, world = tdist.get_rank(), tdist.get_world_size()
    if adapter is None:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        own_eng = Engine(local)
        adapter = kd.CudaAdapter(own_eng, torch.device("cuda", local))
    report = {"mode": "fused-distributed", "world": world, "work_root": work_root, "num_datasets": num_datasets, "k_values": k_values, "stages": []}
    t_start = time.time()
    names = {n: genomes_of(work_root, n) for n in range(1, num_datasets + 1)}
    shardable = all(17 <= int(k) <= 63 and int(k) != 32 for k in k_values) and hasattr(getattr(adapter, "eng", None), "team_alloc")
    T = int(team) if team else kd.team_shape(num_datasets, world)
    if T > 1 and not (shardable and world % T == 0 and T <= 8 and min(len(v) for v in names.values()) >= T):
        if team:
            raise ValueError(f"team={T}: needs a divisor of the {world} ranks (<= 8), groups of at least {T} genomes and 17 <= k <= 63, k != 32")
        T = 1
    n_teams, team_idx, member = world // T, rank // T, rank % T
    mine = kd.groups_of_rank(num_datasets, team_idx, n_teams)
    report["team_size"] = T
    ts, team_pg, group_syms = None, None, {}
    if T > 1:
        for t in range(n_teams):                # every rank creates every team's process group, in the same order
            pg = tdist.new_group(ranks=list(range(t * T, (t + 1) * T)))
            if t == team_idx:
                team_pg = pg
        ts = kd.TeamSharder(adapter.eng, T, member, group=team_pg)
    slices = {n: kd.genome_slices(len(names[n]), T) for n in names}
    packed: Dict[int, object] = {}
    ex = None             # ONE exchanger at a time: a khb_ctx holds one peer exchange (one set of mapped regions, one key width)
    reader = None
    ctrl = torch.device("cpu") if tdist.get_backend() != "nccl" else adapter.new_tensor(0).device
    try:
        if rank == 0:
            write_complex_ops(work_root, k_values, num_datasets)
        zero = np.zeros(tables.HIST_ROWS + 1, dtype=np.uint64)
        if mine:
            # a rank reads (inflates, packs) only what it holds: its groups, or its slice of its team's groups
            reader = ingest.GroupReader({n: [os.path.join(work_root, p_genome(n, g)) for g in names[n][slices[n][member][0]:slices[n][member][1]]]
                                         for n in mine}, mine)
        for k in k_values:
            ki = int(k)
            if ex is not None and kd.key_words(ki) != kd.key_words(ex.k):
                ex.close()                              # collective: the key width changes (k crosses 32), in either direction
                ex = None
            if ex is None:
                ex = kd.AcrossExchanger(adapter, ki, num_datasets, nbins=tables.HIST_ROWS, mode=exchange)
            ex.set_k(ki)
            adapter.reset()
            ex.begin()
            within = np.zeros((num_datasets, tables.HIST_ROWS + 2), dtype=np.int64)     # last column: distinct k-mers of the group
            for num in mine:
                if num not in packed:
                    packed[num] = adapter.pack_group(reader.get(num))
                    if ts:      # symbols of the whole group: the same number on every member (sizes the team's bins and tables)
                        nsym = torch.tensor([int(packed[num].info()["n_symbols"])], dtype=torch.int64, device=ctrl)
                        tdist.all_reduce(nsym, group=team_pg)
                        group_syms[num] = int(nsym.item())
                if ts:
                    # this member's rows: the k-mers of the bins it owns; the all-reduce below adds the members' rows up
                    hist, st = ts.run_group(packed[num], ki, len(names[num]), [hi - lo for lo, hi in slices[num]], group_syms[num], nbins=tables.HIST_ROWS)
                else:
                    hist, st = adapter.group_from_packed(packed[num], ki, tables.HIST_ROWS)
                ex.after_group()
                within[num - 1, :-1] = hist.astype(np.int64)
                within[num - 1, -1] = int(st["distinct"])
                report["stages"].append({"k": ki, "group": num, "rank": rank, **{a: b for a, b in st.items() if isinstance(b, (int, float))}})
            across, info = ex.finish()
            w = torch.from_numpy(within).to(ctrl)
            tdist.all_reduce(w, op=tdist.ReduceOp.SUM)
            within = w.cpu().numpy()
            d_all = torch.tensor([int(info.get("local_distinct", info.get("distinct", 0)))], dtype=torch.int64, device=ctrl)
            if world > 1:
                tdist.all_reduce(d_all, op=tdist.ReduceOp.SUM)      # every k-mer has exactly one owner
            if rank == 0:
                for num in range(1, num_datasets + 1):
                    hist = within[num - 1, :-1].astype(np.uint64)
                    distinct = int(within[num - 1, -1])
                    tables.write_histogram_file(os.path.join(work_root, p_step4(k, num)), hist)
                    if stubs:
                        one = zero.copy()
                        for g in names[num]:
                            kmcdb.write_db(os.path.join(work_root, p_step1(k, num, g)), ki, None, None, zero, cli.KMC_DEFAULT_CS)
                            kmcdb.write_db(os.path.join(work_root, p_step2(k, num, g)), ki, None, None, zero, cli.KMC_DEFAULT_CS)
                        kmcdb.write_db(os.path.join(work_root, p_step3(k, num)), ki, None, None, hist, COUNTER_MAX, distinct)
                        one[1] = distinct
                        kmcdb.write_db(os.path.join(work_root, p_step6(k, num)), ki, None, None, one, COUNTER_MAX, distinct)
                across = np.asarray(across, dtype=np.uint64)
                tables.write_histogram_file(os.path.join(work_root, p_step8(k)), across)
                if stubs:
                    kmcdb.write_db(os.path.join(work_root, p_step7(k)), ki, None, None, across, COUNTER_MAX, int(d_all.item()))
                report["stages"].append({"k": ki, "group": "across", "distinct": int(d_all.item()), "exchange": info.get("exchange", "local")})
        if rank == 0:
            build_tables(work_root, k_values, num_datasets)
        if ex is not None:
            ex.close()
        if ts is not None:
            ts.close()
        tdist.barrier()
    finally:
        if reader is not None:
            reader.close()
        for pk in packed.values():
            pk.free()
        if own_eng is not None:
            own_eng.close()
    report["seconds"] = time.time() - t_start
    if report_path and rank == 0:
        with open(report_path, "w") as fd:
            json.dump(report, fd, indent=1)
    return report


# ---- rule-by-rule mode (mini scheduler) --------------------------------------------------------------
def _rule_jobs(work_root: str, k_values: Sequence[str], num_datasets: int):
    """(rule name, outputs, shell string) for every rule instance, in a valid topological order.  The
    shell strings are the reference's own (exp_type_1.smk:163,173,182,191,241,250,259)."""
    jobs = []
    for k in k_values:
        for num in range(1, num_datasets + 1):
            for g in genomes_of(work_root, num):
                jobs.append(("build_kmc_database_on_genome", [p_step1(k, num, g) + e for e in (".kmc_pre", ".kmc_suf")],
                             f"kmc -fm -m64 -k{k} -ci1 {p_genome(num, g)} {p_step1(k, num, g)} tmp/"))
                jobs.append(("transform_genome_to_set", [p_step2(k, num, g) + e for e in (".kmc_pre", ".kmc_suf")],
                             f"kmc_tools transform {p_step1(k, num, g)} set_counts 1 {p_step2(k, num, g)}"))
            jobs.append(("within_group_union", [p_step3(k, num) + e for e in (".kmc_pre", ".kmc_suf")],
                         f"kmc_tools complex {p_ops_within(k, num)}"))
            jobs.append(("within_group_union_histogram", [p_step4(k, num)],
                         f"kmc_tools transform {p_step3(k, num)} histogram {p_step4(k, num)}"))
            jobs.append(("build_group_kmer_set", [p_step6(k, num) + e for e in (".kmc_pre", ".kmc_suf")],
                         f"kmc_tools transform {p_step3(k, num)} set_counts 1 {p_step6(k, num)}"))
        jobs.append(("across_group_union", [p_step7(k) + e for e in (".kmc_pre", ".kmc_suf")],
                     f"kmc_tools complex {p_ops_across(k)}"))
        jobs.append(("across_group_union_histogram", [p_step8(k)],
                     f"kmc_tools transform {p_step7(k)} histogram {p_step8(k)}"))
    return jobs


def _fused_rule_jobs(work_root: str, k_values: Sequence[str], num_datasets: int):
    """The rules of khoice_b200/workflow/exp_type_1.smk in KHB_MODE=fused, as (rule, outputs, shell): the same ten names and
    output patterns as the reference; `khb group` / `khb across` carry the arithmetic, `khb stub` the per-genome placeholders,
    the three transform rules are the reference's own strings."""
    jobs = []
    pair = lambda p: [p + e for e in (".kmc_pre", ".kmc_suf")]
    for k in k_values:
        for num in range(1, num_datasets + 1):
            names = genomes_of(work_root, num)
            for g in names:
                jobs.append(("build_kmc_database_on_genome", pair(p_step1(k, num, g)), f"khb stub --k {k} {p_step1(k, num, g)}"))
                jobs.append(("transform_genome_to_set", pair(p_step2(k, num, g)), f"khb stub --k {k} {p_step2(k, num, g)}"))
            ins = " ".join(p_genome(num, g) for g in names)
            jobs.append(("within_group_union", pair(p_step3(k, num)), f"khb group --k {k} --table {p_step3(k, num)} {ins}"))
            jobs.append(("within_group_union_histogram", [p_step4(k, num)],
                         f"kmc_tools transform {p_step3(k, num)} histogram {p_step4(k, num)}"))
            jobs.append(("build_group_kmer_set", pair(p_step6(k, num)),
                         f"kmc_tools transform {p_step3(k, num)} set_counts 1 {p_step6(k, num)}"))
        sets = " ".join(p_step6(k, n) for n in range(1, num_datasets + 1))
        jobs.append(("across_group_union", pair(p_step7(k)), f"khb across --k {k} --table {p_step7(k)} {sets}"))
        jobs.append(("across_group_union_histogram", [p_step8(k)], f"kmc_tools transform {p_step7(k)} histogram {p_step8(k)}"))
    return jobs


def run_rules(work_root: str, num_datasets: int, k_values: Optional[Sequence] = None, subprocess_mode: bool = False,
              engine: Optional[Engine] = None, fused_rules: bool = False) -> Dict:
    """Run every exp-1 rule instance separately through the kmc / kmc_tools shims (or, with fused_rules, the two
    rule-granular fused commands of khoice_b200/workflow/exp_type_1.smk)."""
    k_values = [str(k) for k in (k_values or DEFAULT_K_VALUES)]
    write_complex_ops(work_root, k_values, num_datasets)
    cwd = os.getcwd()
    ran, skipped = 0, 0
    env = dict(os.environ, PATH=BIN_DIR + os.pathsep + os.environ.get("PATH", ""))
    own = None
    if not subprocess_mode:
        own = engine or Engine(int(os.environ.get("KHB_DEVICE", "0")))
        cli.set_engine(own)
    try:
        os.chdir(work_root)  # `workdir:` of the Snakefile (Snakefile:45)
        for rule, outputs, shell in (_fused_rule_jobs if fused_rules else _rule_jobs)(".", k_values, num_datasets):
            if all(os.path.exists(o) for o in outputs):
                skipped += 1
                continue
            for o in outputs:
                os.makedirs(os.path.dirname(o) or ".", exist_ok=True)
            if subprocess_mode:
                rc = subprocess.run(["/bin/sh", "-c", shell], env=env).returncode
            else:
                argv = shlex.split(shell)
                rc = cli.main(argv)
            if rc != 0:
                for o in outputs:
                    if os.path.exists(o):
                        os.remove(o)
                raise RuntimeError(f"rule {rule} failed (exit {rc}): {shell}")
            ran += 1
    finally:
        os.chdir(cwd)
        if not subprocess_mode:
            cli.set_engine(None)
            if engine is None and own is not None:
                own.close()
    build_tables(work_root, k_values, num_datasets)
    return {"mode": "rules-subprocess" if subprocess_mode else "rules", "jobs_run": ran, "jobs_skipped": skipped}


def main(argv: Optional[List[str]] = None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="khoice experiment type 1 on the B200 engine")
    ap.add_argument("--work-root", required=True, help="WORK_ROOT of the reference config (config/config.yaml)")
    ap.add_argument("--num-datasets", type=int, required=True, help="NUM_DATASETS")
    ap.add_argument("--k-values", default=None, help="comma separated K_VALUES (default: the reference list, Snakefile:36)")
    ap.add_argument("--mode", choices=["fused", "fused-dist", "rules", "rules-subprocess"], default="fused",
                    help="fused-dist: under `torchrun --nproc-per-node N`, one rank per GPU")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer", help="fused-dist: across-group exchange route")
    ap.add_argument("--team", type=int, default=0, help="fused-dist: GPUs per group (1: whole groups; 0: the smallest team size that deals the groups evenly)")
    ap.add_argument("--report", default=None)
    a = ap.parse_args(argv)
    ks = a.k_values.split(",") if a.k_values else None
    if a.mode == "fused":
        rep = run_fused(a.work_root, a.num_datasets, ks, report_path=a.report)
    elif a.mode == "fused-dist":
        rep = run_fused_distributed(a.work_root, a.num_datasets, ks, exchange=a.exchange, report_path=a.report, team=a.team)
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
    else:
        rep = run_rules(a.work_root, a.num_datasets, ks, subprocess_mode=a.mode == "rules-subprocess")
    print(json.dumps({k: v for k, v in rep.items() if k != "stages"}))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
