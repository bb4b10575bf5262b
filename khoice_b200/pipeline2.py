"""Host driver of experiment type 2 (pivot analysis) on the B200 engine: the rule chain of
/root/reference/workflow/rules/exp_type_2.smk with the same rule names, inputs and outputs
(SURVEY.md section 8f, row N1).

For every dataset n the reference compares the k-mer set of one held-out pivot genome
(``input_type_2/pivot/dataset_{n}/pivot_{n}.fna.gz``) with

* the union of the other genomes of its own dataset (``input_type_2/rest_of_set/dataset_{n}/*.fna.gz``) --
  ``kmc_tools simple pivot union intersect -ocsum`` / ``kmers_subtract`` + histograms ->
  ``within_dataset_analysis_type_2/within_dataset_analysis.csv`` (exp_type_2.smk:297-436), and
* the union of the rest-of-set unions of all OTHER datasets -> ``across_dataset_analysis_type_2/
  across_dataset_analysis.csv`` (exp_type_2.smk:438-553).

Two execution modes, as for experiment type 1 (khoice_b200/pipeline.py):

* ``fused``  -- per (k, dataset) ONE sort of all windows of rest-of-set + pivot with the genome id as payload
  (``Engine.pivot_group_from_packed``: the pivot is the last genome, so a run of equal k-mers that ends with the
  pivot's id is a pivot k-mer and its pair count is the -ocsum counter), then ONE sort per k of all retained
  unions and pivot sets (``Engine.pivot_across``).  Intermediate databases are written as header-only stubs.
* ``rules``  -- every rule instance through the kmc / kmc_tools shims (khoice_b200/cli.py), exchanging real
  intermediate databases.
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
from .pipeline import BIN_DIR, DEFAULT_K_VALUES

OPS = ("subtract", "intersect")  # order of get_{within,across}_group_histogram_files (exp_type_2.smk:153-169)


# ---- layout helpers (paths exactly as in the rules) ---------------------------------------------------
def rest_genomes_of(work_root: str, num: int) -> List[str]:
    """Genome names of input_type_2/rest_of_set/dataset_{num} (exp_type_2.smk:63-66), sorted."""
    d = os.path.join(work_root, "input_type_2", "rest_of_set", f"dataset_{num}")
    return sorted(f.split(".fna.gz")[0] for f in os.listdir(d) if f.endswith(".fna.gz"))


def p_rest(num, g): return f"input_type_2/rest_of_set/dataset_{num}/{g}.fna.gz"
def p_pivot(num): return f"input_type_2/pivot/dataset_{num}/pivot_{num}.fna.gz"
def p_s1_rest(k, num, g): return f"step_1_type_2/rest_of_set/k_{k}/dataset_{num}/{g}"
def p_s1_pivot(k, num): return f"step_1_type_2/pivot/k_{k}/dataset_{num}/pivot_{num}"
def p_set_rest(k, num, g): return f"genome_sets_type_2/rest_of_set/k_{k}/dataset_{num}/{g}.transformed"
def p_set_pivot(k, num): return f"genome_sets_type_2/pivot/k_{k}/dataset_{num}/pivot_{num}.transformed"
def p_union(k, num): return f"within_databases_type_2/rest_of_set/k_{k}/dataset_{num}/dataset_{num}.transformed.combined"
def p_union_set(k, num): return p_union(k, num) + ".transformed"
def p_within(k, num, op): return f"within_dataset_results_type_2/k_{k}/dataset_{num}/{op}/dataset_{num}_pivot_{op}_group"
def p_across_db(k, num): return f"across_databases_type_2/k_{k}/pivot_{num}/all_datasets_pivot_{num}.transformed.combined.transformed.combined"
def p_across(k, num, op): return f"across_dataset_results_type_2/k_{k}/dataset_{num}/{op}/dataset_{num}_pivot_{op}_group"
def p_ops_within(k, num): return f"complex_ops_type_2/within_groups/k_{k}/dataset_{num}/within_dataset_{num}.txt"
def p_ops_across(k, num): return f"complex_ops_type_2/across_groups/k_{k}/pivot_{num}/across_datasets_pivot_{num}.txt"
P_WITHIN_CSV = "within_dataset_analysis_type_2/within_dataset_analysis.csv"
P_ACROSS_CSV = "across_dataset_analysis_type_2/across_dataset_analysis.csv"


def prepare_inputs(work_root: str, database_root: str, trial: int, num_datasets: int) -> None:
    """The copy step of the parse-time block (exp_type_2.smk:31-48): rest-of-set genomes and the pivot of every
    dataset from ``{database_root}/trial_{trial}/exp0_{nonpivot,pivot}_genomes/`` into ``input_type_2/``."""
    if os.path.isdir(os.path.join(work_root, "input_type_2")):
        return
    for i in range(1, num_datasets + 1):
        src = os.path.join(database_root, f"trial_{trial}", "exp0_nonpivot_genomes", f"dataset_{i}")
        dst = os.path.join(work_root, "input_type_2", "rest_of_set", f"dataset_{i}")
        os.makedirs(dst, exist_ok=True)
        for f in os.listdir(src):
            shutil.copy(os.path.join(src, f), dst)
        pdst = os.path.join(work_root, "input_type_2", "pivot", f"dataset_{i}")
        os.makedirs(pdst, exist_ok=True)
        shutil.copy(os.path.join(database_root, f"trial_{trial}", "exp0_pivot_genomes", f"dataset_{i}", f"pivot_{i}.fna.gz"),
                    os.path.join(pdst, f"pivot_{i}.fna.gz"))


def write_complex_ops(work_root: str, k_values: Sequence[str], num_datasets: int) -> None:
    """The rest of the parse-time block (exp_type_2.smk:27-29, 50-113): tmp/ and the `kmc_tools complex` operation
    files -- per (k, dataset) the union of its rest-of-set genome sets, per (k, pivot) the union of the OTHER
    datasets' rest-of-set sets."""
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
            emit(p_ops_within(k, num), [p_set_rest(k, num, g) for g in rest_genomes_of(work_root, num)], p_union(k, num))
        for piv in range(1, num_datasets + 1):
            emit(p_ops_across(k, piv), [p_union_set(k, i) for i in range(1, num_datasets + 1) if i != piv], p_across_db(k, piv))


def histogram_files(scope: str, k_values: Sequence[str], num_datasets: int) -> List[str]:
    """get_within_group_histogram_files / get_across_group_histogram_files (exp_type_2.smk:153-169): dataset-major,
    then k, then subtract before intersect."""
    fn = p_within if scope == "within" else p_across
    return [fn(k, num, op) + ".hist.txt" for num in range(1, num_datasets + 1) for k in k_values for op in OPS]


def build_tables(work_root: str, k_values: Sequence[str], num_datasets: int) -> None:
    """Rules within_group_analysis_exp_type2 and across_group_analysis_exp_type2."""
    members = lambda n: tables.get_num_of_dataset_members_exp2(n, os.path.join(work_root, "input_type_2"))
    tables.within_group_analysis_exp_type2([os.path.join(work_root, f) for f in histogram_files("within", k_values, num_datasets)],
                                           os.path.join(work_root, P_WITHIN_CSV), num_datasets, members)
    tables.across_group_analysis_exp_type2([os.path.join(work_root, f) for f in histogram_files("across", k_values, num_datasets)],
                                           os.path.join(work_root, P_ACROSS_CSV), num_datasets)


def split_pivot_histogram(hist: np.ndarray):
    """Engine histogram (index = 1 + #sets containing the pivot k-mer) -> the two histograms the reference's rules
    read: kmers_subtract (all counters 1) and intersect -ocsum (counters >= 2)."""
    sub = np.zeros_like(hist)
    sub[1] = hist[1]
    inter = hist.copy()
    inter[:2] = 0
    return sub, inter


# ---- fused mode ----------------------------------------------------------------------------------------
def run_fused(work_root: str, num_datasets: int, k_values: Optional[Sequence] = None, engine: Optional[Engine] = None,
              stubs: bool = True, report_path: Optional[str] = None) -> Dict:
    """All of exp type 2 for ``work_root`` (``input_type_2/`` must exist, see prepare_inputs)."""
    k_values = [str(k) for k in (k_values or DEFAULT_K_VALUES)]
    own = engine is None
    eng = engine or Engine(int(os.environ.get("KHB_DEVICE", "0")))
    report = {"mode": "fused", "exp_type": 2, "work_root": work_root, "num_datasets": num_datasets, "k_values": k_values, "stages": []}
    t_start = time.time()
    packed: Dict[int, object] = {}
    zero = np.zeros(tables.HIST_ROWS + 1, dtype=np.uint64)

    def emit(prefix, ki, sub, inter):
        for op, h in zip(OPS, (sub, inter)):
            tables.write_histogram_file(os.path.join(work_root, prefix(op)) + ".hist.txt", h)
            if stubs:
                kmcdb.write_db(os.path.join(work_root, prefix(op)), ki, None, None, h, COUNTER_MAX, int(h.sum()))

    try:
        write_complex_ops(work_root, k_values, num_datasets)
        names = {n: rest_genomes_of(work_root, n) for n in range(1, num_datasets + 1)}
        # parallel inflate, one dataset ahead of the GPU; rest-of-set genomes first, the pivot LAST
        reader = ingest.GroupReader({n: [os.path.join(work_root, p_rest(n, g)) for g in names[n]] + [os.path.join(work_root, p_pivot(n))]
                                     for n in names}, sorted(names))
        for k in k_values:
            ki = int(k)
            eng.group_sets_reset()
            for num in range(1, num_datasets + 1):
                if num not in packed:  # inflate + pack once per dataset
                    texts = reader.get(num)
                    packed[num] = eng.pack_group(texts)
                    del texts
                hist, st = eng.pivot_group_from_packed(packed[num], ki, nbins=tables.HIST_ROWS, keep_sets=True)
                sub, inter = split_pivot_histogram(hist)
                emit(lambda op: p_within(k, num, op), ki, sub, inter)
                if stubs:
                    for g in names[num]:
                        kmcdb.write_db(os.path.join(work_root, p_s1_rest(k, num, g)), ki, None, None, zero, cli.KMC_DEFAULT_CS)
                        kmcdb.write_db(os.path.join(work_root, p_set_rest(k, num, g)), ki, None, None, zero, cli.KMC_DEFAULT_CS)
                    kmcdb.write_db(os.path.join(work_root, p_s1_pivot(k, num)), ki, None, None, zero, cli.KMC_DEFAULT_CS)
                    kmcdb.write_db(os.path.join(work_root, p_set_pivot(k, num)), ki, None, None, zero, cli.KMC_DEFAULT_CS)
                    kmcdb.write_db(os.path.join(work_root, p_union(k, num)), ki, None, None, zero, COUNTER_MAX, st["distinct"])
                    kmcdb.write_db(os.path.join(work_root, p_union_set(k, num)), ki, None, None, zero, COUNTER_MAX, st["distinct"])
                report["stages"].append({"k": ki, "dataset": num, **st})
            hists, st = eng.pivot_across(nbins=tables.HIST_ROWS)
            for num in range(1, num_datasets + 1):
                sub, inter = split_pivot_histogram(hists[num - 1])
                emit(lambda op: p_across(k, num, op), ki, sub, inter)
                if stubs:
                    kmcdb.write_db(os.path.join(work_root, p_across_db(k, num)), ki, None, None, zero, COUNTER_MAX)
            report["stages"].append({"k": ki, "dataset": "across", **st})
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


# ---- rule-by-rule mode ---------------------------------------------------------------------------------
def _rule_jobs(work_root: str, k_values: Sequence[str], num_datasets: int):
    """(rule name, outputs, shell string) for every rule instance, in a valid topological order.  The shell strings
    are the reference's own (exp_type_2.smk:306, 315, 331, 343, 355, 368, 381, 432, 450, 471, 484, 497)."""
    db = lambda p: [p + ".kmc_pre", p + ".kmc_suf"]
    jobs = []
    for k in k_values:
        for num in range(1, num_datasets + 1):
            for g in rest_genomes_of(work_root, num):
                jobs.append(("build_kmc_database_on_genome_exp_type_2", db(p_s1_rest(k, num, g)),
                             f"kmc -fm -m64 -k{k} -ci1 {p_rest(num, g)} {p_s1_rest(k, num, g)} tmp/"))
                jobs.append(("transform_genome_to_set_exp_type2", db(p_set_rest(k, num, g)),
                             f"kmc_tools transform {p_s1_rest(k, num, g)} set_counts 1 {p_set_rest(k, num, g)}"))
            jobs.append(("build_kmc_database_on_pivot_exp_type_2", db(p_s1_pivot(k, num)),
                         f"kmc -fm -m64 -k{k} -ci1 {p_pivot(num)} {p_s1_pivot(k, num)} tmp/"))
            jobs.append(("transform_pivot_to_set_exp_type2", db(p_set_pivot(k, num)),
                         f"kmc_tools transform {p_s1_pivot(k, num)} set_counts 1 {p_set_pivot(k, num)}"))
            jobs.append(("within_group_union_exp_type2", db(p_union(k, num)), f"kmc_tools complex {p_ops_within(k, num)}"))
            jobs.append(("pivot_intersect_within_group_exp_type2", db(p_within(k, num, "intersect")),
                         f"kmc_tools simple {p_set_pivot(k, num)} {p_union(k, num)} intersect {p_within(k, num, 'intersect')} -ocsum"))
            jobs.append(("pivot_subtract_within_group_exp_type2", db(p_within(k, num, "subtract")),
                         f"kmc_tools simple {p_set_pivot(k, num)} {p_union(k, num)} kmers_subtract {p_within(k, num, 'subtract')}"))
            for op in OPS:
                jobs.append(("within_group_histogram_exp_type2", [p_within(k, num, op) + ".hist.txt"],
                             f"kmc_tools transform {p_within(k, num, op)} histogram {p_within(k, num, op)}.hist.txt"))
            jobs.append(("transform_rest_of_set_to_single_counts", db(p_union_set(k, num)),
                         f"kmc_tools transform {p_union(k, num)} set_counts 1 {p_union_set(k, num)}"))
        for num in range(1, num_datasets + 1):
            jobs.append(("across_group_union_for_pivot_exp_type2", db(p_across_db(k, num)), f"kmc_tools complex {p_ops_across(k, num)}"))
            jobs.append(("pivot_intersect_across_group_exp_type2", db(p_across(k, num, "intersect")),
                         f"kmc_tools simple {p_set_pivot(k, num)} {p_across_db(k, num)} intersect {p_across(k, num, 'intersect')} -ocsum"))
            jobs.append(("pivot_subtract_across_group_exp_type2", db(p_across(k, num, "subtract")),
                         f"kmc_tools simple {p_set_pivot(k, num)} {p_across_db(k, num)} kmers_subtract {p_across(k, num, 'subtract')}"))
            for op in OPS:
                jobs.append(("across_group_histogram_exp_type2", [p_across(k, num, op) + ".hist.txt"],
                             f"kmc_tools transform {p_across(k, num, op)} histogram {p_across(k, num, op)}.hist.txt"))
    return jobs


def run_rules(work_root: str, num_datasets: int, k_values: Optional[Sequence] = None, subprocess_mode: bool = False,
              engine: Optional[Engine] = None) -> Dict:
    """Run every exp-2 rule instance separately through the kmc / kmc_tools shims."""
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
        os.chdir(work_root)
        for rule, outputs, shell in _rule_jobs(".", k_values, num_datasets):
            if all(os.path.exists(o) for o in outputs):
                skipped += 1
                continue
            for o in outputs:
                os.makedirs(os.path.dirname(o) or ".", exist_ok=True)
            rc = subprocess.run(["/bin/sh", "-c", shell], env=env).returncode if subprocess_mode else cli.main(shlex.split(shell))
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
    return {"mode": "rules-subprocess" if subprocess_mode else "rules", "exp_type": 2, "jobs_run": ran, "jobs_skipped": skipped}


def main(argv: Optional[List[str]] = None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="khoice experiment type 2 (pivot analysis) on the B200 engine")
    ap.add_argument("--work-root", required=True, help="WORK_ROOT of the reference config (config/config.yaml)")
    ap.add_argument("--num-datasets", type=int, required=True, help="NUM_DATASETS")
    ap.add_argument("--k-values", default=None, help="comma separated K_VALUES (default: the reference list, Snakefile:36)")
    ap.add_argument("--database-root", default=None, help="DB_ROOT: copy inputs from {root}/trial_{t}/exp0_* first")
    ap.add_argument("--trial", type=int, default=1)
    ap.add_argument("--mode", choices=["fused", "rules", "rules-subprocess"], default="fused")
    ap.add_argument("--report", default=None)
    a = ap.parse_args(argv)
    ks = a.k_values.split(",") if a.k_values else None
    if a.database_root:
        prepare_inputs(a.work_root, a.database_root, a.trial, a.num_datasets)
    if a.mode == "fused":
        rep = run_fused(a.work_root, a.num_datasets, ks, report_path=a.report)
    else:
        rep = run_rules(a.work_root, a.num_datasets, ks, subprocess_mode=a.mode == "rules-subprocess")
    print(json.dumps({k: v for k, v in rep.items() if k != "stages"}))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
