"""Host driver of experiment type 4 (feature-level confusion matrix) on the B200 engine: the rule chain of
/root/reference/workflow/rules/exp_type_4.smk with the same rule names, inputs and outputs (SURVEY.md section 8f,
row N4).

For every k the reference builds, per dataset d, the union U_d of its rest-of-set genomes' k-mer sets, per pivot p the
counted k-mers of ``input_type4/pivot/pivot_{p}.fna.gz``, intersects every pivot with every union (G x G
``kmc_tools simple ... intersect -ocsum`` runs), dumps everything as text and joins the dumps in a Python dictionary
(src/merge_lists.py) into ``accuracies_type_4/confusion_matrix/k_{k}_confusion_matrix[_with_unidentified].txt`` and
``accuracies_type_4/values/k_{k}_accuracy_values.csv``; the per-k tables are concatenated into
``accuracies_type_4/accuracy_values.csv``.

* ``fused``  -- per k: the unions come from the single-sort group path (``Engine.group_from_packed``), the pivots from
  sort + run-length count (``Engine.kmer_counts``), and ONE sort of all unions and all pivot k-mers with the set index
  as payload yields every pivot k-mer's group-membership bitmask (``Engine.group_membership``); the ordered floating
  point accumulation of the matrix stays on the host (khoice_b200/merge_lists.py: the order of the additions is part
  of the reference's result).  No text dumps, no G x G intersections.
* ``rules``  -- every rule instance through the kmc / kmc_tools shims, text dumps included; the last step runs
  ``khoice_b200.merge_lists`` with the reference's command line (the reference's own script works on these dumps too).
"""
from __future__ import annotations

import json
import os
import shlex
import shutil
import sys
import time
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import cli, ingest, merge_lists, tables
from .engine import Engine
from .pipeline import DEFAULT_K_VALUES


# ---- layout helpers (paths exactly as in the rules) ---------------------------------------------------
def rest_genomes_of(work_root: str, num: int) -> List[str]:
    d = os.path.join(work_root, "input_type4", "rest_of_set", f"dataset_{num}")
    return sorted(f.split(".fna.gz")[0] for f in os.listdir(d) if f.endswith(".fna.gz"))


def p_rest(num, g): return f"input_type4/rest_of_set/dataset_{num}/{g}.fna.gz"
def p_pivot(num): return f"input_type4/pivot/pivot_{num}.fna.gz"
def p_s1_rest(k, num, g): return f"step_1_type_4/rest_of_set/k_{k}/dataset_{num}/{g}"
def p_s1_pivot(k, num): return f"step_1_type_4/pivot/k_{k}/dataset_{num}/pivot_{num}"
def p_set_rest(k, num, g): return f"genome_sets_type_4/rest_of_set/k_{k}/dataset_{num}/{g}.transformed"
def p_union(k, num): return f"unions_type_4/rest_of_set/k_{k}/dataset_{num}/dataset_{num}.transformed.combined"
def p_union_hist(k, num): return f"unions_type_4/rest_of_set/k_{k}/dataset_{num}/dataset_{num}.hist.txt"
def p_union_set(k, num): return f"genome_sets_type_4/unions_type_4/k_{k}/dataset_{num}/dataset_{num}.transformed.combined.transformed"
def p_inter(k, piv, num): return f"intersection_results_type_4/k_{k}/pivot_{piv}/pivot_{piv}_intersect_dataset_{num}"
def p_dump_pivot(k, piv): return f"text_dump_type_4/k_{k}/pivot/pivot_{piv}.txt"
def p_dump_inter(k, piv, num): return f"text_dump_type_4/k_{k}/intersection/pivot_{piv}/pivot_{piv}_intersect_dataset_{num}.txt"
def p_ops(k, num): return f"complex_ops_type_4/k_{k}/dataset_{num}/ops_{num}.txt"
def p_values(k): return f"accuracies_type_4/values/k_{k}_accuracy_values.csv"
def p_matrix(k): return f"accuracies_type_4/confusion_matrix/k_{k}_confusion_matrix.txt"
P_FINAL = "accuracies_type_4/accuracy_values.csv"


def prepare_inputs(work_root: str, database_root: str, trial: int, num_datasets: int, out_pivot: bool = True) -> None:
    """The copy step of the parse-time block (exp_type_4.smk:31-52)."""
    if os.path.isdir(os.path.join(work_root, "input_type4")):
        return
    os.makedirs(os.path.join(work_root, "input_type4", "pivot"), exist_ok=True)
    for i in range(1, num_datasets + 1):
        src = os.path.join(database_root, f"trial_{trial}", "exp0_nonpivot_genomes", f"dataset_{i}")
        dst = os.path.join(work_root, "input_type4", "rest_of_set", f"dataset_{i}")
        os.makedirs(dst, exist_ok=True)
        for f in os.listdir(src):
            shutil.copy(os.path.join(src, f), dst)
        pivot = os.path.join(database_root, f"trial_{trial}", "exp0_pivot_genomes", f"dataset_{i}", f"pivot_{i}.fna.gz")
        shutil.copy(pivot, os.path.join(work_root, "input_type4", "pivot", f"pivot_{i}.fna.gz"))
        if not out_pivot:
            shutil.copy(pivot, dst)


def write_parse_time_files(work_root: str, k_values: Sequence[str], num_datasets: int) -> None:
    """tmp/, the `kmc_tools complex` operation files and the two file lists per k (exp_type_4.smk:27-29, 54-103)."""
    os.makedirs(os.path.join(work_root, "tmp"), exist_ok=True)
    base_dir = os.path.abspath(work_root)
    for k in k_values:
        for num in range(1, num_datasets + 1):
            full = os.path.join(work_root, p_ops(k, num))
            os.makedirs(os.path.dirname(full), exist_ok=True)
            inputs = [p_set_rest(k, num, g) for g in rest_genomes_of(work_root, num)]
            lines = ["INPUT:"] + [f"set{i + 1} = {p}" for i, p in enumerate(inputs)]
            lines += ["OUTPUT:", f"{p_union(k, num)} = (" + " + ".join(f"set{i + 1}" for i in range(len(inputs))) + ")",
                      "OUTPUT_PARAMS:", "-cs5000"]
            with open(full, "w") as fd:
                fd.write("\n".join(lines) + "\n")
        d = os.path.join(work_root, "filelists_type_4", f"k_{k}")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "pivots_filelist.txt"), "w") as fd:
            for piv in range(1, num_datasets + 1):
                fd.write(f"{base_dir}/{p_dump_pivot(k, piv)}\n")
        with open(os.path.join(d, "intersections_filelist.txt"), "w") as fd:
            for piv in range(1, num_datasets + 1):
                for num in range(1, num_datasets + 1):
                    fd.write(f"{base_dir}/{p_dump_inter(k, piv, num)}\n")


def concatenate_accuracies(work_root: str) -> None:
    """Rule concatenate_accuracies_exp_type4: `cat accuracies_type_4/values/*.csv` (shell glob order = sorted names)."""
    d = os.path.join(work_root, "accuracies_type_4", "values")
    with open(os.path.join(work_root, P_FINAL), "w") as out:
        for f in sorted(x for x in os.listdir(d) if x.endswith(".csv")):
            with open(os.path.join(d, f)) as fd:
                out.write(fd.read())


# ---- fused mode ----------------------------------------------------------------------------------------
def run_fused(work_root: str, num_datasets: int, k_values: Optional[Sequence] = None, engine: Optional[Engine] = None,
              report_path: Optional[str] = None) -> Dict:
    """All of exp type 4 (feature level) for ``work_root`` (``input_type4/`` must exist, see prepare_inputs)."""
    k_values = [str(k) for k in (k_values or DEFAULT_K_VALUES)]
    own = engine is None
    eng = engine or Engine(int(os.environ.get("KHB_DEVICE", "0")))
    report = {"mode": "fused", "exp_type": 4, "work_root": work_root, "num_datasets": num_datasets, "k_values": k_values, "stages": []}
    t_start = time.time()
    packed: Dict[int, object] = {}
    pivot_text: Dict[int, bytes] = {}
    try:
        write_parse_time_files(work_root, k_values, num_datasets)
        names = {n: rest_genomes_of(work_root, n) for n in range(1, num_datasets + 1)}
        reader = ingest.GroupReader({n: [os.path.join(work_root, p_rest(n, g)) for g in names[n]] for n in names}, sorted(names))
        for k in k_values:
            ki = int(k)
            eng.group_sets_reset()
            group_off = [0]
            for num in range(1, num_datasets + 1):
                if num not in packed:
                    packed[num] = eng.pack_group(reader.get(num))
                hist, st = eng.group_from_packed(packed[num], ki, nbins=tables.HIST_ROWS, keep_set=True)
                tables.write_histogram_file(os.path.join(work_root, p_union_hist(k, num)), hist)   # rule union_histogram_exp_type_4
                group_off.append(eng.group_sets_info()["n_keys"])
                report["stages"].append({"k": ki, "dataset": num, **st})
            bufs, counts, sizes = [], [], []
            try:
                for piv in range(1, num_datasets + 1):
                    if piv not in pivot_text:
                        pivot_text[piv] = ingest.read_fasta(os.path.join(work_root, p_pivot(piv)))
                    buf, cnt, n = eng.kmer_counts(pivot_text[piv], ki, cs=cli.KMC_DEFAULT_CS)
                    bufs.append(buf); counts.append(cnt); sizes.append(n)
                masks = eng.group_membership(group_off, bufs, sizes, ki)
            finally:
                for b in bufs:
                    b.free()
            per_pivot, at = [], 0
            for n in sizes:
                per_pivot.append(masks[at:at + n])
                at += n
            matrix, matrix_u = merge_lists.confusion_from_masks(counts, per_pivot, num_datasets)
            merge_lists.write_outputs(os.path.join(work_root, "accuracies_type_4") + "/", k, matrix, matrix_u, num_datasets)
            report["stages"].append({"k": ki, "dataset": "pivots", "pivot_kmers": int(sum(sizes))})
        concatenate_accuracies(work_root)
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
    """(rule name, outputs, shell string) per rule instance in topological order; the shell strings are the reference's
    (exp_type_4.smk:143, 152, 167, 185, 197, 210, 229, 242, 254, 267, 284) minus the `rm` clean-ups, with
    `python3 -m khoice_b200.merge_lists` in place of `python3 {repo_dir}/src/merge_lists.py` (same arguments)."""
    db = lambda p: [p + ".kmc_pre", p + ".kmc_suf"]
    base_dir = os.path.abspath(work_root)
    jobs = []
    for k in k_values:
        for num in range(1, num_datasets + 1):
            for g in rest_genomes_of(work_root, num):
                jobs.append(("build_kmc_database_on_genome_exp_type_4", db(p_s1_rest(k, num, g)),
                             f"kmc -fm -m64 -k{k} -ci1 {p_rest(num, g)} {p_s1_rest(k, num, g)} tmp/"))
                jobs.append(("transform_genome_to_set_exp_type_4", db(p_set_rest(k, num, g)),
                             f"kmc_tools transform {p_s1_rest(k, num, g)} set_counts 1 {p_set_rest(k, num, g)}"))
            jobs.append(("build_kmc_database_on_pivot_exp_type_4", db(p_s1_pivot(k, num)),
                         f"kmc -fm -m64 -k{k} -ci1 {p_pivot(num)} {p_s1_pivot(k, num)} tmp/"))
            jobs.append(("rest_of_set_union_exp_type_4", db(p_union(k, num)), f"kmc_tools complex {p_ops(k, num)}"))
            jobs.append(("union_histogram_exp_type_4", [p_union_hist(k, num)],
                         f"kmc_tools transform {p_union(k, num)} histogram {p_union_hist(k, num)}"))
            jobs.append(("transform_union_to_set_exp_type_4", db(p_union_set(k, num)),
                         f"kmc_tools transform {p_union(k, num)} set_counts 1 {p_union_set(k, num)}"))
        for piv in range(1, num_datasets + 1):
            jobs.append(("pivot_text_dump_exp_type_4", [p_dump_pivot(k, piv)],
                         f"kmc_tools transform {p_s1_pivot(k, piv)} dump -s {p_dump_pivot(k, piv)}"))
            for num in range(1, num_datasets + 1):
                jobs.append(("pivot_intersect_exp_type_4", db(p_inter(k, piv, num)),
                             f"kmc_tools simple {p_union_set(k, num)} {p_s1_pivot(k, piv)} intersect {p_inter(k, piv, num)} -ocsum"))
                jobs.append(("intersection_text_dump_exp_type_4", [p_dump_inter(k, piv, num)],
                             f"kmc_tools transform {p_inter(k, piv, num)} dump -s {p_dump_inter(k, piv, num)}"))
        jobs.append(("run_merge_list_exp_type_4", [p_values(k), p_matrix(k)],
                     f"{shlex.quote(sys.executable)} -m khoice_b200.merge_lists -p {base_dir}/filelists_type_4/k_{k}/pivots_filelist.txt "
                     f"-i {base_dir}/filelists_type_4/k_{k}/intersections_filelist.txt -o {base_dir}/accuracies_type_4/ -n {num_datasets} -k {k}"))
    return jobs


def run_rules(work_root: str, num_datasets: int, k_values: Optional[Sequence] = None, engine: Optional[Engine] = None) -> Dict:
    """Run every exp-4 rule instance separately through the kmc / kmc_tools shims (in-process)."""
    k_values = [str(k) for k in (k_values or DEFAULT_K_VALUES)]
    write_parse_time_files(work_root, k_values, num_datasets)
    cwd = os.getcwd()
    ran, skipped = 0, 0
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
            argv = shlex.split(shell)
            if argv[0] in ("kmc", "kmc_tools"):
                rc = cli.main(argv)
            else:  # the merge step: `python -m khoice_b200.merge_lists ...`
                rc = merge_lists.main(argv[3:])
            if rc != 0:
                for o in outputs:
                    if os.path.exists(o):
                        os.remove(o)
                raise RuntimeError(f"rule {rule} failed (exit {rc}): {shell}")
            ran += 1
    finally:
        os.chdir(cwd)
        cli.set_engine(None)
        if engine is None:
            own.close()
    concatenate_accuracies(work_root)
    return {"mode": "rules", "exp_type": 4, "jobs_run": ran, "jobs_skipped": skipped}


def main(argv: Optional[List[str]] = None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="khoice experiment type 4 (feature-level confusion matrix) on the B200 engine")
    ap.add_argument("--work-root", required=True)
    ap.add_argument("--num-datasets", type=int, required=True)
    ap.add_argument("--k-values", default=None)
    ap.add_argument("--database-root", default=None, help="DB_ROOT: copy inputs from {root}/trial_{t}/exp0_* first")
    ap.add_argument("--trial", type=int, default=1)
    ap.add_argument("--in-pivot", action="store_true", help="OUT_PIVOT: False -- the pivot also joins its rest of set")
    ap.add_argument("--mode", choices=["fused", "rules"], default="fused")
    ap.add_argument("--report", default=None)
    a = ap.parse_args(argv)
    ks = a.k_values.split(",") if a.k_values else None
    if a.database_root:
        prepare_inputs(a.work_root, a.database_root, a.trial, a.num_datasets, out_pivot=not a.in_pivot)
    rep = run_fused(a.work_root, a.num_datasets, ks, report_path=a.report) if a.mode == "fused" else run_rules(a.work_root, a.num_datasets, ks)
    print(json.dumps({k: v for k, v in rep.items() if k != "stages"}))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
