"""Host driver of experiment type 6 (confusion matrix from simulated reads) on the B200 engine: the rule chain of
/root/reference/workflow/rules/exp_type_6.smk with the same rule names, inputs and outputs.

Experiment 6 is experiment 4 with the pivots replaced by two sets of simulated reads per out-pivot genome
(``exp6_input/pivot_reads_subset/{illumina,ont}/pivot_{n}.fa``) and the confusion matrix built at the READ level
(rule run_merge_list_exp6 passes ``-r``, exp_type_6.smk:327-346): every read votes with its k-mers and counts once.  The
read simulators themselves (prepare_data.smk) are outside the path; ``khoice_b200.synth.make_reads`` writes stand-ins.

* ``fused``  -- per k: the unions of the rest-of-set genomes from the group path (``Engine.group_from_packed``), per read
  type the counted k-mers of every reads file (``Engine.kmer_counts``), ONE sort for all membership bit masks
  (``Engine.group_membership``), then per pivot ``Engine.read_votes`` (K2 over the reads, a sorted lookup of every window's
  k-mer, and one thread per (read, dataset) adding the votes in window order in IEEE doubles).  The argmax with the
  reference's ``random.choice`` tie-breaking stays on the host (khoice_b200/merge_lists.py).  ``level="feature"`` gives the
  feature-level matrix instead (the reference's rule with its ``-r`` line commented out, as its comment suggests).
* ``rules``  -- every rule instance through the kmc / kmc_tools shims, text dumps included; the last step runs
  ``khoice_b200.merge_lists`` with the reference's command line (the reference's own script works on these dumps too).
"""
from __future__ import annotations

import json
import os
import random
import shlex
import shutil
import sys
import time
from typing import Dict, List, Optional, Sequence

from . import cli, ingest, merge_lists, tables
from .engine import Engine
from .pipeline import DEFAULT_K_VALUES

READ_TYPES = ("illumina", "ont")
HEADER = "k,pivotnum,TP,TN,FP,FN,TP-U,TN-U,FP-U,FN-U\n"


# ---- layout helpers (paths exactly as in the rules) ---------------------------------------------------
def rest_genomes_of(work_root: str, num: int) -> List[str]:
    d = os.path.join(work_root, "exp6_input", "rest_of_set", f"dataset_{num}")
    return sorted(f.split(".fna.gz")[0] for f in os.listdir(d) if f.endswith(".fna.gz"))


def p_rest(num, g): return f"exp6_input/rest_of_set/dataset_{num}/{g}.fna.gz"
def p_reads(rt, num): return f"exp6_input/pivot_reads_subset/{rt}/pivot_{num}.fa"
def p_s1_rest(k, num, g): return f"exp6_intermediate/step_1/rest_of_set/k_{k}/dataset_{num}/{g}"
def p_s1_pivot(k, rt, num): return f"exp6_intermediate/step_1/pivot/k_{k}/{rt}/pivot_{num}"
def p_set_rest(k, num, g): return f"exp6_genome_sets/rest_of_set/k_{k}/dataset_{num}/{g}.transformed"
def p_union(k, num): return f"exp6_unions/rest_of_set/k_{k}/dataset_{num}/dataset_{num}.transformed.combined"
def p_union_hist(k, num): return f"exp6_unions/rest_of_set/k_{k}/dataset_{num}/dataset_{num}.hist.txt"
def p_union_set(k, num): return f"exp6_genome_sets/unions/k_{k}/dataset_{num}/dataset_{num}.transformed.combined.transformed"
def p_inter(k, rt, piv, num): return f"exp6_intersection_results/k_{k}/{rt}/pivot_{piv}/pivot_{piv}_intersect_dataset_{num}"
def p_dump_pivot(k, rt, piv): return f"exp6_text_dump/k_{k}/{rt}/pivot/pivot_{piv}.txt"
def p_dump_inter(k, rt, piv, num): return f"exp6_text_dump/k_{k}/{rt}/intersection/pivot_{piv}/pivot_{piv}_intersect_dataset_{num}.txt"
def p_ops(k, num): return f"exp6_complex_ops/k_{k}/dataset_{num}/ops_{num}.txt"
def p_values(rt, k): return f"exp6_accuracies/{rt}/values/k_{k}_accuracy_values.csv"
def p_matrix(rt, k): return f"exp6_accuracies/{rt}/confusion_matrix/k_{k}_confusion_matrix.txt"
def p_final(trial, rt): return f"exp6_accuracies/trial_{trial}_{'short' if rt == 'illumina' else 'long'}_acc.csv"


def prepare_inputs(work_root: str, database_root: str, trial: int, num_datasets: int) -> None:
    """The copy step of the parse-time block (exp_type_6.smk:31-57)."""
    if os.path.isdir(os.path.join(work_root, "exp6_input")):
        return
    base = os.path.join(database_root, f"trial_{trial}")
    os.makedirs(os.path.join(work_root, "exp6_input", "pivot"), exist_ok=True)
    for rt in READ_TYPES:
        os.makedirs(os.path.join(work_root, "exp6_input", "pivot_reads_subset", rt), exist_ok=True)
    for i in range(1, num_datasets + 1):
        src = os.path.join(base, "exp0_nonpivot_genomes", f"dataset_{i}")
        dst = os.path.join(work_root, "exp6_input", "rest_of_set", f"dataset_{i}")
        os.makedirs(dst, exist_ok=True)
        for f in os.listdir(src):
            shutil.copy(os.path.join(src, f), dst)
        shutil.copy(os.path.join(base, "exp0_pivot_genomes", f"dataset_{i}", f"pivot_{i}.fna.gz"), os.path.join(work_root, "exp6_input", "pivot", f"pivot_{i}.fna.gz"))
        for rt in READ_TYPES:
            shutil.copy(os.path.join(base, "exp0_pivot_reads", f"dataset_{i}", rt, f"pivot_{i}_subset.fa"), os.path.join(work_root, p_reads(rt, i)))


def write_parse_time_files(work_root: str, k_values: Sequence[str], num_datasets: int) -> None:
    """tmp/, the `kmc_tools complex` operation files and the file lists per (k, read type) (exp_type_6.smk:27-29, 60-111)."""
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
        for rt in READ_TYPES:
            d = os.path.join(work_root, "exp6_filelists", f"k_{k}", rt)
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "pivots_filelist.txt"), "w") as fd:
                for piv in range(1, num_datasets + 1):
                    fd.write(f"{base_dir}/{p_dump_pivot(k, rt, piv)}\n")
            with open(os.path.join(d, "intersections_filelist.txt"), "w") as fd:
                for piv in range(1, num_datasets + 1):
                    for num in range(1, num_datasets + 1):
                        fd.write(f"{base_dir}/{p_dump_inter(k, rt, piv, num)}\n")


def concatenate_accuracies(work_root: str, trial: int) -> None:
    """Rule concatenate_accuracies_exp6 (exp_type_6.smk:351-365): a header line, then `cat .../values/k_*_accuracy_values.csv`
    (shell glob order = sorted names) per read type."""
    for rt in READ_TYPES:
        d = os.path.join(work_root, "exp6_accuracies", rt, "values")
        with open(os.path.join(work_root, p_final(trial, rt)), "w") as out:
            out.write(HEADER)
            for f in sorted(x for x in os.listdir(d) if x.startswith("k_") and x.endswith("_accuracy_values.csv")):
                with open(os.path.join(d, f)) as fd:
                    out.write(fd.read())


# ---- fused mode ----------------------------------------------------------------------------------------
def run_fused(work_root: str, num_datasets: int, k_values: Optional[Sequence] = None, engine: Optional[Engine] = None, trial: int = 1,
              level: str = "read", report_path: Optional[str] = None, seed_fn=None) -> Dict:
    """All of exp type 6 for ``work_root`` (``exp6_input/`` must exist, see prepare_inputs / synth.write_dataset_type6).
    seed_fn(read_type, k) -> int (optional): seed Python's global generator before every (read type, k) matrix, which makes
    the reference's random tie-breaking reproducible (the reference itself never seeds)."""
    if level not in ("read", "feature"):
        raise ValueError(level)
    k_values = [str(k) for k in (k_values or DEFAULT_K_VALUES)]
    own = engine is None
    eng = engine or Engine(int(os.environ.get("KHB_DEVICE", "0")))
    report = {"mode": "fused", "exp_type": 6, "level": level, "work_root": work_root, "num_datasets": num_datasets, "k_values": k_values, "stages": []}
    t_start = time.time()
    packed: Dict[int, object] = {}
    read_text: Dict[tuple, bytes] = {}
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
                tables.write_histogram_file(os.path.join(work_root, p_union_hist(k, num)), hist)   # rule union_histogram_exp6
                group_off.append(eng.group_sets_info()["n_keys"])
                report["stages"].append({"k": ki, "dataset": num, **st})
            for rt in READ_TYPES:
                bufs, counts, sizes = [], [], []
                try:
                    for piv in range(1, num_datasets + 1):
                        if (rt, piv) not in read_text:
                            read_text[(rt, piv)] = ingest.read_fasta(os.path.join(work_root, p_reads(rt, piv)))
                        buf, cnt, n = eng.kmer_counts(read_text[(rt, piv)], ki, cs=cli.KMC_DEFAULT_CS)   # rule build_kmc_database_on_pivot_exp6
                        bufs.append(buf); counts.append(cnt); sizes.append(n)
                    masks = eng.group_membership(group_off, bufs, sizes, ki)
                    per_pivot, at = [], 0
                    for n in sizes:
                        per_pivot.append(masks[at:at + n])
                        at += n
                    if level == "feature":
                        matrix, matrix_u = merge_lists.confusion_from_masks(counts, per_pivot, num_datasets)
                    else:
                        matrix, n_reads = [], 0
                        if seed_fn is not None:
                            random.seed(seed_fn(rt, k))
                        for piv in range(1, num_datasets + 1):
                            reads = merge_lists.split_reads(read_text[(rt, piv)])
                            merge_lists.check_reads(reads, ki)
                            votes, _ = eng.read_votes(reads, ki, bufs[piv - 1], sizes[piv - 1], per_pivot[piv - 1], num_datasets)
                            matrix.append(merge_lists.read_level_row(votes, num_datasets))
                            n_reads += len(reads)
                        matrix_u = [list(r) for r in matrix]
                        report["stages"].append({"k": ki, "read_type": rt, "reads": n_reads})
                finally:
                    for b in bufs:
                        b.free()
                merge_lists.write_outputs(os.path.join(work_root, "exp6_accuracies", rt) + "/", k, matrix, matrix_u, num_datasets)
                report["stages"].append({"k": ki, "read_type": rt, "pivot_kmers": int(sum(sizes))})
        concatenate_accuracies(work_root, trial)
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
def _rule_jobs(work_root: str, k_values: Sequence[str], num_datasets: int, level: str = "read"):
    """(rule name, outputs, shell string) per rule instance in topological order; the shell strings are the reference's
    (exp_type_6.smk:177, 187, 203-206, 221-224, 238-241, 256-259, 277-280, 291-294, 308-311, 322-325, 337-344) minus the
    `rm` clean-ups, with `python3 -m khoice_b200.merge_lists` in place of `python3 {repo_dir}/src/merge_lists.py`."""
    db = lambda p: [p + ".kmc_pre", p + ".kmc_suf"]
    base_dir = os.path.abspath(work_root)
    jobs = []
    for k in k_values:
        for num in range(1, num_datasets + 1):
            for g in rest_genomes_of(work_root, num):
                jobs.append(("build_kmc_database_on_genome_exp6", db(p_s1_rest(k, num, g)),
                             f"kmc -fm -m64 -k{k} -ci1 {p_rest(num, g)} {p_s1_rest(k, num, g)} tmp/"))
                jobs.append(("transform_genome_to_set_exp6", db(p_set_rest(k, num, g)),
                             f"kmc_tools transform {p_s1_rest(k, num, g)} set_counts 1 {p_set_rest(k, num, g)}"))
            jobs.append(("rest_of_set_union_exp6", db(p_union(k, num)), f"kmc_tools complex {p_ops(k, num)}"))
            jobs.append(("union_histogram_exp6", [p_union_hist(k, num)], f"kmc_tools transform {p_union(k, num)} histogram {p_union_hist(k, num)}"))
            jobs.append(("transform_union_to_set_exp6", db(p_union_set(k, num)), f"kmc_tools transform {p_union(k, num)} set_counts 1 {p_union_set(k, num)}"))
        for rt in READ_TYPES:
            for piv in range(1, num_datasets + 1):
                jobs.append(("build_kmc_database_on_pivot_exp6", db(p_s1_pivot(k, rt, piv)),
                             f"kmc -fm -m64 -k{k} -ci1 {p_reads(rt, piv)} {p_s1_pivot(k, rt, piv)} tmp/"))
                jobs.append(("pivot_text_dump_exp6", [p_dump_pivot(k, rt, piv)],
                             f"kmc_tools transform {p_s1_pivot(k, rt, piv)} dump -s {p_dump_pivot(k, rt, piv)}"))
                for num in range(1, num_datasets + 1):
                    jobs.append(("pivot_intersect_exp6", db(p_inter(k, rt, piv, num)),
                                 f"kmc_tools simple {p_union_set(k, num)} {p_s1_pivot(k, rt, piv)} intersect {p_inter(k, rt, piv, num)} -ocsum"))
                    jobs.append(("intersection_text_dump_exp6", [p_dump_inter(k, rt, piv, num)],
                                 f"kmc_tools transform {p_inter(k, rt, piv, num)} dump -s {p_dump_inter(k, rt, piv, num)}"))
            jobs.append(("run_merge_list_exp6", [p_values(rt, k), p_matrix(rt, k)],
                         f"{shlex.quote(sys.executable)} -m khoice_b200.merge_lists -p {base_dir}/exp6_filelists/k_{k}/{rt}/pivots_filelist.txt "
                         f"-i {base_dir}/exp6_filelists/k_{k}/{rt}/intersections_filelist.txt -o {base_dir}/exp6_accuracies/{rt}/ -n {num_datasets} -k {k}"
                         + (f" -r {base_dir}/exp6_input/pivot_reads_subset/{rt}/" if level == "read" else "")))
    return jobs


def run_rules(work_root: str, num_datasets: int, k_values: Optional[Sequence] = None, engine: Optional[Engine] = None, trial: int = 1,
              level: str = "read", seed_fn=None) -> Dict:
    """Run every exp-6 rule instance separately through the kmc / kmc_tools shims (in-process).  seed_fn: see run_fused."""
    k_values = [str(k) for k in (k_values or DEFAULT_K_VALUES)]
    write_parse_time_files(work_root, k_values, num_datasets)
    cwd = os.getcwd()
    ran, skipped = 0, 0
    own = engine or Engine(int(os.environ.get("KHB_DEVICE", "0")))
    cli.set_engine(own)
    try:
        os.chdir(work_root)
        for rule, outputs, shell in _rule_jobs(".", k_values, num_datasets, level):
            if all(os.path.exists(o) for o in outputs):
                skipped += 1
                continue
            for o in outputs:
                os.makedirs(os.path.dirname(o) or ".", exist_ok=True)
            argv = shlex.split(shell)
            if argv[0] in ("kmc", "kmc_tools"):
                rc = cli.main(argv)
            else:
                if seed_fn is not None:   # outputs[0] = exp6_accuracies/{rt}/values/k_{k}_accuracy_values.csv
                    random.seed(seed_fn(outputs[0].split("/")[-3], outputs[0].split("/")[-1].split("_")[1]))
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
    concatenate_accuracies(work_root, trial)
    return {"mode": "rules", "exp_type": 6, "level": level, "jobs_run": ran, "jobs_skipped": skipped}


def main(argv: Optional[List[str]] = None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="khoice experiment type 6 (confusion matrix from simulated reads) on the B200 engine")
    ap.add_argument("--work-root", required=True)
    ap.add_argument("--num-datasets", type=int, required=True)
    ap.add_argument("--k-values", default=None)
    ap.add_argument("--database-root", default=None, help="DB_ROOT: copy inputs from {root}/trial_{t}/exp0_* first")
    ap.add_argument("--trial", type=int, default=1)
    ap.add_argument("--level", choices=["read", "feature"], default="read")
    ap.add_argument("--mode", choices=["fused", "rules"], default="fused")
    ap.add_argument("--report", default=None)
    a = ap.parse_args(argv)
    ks = a.k_values.split(",") if a.k_values else None
    if a.database_root:
        prepare_inputs(a.work_root, a.database_root, a.trial, a.num_datasets)
    if a.mode == "fused":
        rep = run_fused(a.work_root, a.num_datasets, ks, trial=a.trial, level=a.level, report_path=a.report)
    else:
        rep = run_rules(a.work_root, a.num_datasets, ks, trial=a.trial, level=a.level)
    print(json.dumps({k: v for k, v in rep.items() if k != "stages"}))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
