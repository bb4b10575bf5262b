"""Golden outputs of experiment type 6 (read-level confusion matrix) made by running the REFERENCE's own program,
/root/reference/src/merge_lists.py -- executed unmodified, in this interpreter, with the command line of rule
run_merge_list_exp6 (exp_type_6.smk:337-344, i.e. with -r) -- on text dumps that the CPU oracle writes for deterministic
synthetic genomes and reads.  The reference breaks ties with random.choice on Python's global, unseeded generator; to
have reproducible fixtures the generator is seeded (random.seed(seed_of(...))) right before the script runs, and the tests
seed it the same way before running this package's code, which makes the same random.choice calls in the same order.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_exp6.py
Outputs (committed): exp6_case{c}_{read_type}_k{k}_confusion_matrix.txt, ..._accuracy_values.csv, exp6_cases.json
"""
import contextlib
import io
import json
import os
import random
import runpy
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_SCRIPT = "/root/reference/src/merge_lists.py"
READ_TYPES = ("illumina", "ont")

CASES = [  # genomes_per_group includes the pivot (the last genome), whose reads are the pivots of this experiment
    dict(n_groups=3, genomes_per_group=3, genome_len=8_000, seed=61, k_values=[15, 31], n_reads=60),
    dict(n_groups=2, genomes_per_group=4, genome_len=5_000, seed=62, k_values=[9, 21], n_reads=40),
]


def seed_of(c, read_type, k):
    return 20240606 + 1000 * c + 10 * int(k) + READ_TYPES.index(read_type)


def inputs_of(case):
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=case["n_groups"], genomes_per_group=case["genomes_per_group"], genome_len=case["genome_len"], seed=case["seed"])
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, cfg.genomes_per_group)] for g in range(1, cfg.n_groups + 1)]
    reads = {rt: [synth.make_reads(cfg, g, rt, case["n_reads"]) for g in range(1, cfg.n_groups + 1)] for rt in READ_TYPES}
    # ties and reads without k-mers must be exercised: an empty read, a read shorter than any k, a read of a foreign sequence
    reads["illumina"][0] += b">empty\n\n>short\nACGTA\n>foreign\n" + b"ACGTTGCATG" * 12 + b"\n"
    return cfg, groups, reads


def write_case_files(work, groups, read_texts, k, oracle, kmcdb):
    """pivot / intersection dumps (what the rules' `dump -s` writes), the two file lists and the reads directory."""
    G = len(groups)
    tables, inters = oracle.exp4(groups, read_texts, k)
    pl, il = [], []
    os.makedirs(f"{work}/reads")
    for p in range(G):
        f = f"{work}/pivot_{p + 1}.txt"
        kmcdb.write_text_dump(f, tables[p][0], tables[p][1], k)
        pl.append(f)
        open(f"{work}/reads/pivot_{p + 1}.fa", "wb").write(read_texts[p])
        for d in range(G):
            f = f"{work}/pivot_{p + 1}_intersect_dataset_{d + 1}.txt"
            kmcdb.write_text_dump(f, inters[p][d][0], inters[p][d][1], k)
            il.append(f)
    open(f"{work}/pl.txt", "w").write("\n".join(pl) + "\n")
    open(f"{work}/il.txt", "w").write("\n".join(il) + "\n")
    os.makedirs(f"{work}/out/confusion_matrix")
    os.makedirs(f"{work}/out/values")
    return ["-p", f"{work}/pl.txt", "-i", f"{work}/il.txt", "-o", f"{work}/out/", "-n", str(G), "-k", str(k), "-r", f"{work}/reads/"]


def main():
    if not os.path.isfile(REF_SCRIPT):
        sys.exit("needs /root/reference (build container only)")
    from khoice_b200 import kmcdb
    from oracle import oracle as O
    for c, case in enumerate(CASES):
        cfg, groups, reads = inputs_of(case)
        for rt in READ_TYPES:
            for k in case["k_values"]:
                work = tempfile.mkdtemp(prefix="khb_golden6_")
                try:
                    argv = write_case_files(work, groups, reads[rt], k, O, kmcdb)
                    old_argv = sys.argv
                    sys.argv = [REF_SCRIPT] + argv
                    random.seed(seed_of(c, rt, k))
                    try:
                        with contextlib.redirect_stdout(io.StringIO()):
                            runpy.run_path(REF_SCRIPT, run_name="__main__")
                    finally:
                        sys.argv = old_argv
                    a = open(f"{work}/out/confusion_matrix/k_{k}_confusion_matrix.txt", "rb").read()
                    b = open(f"{work}/out/confusion_matrix/k_{k}_confusion_matrix_with_unidentified.txt", "rb").read()
                    assert a == b      # read level: the two matrices receive the same increments (merge_lists.py:180-181)
                    shutil.copyfile(f"{work}/out/confusion_matrix/k_{k}_confusion_matrix.txt", os.path.join(HERE, f"exp6_case{c}_{rt}_k{k}_confusion_matrix.txt"))
                    shutil.copyfile(f"{work}/out/values/k_{k}_accuracy_values.csv", os.path.join(HERE, f"exp6_case{c}_{rt}_k{k}_accuracy_values.csv"))
                finally:
                    shutil.rmtree(work)
        print("case", c, "done")
    with open(os.path.join(HERE, "exp6_cases.json"), "w") as fd:
        json.dump({"source": "src/merge_lists.py (read level, -r) executed unmodified under random.seed(seed_of(case, read type, k)) on oracle-written "
                             "text dumps (tests/golden/make_golden_exp6.py)", "cases": CASES}, fd, indent=1)


if __name__ == "__main__":
    main()
