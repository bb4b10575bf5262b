"""Golden outputs of experiment type 4 made by running the REFERENCE's own program, /root/reference/src/merge_lists.py
(unmodified, as a sub-process with the command line of rule run_merge_list_exp_type_4, exp_type_4.smk:284-290), on text
dumps that the CPU oracle writes for deterministic synthetic genomes.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_exp4.py
Outputs (committed; the dumps themselves are not -- the tests regenerate the same genomes from the same seeds):
  exp4_case{c}_k{k}_confusion_matrix.txt, ..._with_unidentified.txt, ..._accuracy_values.csv, exp4_cases.json
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_SCRIPT = "/root/reference/src/merge_lists.py"

CASES = [  # (n_groups, genomes_per_group incl. the pivot, genome_len, seed, k values, out_pivot)
    dict(n_groups=3, genomes_per_group=4, genome_len=12_000, seed=41, k_values=[9, 21, 31], out_pivot=True),
    dict(n_groups=4, genomes_per_group=3, genome_len=6_000, seed=42, k_values=[15, 35], out_pivot=False),
]


def inputs_of(case):
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=case["n_groups"], genomes_per_group=case["genomes_per_group"], genome_len=case["genome_len"], seed=case["seed"])
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, cfg.genomes_per_group)] for g in range(1, cfg.n_groups + 1)]
    pivots = [synth.make_genome(cfg, g, cfg.genomes_per_group) for g in range(1, cfg.n_groups + 1)]
    if not case["out_pivot"]:
        groups = [grp + [p] for grp, p in zip(groups, pivots)]
    return cfg, groups, pivots


def main():
    if not os.path.isfile(REF_SCRIPT):
        sys.exit("needs /root/reference (build container only)")
    from khoice_b200 import kmcdb
    from oracle import oracle as O
    for c, case in enumerate(CASES):
        cfg, groups, pivots = inputs_of(case)
        G = case["n_groups"]
        for k in case["k_values"]:
            work = tempfile.mkdtemp(prefix="khb_golden4_")
            try:
                tables, inters = O.exp4(groups, pivots, k)
                pl, il = [], []
                for p in range(G):
                    f = f"{work}/pivot_{p + 1}.txt"
                    kmcdb.write_text_dump(f, tables[p][0], tables[p][1], k)
                    pl.append(f)
                    for d in range(G):
                        f = f"{work}/pivot_{p + 1}_intersect_dataset_{d + 1}.txt"
                        kmcdb.write_text_dump(f, inters[p][d][0], inters[p][d][1], k)
                        il.append(f)
                open(f"{work}/pl.txt", "w").write("\n".join(pl) + "\n")
                open(f"{work}/il.txt", "w").write("\n".join(il) + "\n")
                os.makedirs(f"{work}/out/confusion_matrix")
                os.makedirs(f"{work}/out/values")
                r = subprocess.run([sys.executable, REF_SCRIPT, "-p", f"{work}/pl.txt", "-i", f"{work}/il.txt", "-o", f"{work}/out/",
                                    "-n", str(G), "-k", str(k)], capture_output=True, text=True)
                assert r.returncode == 0, r.stdout + r.stderr
                for src, dst in ((f"confusion_matrix/k_{k}_confusion_matrix.txt", "confusion_matrix.txt"),
                                 (f"confusion_matrix/k_{k}_confusion_matrix_with_unidentified.txt", "confusion_matrix_with_unidentified.txt"),
                                 (f"values/k_{k}_accuracy_values.csv", "accuracy_values.csv")):
                    shutil.copyfile(f"{work}/out/{src}", os.path.join(HERE, f"exp4_case{c}_k{k}_{dst}"))
            finally:
                shutil.rmtree(work)
        print("case", c, "done")
    with open(os.path.join(HERE, "exp4_cases.json"), "w") as fd:
        json.dump({"source": "src/merge_lists.py run unmodified on oracle-written text dumps (tests/golden/make_golden_exp4.py)", "cases": CASES}, fd, indent=1)


if __name__ == "__main__":
    main()
