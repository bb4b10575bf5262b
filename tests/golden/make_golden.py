"""Generate the golden fixtures under tests/golden/ by executing the REFERENCE's own Python code.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The GPU box has no /root/reference, so the outputs are committed:

  summarize_type1.json   inputs/outputs of summarize_histogram_type1 (exp_type_1.smk:115-150)
  step5_case*.csv, step9_case*.csv + tables_cases.json
                         the `run:` bodies of within_group_union_analysis (exp_type_1.smk:199-231) and
                         across_group_union_analysis (:268-297) executed on synthetic histogram files
  complex_ops.json       the operation files written by the parse-time block (exp_type_1.smk:26-84)
  canonical.json         get_canonical_kmer (src/merge_lists.py:60-73) on random k-mers
  summarize_type2.json   inputs/outputs of summarize_histogram_type2 (exp_type_2.smk:171-216)
  t2_within_case*.csv, t2_across_case*.csv + tables_cases_type2.json
                         the `run:` bodies of within_group_analysis_exp_type2 (exp_type_2.smk:404-438) and
                         across_group_analysis_exp_type2 (:521-554) executed on synthetic histogram files

Nothing from the reference is copied into the repo: its source is read from /root/reference at run time,
sliced by line number and exec'd.  The reference environment is Python 3.10, where sum() over floats is
a plain left-to-right accumulation; newer interpreters compensate, so the exec namespace gets a 3.10-style
`sum` (documented deviation from "exec as is"; outputs of both variants are compared and any difference is
recorded in the fixture under "py312_sum_differs").
"""
import json
import os
import random
import shutil
import sys
import tempfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def py310_sum(values, start=0):
    acc = start
    for v in values:
        acc = acc + v
    return acc


def slice_source(path, first, last, dedent=0):
    with open(path) as fd:
        lines = fd.readlines()[first - 1:last]
    return "".join(line[dedent:] if line.strip() else line for line in lines)


def load_summarize(sum_impl):
    ns = {"sum": sum_impl}
    exec(slice_source(f"{REF}/workflow/rules/exp_type_1.smk", 115, 150), ns)
    return ns["summarize_histogram_type1"]


def hist_cases():
    rnd = random.Random(1234)
    cases = []
    # (hist, members, across, k)
    cases.append(([6, 1] + [0] * 4998, 2, False, 3))
    cases.append(([7, 0] + [0] * 4998, 2, False, 3))
    cases.append(([10, 2] + [0] * 4998, 2, True, 3))
    cases.append(([1000, 200, 50, 10, 4000] + [0] * 250, 5, False, 31))
    for _ in range(60):
        members = rnd.choice([2, 3, 4, 5, 8, 10, 37, 50, 100, 200, 1000])
        across = rnd.random() < 0.4
        k = rnd.choice([7, 8, 15, 21, 30, 31, 34, 49, 63])
        n_rows = rnd.choice([20, 255, 5000])
        top = min(members, n_rows)
        h = [0] * n_rows
        for i in range(top):
            if rnd.random() < 0.8:
                h[i] = int(rnd.paretovariate(0.6) * rnd.choice([1, 10, 1000, 100000]))
        if sum(h) == 0:
            h[0] = 1
        if h[0] == 0 and rnd.random() < 0.5:
            h[0] = rnd.randrange(1, 10**7)
        cases.append((h, members, across, k))
    return cases


def trim(h):
    """Drop trailing zero rows (the fixture stores the row count separately)."""
    n = len(h)
    while n > 1 and h[n - 1] == 0:
        n -= 1
    return h[:n]


def make_summarize():
    f310, f312 = load_summarize(py310_sum), load_summarize(sum)
    out, differs = [], 0
    for h, members, across, k in hist_cases():
        try:
            m = f310(list(h), members, across, k)
            m2 = f312(list(h), members, across, k)
            differs += int([float(x) for x in m] != [float(x) for x in m2])
            out.append({"hist": trim(h), "rows": len(h), "members": members, "across": across, "k": k,
                        "metrics": [float(x) for x in m], "repr": [str(x) for x in m]})
        except (AssertionError, IndexError) as e:
            out.append({"hist": trim(h), "rows": len(h), "members": members, "across": across, "k": k, "raises": type(e).__name__})
    with open(os.path.join(HERE, "summarize_type1.json"), "w") as fd:
        json.dump({"source": "exp_type_1.smk:115-150 exec'd with py3.10 sum", "py312_sum_differs": differs, "cases": out}, fd)
    print("summarize cases:", len(out), "py3.12-sum differences:", differs)


def write_hist(path, h):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fd:
        for i, c in enumerate(h):
            fd.write(f"{i + 1}\t{c}\n")


def make_tables():
    rnd = random.Random(99)
    cases = []
    for case, (num_datasets, members, k_values) in enumerate([(2, [5, 5], ["7", "15", "31"]), (3, [4, 9, 2], ["8", "21", "30", "34"]),
                                                              (10, [50] * 10, ["31"])]):
        work = tempfile.mkdtemp(prefix="khb_golden_")
        cwd = os.getcwd()
        os.chdir(work)
        try:
            for n in range(1, num_datasets + 1):
                os.makedirs(f"data/dataset_{n}")
                for g in range(members[n - 1]):
                    open(f"data/dataset_{n}/genome_{g}.fna.gz", "wb").close()
                open(f"data/dataset_{n}/README.txt", "w").close()  # must not be counted
            hists = {}
            for k in k_values:
                for n in range(1, num_datasets + 1):
                    h = [0] * 5000
                    for i in range(members[n - 1]):
                        h[i] = rnd.randrange(0, 4 ** min(int(k), 11)) if rnd.random() < 0.9 else 0
                    h[0] += 1
                    hists[f"step_4/k_{k}/dataset_{n}/dataset_{n}_k{k}_hist.txt"] = h
                h = [0] * 5000
                for i in range(num_datasets):
                    h[i] = rnd.randrange(0, 4 ** min(int(k), 11))
                h[0] += 1
                hists[f"step_8/k_{k}/all_datasets_k{k}_hist.txt"] = h
            for p, h in hists.items():
                write_hist(p, h)
            ns = {"sum": py310_sum, "os": os, "num_datasets": num_datasets, "k_values": k_values}
            exec(slice_source(f"{REF}/workflow/rules/exp_type_1.smk", 107, 150), ns)   # helpers
            # step_5: rule body, `input` in expand() order (k-major), `output[0]`
            ns["input"] = [f"step_4/k_{k}/dataset_{n}/dataset_{n}_k{k}_hist.txt" for k in k_values for n in range(1, num_datasets + 1)]
            ns["output"] = ["step_5/within_datasets_analysis.csv"]
            os.makedirs("step_5"); os.makedirs("step_9")
            exec(slice_source(f"{REF}/workflow/rules/exp_type_1.smk", 199, 231, dedent=8), ns)
            ns["input"] = [f"step_8/k_{k}/all_datasets_k{k}_hist.txt" for k in k_values]
            ns["output"] = ["step_9/across_datasets_analysis.csv"]
            exec(slice_source(f"{REF}/workflow/rules/exp_type_1.smk", 268, 297, dedent=8), ns)
            shutil.copyfile("step_5/within_datasets_analysis.csv", os.path.join(HERE, f"step5_case{case}.csv"))
            shutil.copyfile("step_9/across_datasets_analysis.csv", os.path.join(HERE, f"step9_case{case}.csv"))
            cases.append({"num_datasets": num_datasets, "members": members, "k_values": k_values,
                          "hists": {p: [c for c in h if True][:max(max(members), num_datasets) + 2] for p, h in hists.items()}})
        finally:
            os.chdir(cwd)
            shutil.rmtree(work)
    with open(os.path.join(HERE, "tables_cases.json"), "w") as fd:
        json.dump({"source": "exp_type_1.smk:199-231 and :268-297 exec'd with py3.10 sum; hists truncated (rest is zeros, 5000 rows)",
                   "cases": cases}, fd)
    print("table cases:", len(cases))


def make_complex_ops():
    work = tempfile.mkdtemp(prefix="khb_golden_")
    cwd = os.getcwd()
    os.chdir(work)
    try:
        members = {1: ["b_genome", "a_genome", "c.v2_genome"], 2: ["solo"]}
        for n, names in members.items():
            os.makedirs(f"data/dataset_{n}")
            for g in names:
                open(f"data/dataset_{n}/{g}.fna.gz", "wb").close()
            open(f"data/dataset_{n}/notes.txt", "w").close()
        ns = {"os": os, "exp_type": 1, "k_values": ["7", "31"], "num_datasets": 2}
        exec(slice_source(f"{REF}/workflow/rules/exp_type_1.smk", 26, 84), ns)
        files = {}
        for root, _, fs in os.walk("complex_ops"):
            for f in fs:
                p = os.path.join(root, f)
                files[p] = open(p).read()
        assert os.path.isdir("tmp")
    finally:
        os.chdir(cwd)
        shutil.rmtree(work)
    with open(os.path.join(HERE, "complex_ops.json"), "w") as fd:
        json.dump({"source": "exp_type_1.smk:26-84 exec'd; set numbering follows os.listdir order (arbitrary), compare as sets",
                   "members": {str(k): v for k, v in members.items()}, "k_values": ["7", "31"], "files": files}, fd, indent=1)
    print("complex_ops files:", len(files))


def make_canonical():
    ns = {}
    exec(slice_source(f"{REF}/src/merge_lists.py", 60, 73), ns)
    canon = ns["get_canonical_kmer"]
    rnd = random.Random(7)
    vec = []
    for k in [1, 2, 3, 4, 5, 8, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64]:
        for _ in range(12):
            s = "".join(rnd.choice("ACGT") for _ in range(k))
            vec.append([s, canon(s)])
        pal = "".join(rnd.choice("ACGT") for _ in range(k // 2))
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        if k % 2 == 0:
            s = pal + "".join(comp[c] for c in reversed(pal))   # reverse-complement palindrome
            vec.append([s, canon(s)])
        vec.append(["A" * k, canon("A" * k)])
        vec.append(["T" * k, canon("T" * k)])
    with open(os.path.join(HERE, "canonical.json"), "w") as fd:
        json.dump({"source": "src/merge_lists.py:60-73 get_canonical_kmer exec'd", "vectors": vec}, fd)
    print("canonical vectors:", len(vec))


def load_summarize2(sum_impl):
    ns = {"sum": sum_impl}
    exec(slice_source(f"{REF}/workflow/rules/exp_type_2.smk", 171, 216), ns)
    return ns["summarize_histogram_type2"]


def t2_hist_pair(rnd, members, rows, scale):
    """(sub_counts, inter_counts) like the two kmc_tools histograms of exp 2: sub has only row 1, inter has no row 1."""
    sub = [0] * rows
    sub[0] = rnd.randrange(0, scale)
    inter = [0] * rows
    for i in range(1, min(members + 1, rows)):
        if rnd.random() < 0.85:
            inter[i] = rnd.randrange(0, scale)
    if sub[0] + sum(inter) == 0:
        sub[0] = 1
    return sub, inter


def make_summarize2():
    f310, f312 = load_summarize2(py310_sum), load_summarize2(sum)
    rnd = random.Random(4321)
    cases = [([5] + [0] * 19, [0, 3, 2] + [0] * 17, 4, False, 31),
             ([5] + [0] * 19, [1, 3, 2] + [0] * 17, 4, False, 31),        # raises: intersection has unique k-mers
             ([5, 1] + [0] * 18, [0, 3, 2] + [0] * 17, 4, False, 31),     # raises: subtract has non-unique k-mers
             ([0] * 5000, [0, 7] + [0] * 4998, 9, True, 12)]
    for _ in range(60):
        members = rnd.choice([1, 2, 3, 4, 5, 9, 10, 37, 50, 100, 200])
        across = rnd.random() < 0.5
        rows = rnd.choice([20, 255, 5000])
        sub, inter = t2_hist_pair(rnd, members, rows, rnd.choice([10, 1000, 10**6, 10**9]))
        cases.append((sub, inter, members, across, rnd.choice([7, 8, 15, 21, 30, 31, 34, 49, 63])))
    out, differs = [], 0
    for sub, inter, members, across, k in cases:
        rec = {"sub": trim(sub), "inter": trim(inter), "rows": len(inter), "members": members, "across": across, "k": k}
        try:
            m = f310(list(sub), list(inter), members, across, k)
            m2 = f312(list(sub), list(inter), members, across, k)
            differs += int([float(x) for x in m] != [float(x) for x in m2])
            rec.update(metrics=[float(x) for x in m], repr=[str(x) for x in m])
        except (AssertionError, IndexError, ZeroDivisionError) as e:
            rec["raises"] = type(e).__name__
        out.append(rec)
    with open(os.path.join(HERE, "summarize_type2.json"), "w") as fd:
        json.dump({"source": "exp_type_2.smk:171-216 exec'd with py3.10 sum", "py312_sum_differs": differs, "cases": out}, fd)
    print("summarize2 cases:", len(out), "raising:", sum(1 for c in out if "raises" in c), "py3.12-sum differences:", differs)


def make_tables2():
    rnd = random.Random(777)
    cases = []
    for case, (num_datasets, members, k_values) in enumerate([(2, [4, 4], ["7", "15", "31"]), (3, [3, 8, 1], ["8", "21", "30", "34"]),
                                                              (10, [49] * 10, ["31"])]):
        work = tempfile.mkdtemp(prefix="khb_golden_")
        cwd = os.getcwd()
        os.chdir(work)
        try:
            for n in range(1, num_datasets + 1):
                os.makedirs(f"input_type_2/rest_of_set/dataset_{n}")
                for g in range(members[n - 1]):
                    open(f"input_type_2/rest_of_set/dataset_{n}/genome_{g}.fna.gz", "wb").close()
                open(f"input_type_2/rest_of_set/dataset_{n}/nonpivot_names.txt", "w").close()  # must not be counted
            hists = {}
            for scope, top_of in (("within", lambda n: members[n - 1]), ("across", lambda n: num_datasets - 1)):
                for n in range(1, num_datasets + 1):
                    for k in k_values:
                        sub, inter = t2_hist_pair(rnd, top_of(n), 5000, 4 ** min(int(k), 11))
                        base = f"{scope}_dataset_results_type_2/k_{k}/dataset_{n}"
                        hists[f"{base}/subtract/dataset_{n}_pivot_subtract_group.hist.txt"] = sub
                        hists[f"{base}/intersect/dataset_{n}_pivot_intersect_group.hist.txt"] = inter
            for p, h in hists.items():
                write_hist(p, h)
            ns = {"sum": py310_sum, "os": os, "num_datasets": num_datasets, "k_values": k_values}
            exec(slice_source(f"{REF}/workflow/rules/exp_type_2.smk", 123, 129), ns)   # get_num_of_dataset_members_exp2
            exec(slice_source(f"{REF}/workflow/rules/exp_type_2.smk", 153, 216), ns)   # file lists + summarize_histogram_type2
            os.makedirs("within_dataset_analysis_type_2"); os.makedirs("across_dataset_analysis_type_2")
            ns["input"] = ns["get_within_group_histogram_files"](None)
            ns["output"] = ["within_dataset_analysis_type_2/within_dataset_analysis.csv"]
            exec(slice_source(f"{REF}/workflow/rules/exp_type_2.smk", 404, 438, dedent=8), ns)
            ns["input"] = ns["get_across_group_histogram_files"](None)
            ns["output"] = ["across_dataset_analysis_type_2/across_dataset_analysis.csv"]
            exec(slice_source(f"{REF}/workflow/rules/exp_type_2.smk", 521, 554, dedent=8), ns)
            shutil.copyfile(ns["output"][0], os.path.join(HERE, f"t2_across_case{case}.csv"))
            shutil.copyfile("within_dataset_analysis_type_2/within_dataset_analysis.csv", os.path.join(HERE, f"t2_within_case{case}.csv"))
            cases.append({"num_datasets": num_datasets, "members": members, "k_values": k_values,
                          "hists": {p: h[:max(max(members), num_datasets) + 3] for p, h in hists.items()}})
        finally:
            os.chdir(cwd)
            shutil.rmtree(work)
    with open(os.path.join(HERE, "tables_cases_type2.json"), "w") as fd:
        json.dump({"source": "exp_type_2.smk:404-438 and :521-554 exec'd with py3.10 sum; hists truncated (rest is zeros, 5000 rows)",
                   "cases": cases}, fd)
    print("type-2 table cases:", len(cases))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    make_summarize()
    make_tables()
    make_complex_ops()
    make_canonical()
    make_summarize2()
    make_tables2()
