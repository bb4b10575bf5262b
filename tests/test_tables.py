"""step_5 / step_9 builders and the histogram text format against fixtures produced by executing the
reference's own code (tests/golden/make_golden.py)."""
import json
import os

import pytest

from khoice_b200 import tables

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_summarize_matches_reference_outputs():
    data = json.load(open(os.path.join(GOLDEN, "summarize_type1.json")))
    assert data["py312_sum_differs"] == 0
    n_ok = 0
    for c in data["cases"]:
        hist = c["hist"] + [0] * (c["rows"] - len(c["hist"]))
        if "raises" in c:
            with pytest.raises((AssertionError, IndexError)):
                tables.summarize_histogram_type1(hist, c["members"], c["across"], c["k"])
            continue
        got = tables.summarize_histogram_type1(hist, c["members"], c["across"], c["k"])
        assert got == c["metrics"]
        assert [str(x) for x in got] == c["repr"]      # the CSV is str() of these values
        n_ok += 1
    assert n_ok >= 40


def test_step5_and_step9_csv_bytes(tmp_path):
    cases = json.load(open(os.path.join(GOLDEN, "tables_cases.json")))["cases"]
    for idx, c in enumerate(cases):
        root = tmp_path / f"case{idx}"
        for n, m in enumerate(c["members"], start=1):
            d = root / "data" / f"dataset_{n}"
            d.mkdir(parents=True)
            for g in range(m):
                (d / f"genome_{g}.fna.gz").write_bytes(b"")
            (d / "README.txt").write_text("not a genome")
        for p, h in c["hists"].items():
            tables.write_histogram_file(str(root / p), [0] + h)
            lines = (root / p).read_text().splitlines()
            assert len(lines) == 5000 and lines[0].split("\t") == ["1", str(h[0])] and lines[-1] == "5000\t0"
            assert tables.read_histogram_file(str(root / p))[:len(h)] == h
        nums = range(1, c["num_datasets"] + 1)
        s4 = [str(root / f"step_4/k_{k}/dataset_{n}/dataset_{n}_k{k}_hist.txt") for k in c["k_values"] for n in nums]
        s8 = [str(root / f"step_8/k_{k}/all_datasets_k{k}_hist.txt") for k in c["k_values"]]
        tables.within_group_union_analysis(s4, str(root / "step_5/within_datasets_analysis.csv"), c["num_datasets"],
                                           lambda n: tables.get_num_of_dataset_members(n, str(root / "data")))
        tables.across_group_union_analysis(s8, str(root / "step_9/across_datasets_analysis.csv"), c["num_datasets"])
        assert (root / "step_5/within_datasets_analysis.csv").read_bytes() == open(os.path.join(GOLDEN, f"step5_case{idx}.csv"), "rb").read()
        assert (root / "step_9/across_datasets_analysis.csv").read_bytes() == open(os.path.join(GOLDEN, f"step9_case{idx}.csv"), "rb").read()


def test_histogram_needs_twenty_rows_like_the_reference():
    # the across-group summary indexes rows 5..19 even when there are 2 groups
    with pytest.raises(IndexError):
        tables.summarize_histogram_type1([3, 1], 2, True, 7)
    assert tables.summarize_histogram_type1([3, 1] + [0] * 18, 2, True, 7)[0] == 0.75


# ---- experiment type 2 (exp_type_2.smk) ------------------------------------------------------------------
def test_summarize_type2_matches_reference_outputs():
    data = json.load(open(os.path.join(GOLDEN, "summarize_type2.json")))
    assert data["py312_sum_differs"] == 0
    n_ok = n_raise = 0
    for c in data["cases"]:
        sub = c["sub"] + [0] * (c["rows"] - len(c["sub"]))
        inter = c["inter"] + [0] * (c["rows"] - len(c["inter"]))
        if "raises" in c:
            with pytest.raises((AssertionError, IndexError, ZeroDivisionError)):
                tables.summarize_histogram_type2(sub, inter, c["members"], c["across"], c["k"])
            n_raise += 1
            continue
        got = tables.summarize_histogram_type2(sub, inter, c["members"], c["across"], c["k"])
        assert got == c["metrics"]
        assert [str(x) for x in got] == c["repr"]
        n_ok += 1
    assert n_ok >= 40 and n_raise >= 2


def test_type2_csv_bytes(tmp_path):
    cases = json.load(open(os.path.join(GOLDEN, "tables_cases_type2.json")))["cases"]
    for idx, c in enumerate(cases):
        root = tmp_path / f"case{idx}"
        for n, m in enumerate(c["members"], start=1):
            d = root / "input_type_2" / "rest_of_set" / f"dataset_{n}"
            d.mkdir(parents=True)
            for g in range(m):
                (d / f"genome_{g}.fna.gz").write_bytes(b"")
            (d / "nonpivot_names.txt").write_text("not a genome")
        for p, h in c["hists"].items():
            tables.write_histogram_file(str(root / p), [0] + h)
        nums = range(1, c["num_datasets"] + 1)

        def files(scope):
            return [str(root / f"{scope}_dataset_results_type_2/k_{k}/dataset_{n}/{op}/dataset_{n}_pivot_{op}_group.hist.txt")
                    for n in nums for k in c["k_values"] for op in ("subtract", "intersect")]

        w_csv = root / "within_dataset_analysis_type_2/within_dataset_analysis.csv"
        a_csv = root / "across_dataset_analysis_type_2/across_dataset_analysis.csv"
        tables.within_group_analysis_exp_type2(files("within"), str(w_csv), c["num_datasets"],
                                               lambda n: tables.get_num_of_dataset_members_exp2(n, str(root / "input_type_2")))
        tables.across_group_analysis_exp_type2(files("across"), str(a_csv), c["num_datasets"])
        assert w_csv.read_bytes() == open(os.path.join(GOLDEN, f"t2_within_case{idx}.csv"), "rb").read()
        assert a_csv.read_bytes() == open(os.path.join(GOLDEN, f"t2_across_case{idx}.csv"), "rb").read()
