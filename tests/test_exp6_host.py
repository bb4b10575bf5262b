"""Experiment type 6, host side (no GPU): khoice_b200.merge_lists at the READ level (-r) against the outputs of the
reference's own src/merge_lists.py (tests/golden/make_golden_exp6.py) on dumps the CPU oracle writes for the same synthetic
genomes and reads, under the same random.seed (the reference draws ties with random.choice)."""
import json
import os
import random
import sys

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)


def _golden(c, rt, k, name):
    return open(os.path.join(GOLDEN, f"exp6_case{c}_{rt}_k{k}_{name}"), "rb").read()


def test_merge_lists_read_level_matches_reference_program(oracle, tmp_path):
    from khoice_b200 import kmcdb, merge_lists
    import make_golden_exp6 as G6
    cases = json.load(open(os.path.join(GOLDEN, "exp6_cases.json")))["cases"]
    assert cases == G6.CASES
    for c, case in enumerate(cases):
        cfg, groups, reads = G6.inputs_of(case)
        for rt in G6.READ_TYPES:
            for k in case["k_values"]:
                work = tmp_path / f"c{c}_{rt}_k{k}"
                work.mkdir()
                argv = G6.write_case_files(str(work), groups, reads[rt], k, oracle, kmcdb)
                random.seed(G6.seed_of(c, rt, k))
                assert merge_lists.main(argv) == 0
                out = str(work / "out") + "/"
                got = open(out + f"confusion_matrix/k_{k}_confusion_matrix.txt", "rb").read()
                assert got == _golden(c, rt, k, "confusion_matrix.txt"), (c, rt, k)
                assert open(out + f"confusion_matrix/k_{k}_confusion_matrix_with_unidentified.txt", "rb").read() == got
                assert open(out + f"values/k_{k}_accuracy_values.csv", "rb").read() == _golden(c, rt, k, "accuracy_values.csv")


def test_reads_are_split_and_checked_like_the_reference():
    from khoice_b200 import merge_lists
    text = b">r1\nACGT\n\n>r2 x\nTTGA \r\nAC>GT\nNNNN"
    assert merge_lists.split_reads(text) == [b"ACGT", b"", b"TTGA", b"NNNN"]
    merge_lists.check_reads([b"ACGT", b"", b"NNN"], 4)              # too short to have a k-mer: never looked at
    with pytest.raises(KeyError):
        merge_lists.check_reads([b"ACGTN"], 4)
    with pytest.raises(KeyError):
        merge_lists.check_reads([b"acgta"], 4)                      # lower case is a KeyError in the reference's rev_comp_dict


def test_tie_breaking_consumes_the_generator_like_the_reference():
    from khoice_b200 import merge_lists
    votes = np.array([[1.0, 1.0, 0.5], [0.0, 0.0, 0.0], [0.25, 0.75, 0.75], [2.0, 1.0, 0.0]])
    random.seed(5)
    ours = merge_lists.read_level_row(votes, 3)
    random.seed(5)
    ref = [0, 0, 0, 0]
    for v in votes.tolist():
        votes_np = np.array(v)
        max_indexes = np.where(votes_np == max(v))[0]               # merge_lists.py:176-178
        ref[random.choice(max_indexes)] += 1
    assert ours == ref and sum(ours) == 4
