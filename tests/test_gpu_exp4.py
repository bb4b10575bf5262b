"""Experiment type 4 (feature-level confusion matrix, exp_type_4.smk + src/merge_lists.py) on a GPU: group membership
bitmasks, counted pivot k-mers, and the whole rule chain (fused and rule-by-rule) against fixtures produced by the
reference's own merge_lists.py (tests/golden/make_golden_exp4.py)."""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)


@pytest.mark.parametrize("k", [7, 21, 31, 32, 45, 64])
def test_kmer_counts_and_membership_match_oracle(engine, oracle, k):
    import make_golden_exp4 as G4
    cfg, groups, pivots = G4.inputs_of(dict(n_groups=3, genomes_per_group=4, genome_len=9_000, seed=9, out_pivot=True))
    pivots[0] = pivots[0] + b">rep\n" + b"ACGTTGCATTGACCAGTAGGATCCATGCAAGTACCATGGATTTCAGATTACAGATTACAGGGCATCATC" * 300 + b"\n"   # counts above 255
    tables, inters = oracle.exp4(groups, pivots, k)
    engine.group_sets_reset()
    off = [0]
    for grp in groups:
        engine.group_from_fasta(grp, k)
        off.append(engine.group_sets_info()["n_keys"])
    bufs, sizes = [], []
    try:
        for p, text in enumerate(pivots):
            buf, counts, n = engine.kmer_counts(text, k)
            w = 1 if k <= 32 else 2
            keys = buf.download(np.uint64, n * w).reshape((n,) if w == 1 else (n, 2))
            assert np.array_equal(keys, tables[p][0]) and np.array_equal(counts, tables[p][1]), (k, p)
            bufs.append(buf); sizes.append(n)
        assert max(int(t[1].max()) for t in tables) == 255
        masks = engine.group_membership(off, bufs, sizes, k)
    finally:
        for b in bufs:
            b.free()
        engine.group_sets_reset()
    at = 0
    for p in range(len(pivots)):
        ids = oracle._row_ids([tables[p][0]] + [inters[p][d][0] for d in range(len(groups))], k)
        ref = np.zeros(sizes[p], dtype=np.uint64)
        for d in range(len(groups)):
            ref[np.isin(ids[0], ids[1 + d])] |= np.uint64(1 << d)
        assert np.array_equal(masks[at:at + sizes[p], 0], ref), (k, p)
        at += sizes[p]


def test_exp4_rule_chain_matches_reference_merge_lists(engine, tmp_path):
    from khoice_b200 import pipeline4, synth
    import make_golden_exp4 as G4
    cases = json.load(open(os.path.join(GOLDEN, "exp4_cases.json")))["cases"]
    for c, case in enumerate(cases):
        cfg, _, _ = G4.inputs_of(case)
        ks = [str(k) for k in case["k_values"]]
        for mode in ("fused", "rules"):
            if mode == "rules" and c == 1:
                continue
            root = str(tmp_path / f"case{c}_{mode}")
            synth.write_dataset_type4(cfg, root, out_pivot=case["out_pivot"])
            if mode == "fused":
                pipeline4.run_fused(root, case["n_groups"], ks, engine=engine)
            else:
                rep = pipeline4.run_rules(root, case["n_groups"], ks[:2], engine=engine)
                assert rep["jobs_run"] > 0
            for k in (ks if mode == "fused" else ks[:2]):
                for ours, gold in ((f"accuracies_type_4/confusion_matrix/k_{k}_confusion_matrix.txt", "confusion_matrix.txt"),
                                   (f"accuracies_type_4/confusion_matrix/k_{k}_confusion_matrix_with_unidentified.txt", "confusion_matrix_with_unidentified.txt"),
                                   (f"accuracies_type_4/values/k_{k}_accuracy_values.csv", "accuracy_values.csv")):
                    got = open(os.path.join(root, ours), "rb").read()
                    assert got == open(os.path.join(GOLDEN, f"exp4_case{c}_k{k}_{gold}"), "rb").read(), (c, mode, k, ours)
            final = open(os.path.join(root, pipeline4.P_FINAL)).read()
            assert final.count("\n") == case["n_groups"] * len(ks if mode == "fused" else ks[:2])
