"""Experiment type 4, host side (no GPU): khoice_b200.merge_lists against the outputs of the reference's own
src/merge_lists.py (tests/golden/make_golden_exp4.py) on dumps the CPU oracle writes for the same synthetic genomes."""
import json
import os
import sys

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)


def _golden(c, k, name):
    return open(os.path.join(GOLDEN, f"exp4_case{c}_k{k}_{name}"), "rb").read()


def test_merge_lists_cli_matches_reference_program(oracle, tmp_path):
    from khoice_b200 import kmcdb, merge_lists
    import make_golden_exp4 as G4
    cases = json.load(open(os.path.join(GOLDEN, "exp4_cases.json")))["cases"]
    assert cases == G4.CASES
    for c, case in enumerate(cases):
        cfg, groups, pivots = G4.inputs_of(case)
        G = case["n_groups"]
        for k in case["k_values"][:2]:
            work = tmp_path / f"c{c}_k{k}"
            work.mkdir()
            tables, inters = oracle.exp4(groups, pivots, k)
            pl, il = [], []
            for p in range(G):
                f = str(work / f"pivot_{p + 1}.txt")
                kmcdb.write_text_dump(f, tables[p][0], tables[p][1], k)
                pl.append(f)
                for d in range(G):
                    f = str(work / f"p{p + 1}_d{d + 1}.txt")
                    kmcdb.write_text_dump(f, inters[p][d][0], inters[p][d][1], k)
                    il.append(f)
            (work / "pl.txt").write_text("\n".join(pl) + "\n")
            (work / "il.txt").write_text("\n".join(il) + "\n")
            out = str(work / "out") + "/"
            assert merge_lists.main(["-p", str(work / "pl.txt"), "-i", str(work / "il.txt"), "-o", out, "-n", str(G), "-k", str(k)]) == 0
            assert open(out + f"confusion_matrix/k_{k}_confusion_matrix.txt", "rb").read() == _golden(c, k, "confusion_matrix.txt")
            assert open(out + f"confusion_matrix/k_{k}_confusion_matrix_with_unidentified.txt", "rb").read() == _golden(c, k, "confusion_matrix_with_unidentified.txt")
            assert open(out + f"values/k_{k}_accuracy_values.csv", "rb").read() == _golden(c, k, "accuracy_values.csv")
            # the mask front end (what the GPU path feeds) gives the same rows as the dump front end
            counts, masks = [], []
            for p in range(G):
                pk = tables[p][0]
                ids = oracle._row_ids([pk] + [inters[p][d][0] for d in range(G)], k)
                m = np.zeros((pk.shape[0], 1), dtype=np.uint64)
                for d in range(G):
                    m[np.isin(ids[0], ids[1 + d]), 0] |= np.uint64(1 << d)
                counts.append(tables[p][1])
                masks.append(m)
            a = merge_lists.confusion_from_masks(counts, masks, G)
            b = merge_lists.confusion_from_dumps(pl, il, G)
            assert a == b
