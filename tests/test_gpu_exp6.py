"""Experiment type 6 (read-level confusion matrix from simulated reads, exp_type_6.smk + src/merge_lists.py -r) on a GPU:
the per-read votes kernel against a host restatement, and the whole rule chain (fused and rule-by-rule) against fixtures
produced by the reference's own merge_lists.py under a seeded generator (tests/golden/make_golden_exp6.py)."""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)


@pytest.mark.parametrize("k", [9, 21, 31, 40])
def test_read_votes_match_the_dictionary_walk(engine, oracle, k):
    """Engine.read_votes (K2 + sorted lookup + one thread per (read, dataset)) == the reference's loop restated on the host
    over the oracle's tables (merge_lists.votes_from_dump_index), bit for bit in float64."""
    from khoice_b200 import merge_lists
    import make_golden_exp6 as G6
    cfg, groups, reads = G6.inputs_of(dict(n_groups=3, genomes_per_group=3, genome_len=7_000, seed=77, n_reads=50))
    G = len(groups)
    for rt in G6.READ_TYPES:
        tables, inters = oracle.exp4(groups, reads[rt], k)
        engine.group_sets_reset()
        off = [0]
        for grp in groups:
            engine.group_from_fasta(grp, k)
            off.append(engine.group_sets_info()["n_keys"])
        bufs, sizes = [], []
        try:
            for text in reads[rt]:
                buf, _, n = engine.kmer_counts(text, k)
                bufs.append(buf); sizes.append(n)
            masks = engine.group_membership(off, bufs, sizes, k)
            at = 0
            for p in range(G):
                rd = merge_lists.split_reads(reads[rt][p])
                merge_lists.check_reads(rd, k)
                votes, unmatched = engine.read_votes(rd, k, bufs[p], sizes[p], masks[at:at + sizes[p]], G)
                ids = oracle._row_ids([tables[p][0]] + [inters[p][d][0] for d in range(G)], k)
                member = np.stack([np.isin(ids[0], ids[1 + d]) for d in range(G)], axis=1)
                keys = tables[p][0]
                as_int = [int(x) for x in keys] if keys.ndim == 1 else [(int(h) << 64) | int(l) for l, h in keys]
                ref = merge_lists.votes_from_dump_index(rd, k, {x: i for i, x in enumerate(as_int)}, member, G)
                assert votes.shape == ref.shape and np.array_equal(votes, ref), (k, rt, p)
                assert votes.max() > 0 and len(rd) >= 50
                assert int(unmatched.sum()) >= 0
                at += sizes[p]
        finally:
            for b in bufs:
                b.free()
            engine.group_sets_reset()


def test_exp6_rule_chain_matches_reference_merge_lists(engine, tmp_path):
    from khoice_b200 import pipeline6, synth
    import make_golden_exp6 as G6
    cases = json.load(open(os.path.join(GOLDEN, "exp6_cases.json")))["cases"]
    for c, case in enumerate(cases):
        cfg, _, reads = G6.inputs_of(case)
        ks = [str(k) for k in case["k_values"]]
        for mode in ("fused", "rules"):
            if mode == "rules" and c == 0:
                continue
            root = str(tmp_path / f"case{c}_{mode}")
            synth.write_dataset_type6(cfg, root, n_reads=case["n_reads"])
            for rt in G6.READ_TYPES:            # the golden inputs carry three extra reads (empty, short, foreign) in illumina / pivot 1
                for p in range(case["n_groups"]):
                    with open(os.path.join(root, pipeline6.p_reads(rt, p + 1)), "wb") as fd:
                        fd.write(reads[rt][p])
            seed_fn = lambda rt, k, c=c: G6.seed_of(c, rt, k)
            if mode == "fused":
                pipeline6.run_fused(root, case["n_groups"], ks, engine=engine, seed_fn=seed_fn)
            else:
                rep = pipeline6.run_rules(root, case["n_groups"], ks, engine=engine, seed_fn=seed_fn)
                assert rep["jobs_run"] > 0
            for rt in G6.READ_TYPES:
                for k in ks:
                    for ours, gold in ((f"exp6_accuracies/{rt}/confusion_matrix/k_{k}_confusion_matrix.txt", "confusion_matrix.txt"),
                                       (f"exp6_accuracies/{rt}/confusion_matrix/k_{k}_confusion_matrix_with_unidentified.txt", "confusion_matrix.txt"),
                                       (f"exp6_accuracies/{rt}/values/k_{k}_accuracy_values.csv", "accuracy_values.csv")):
                        got = open(os.path.join(root, ours), "rb").read()
                        assert got == open(os.path.join(GOLDEN, f"exp6_case{c}_{rt}_k{k}_{gold}"), "rb").read(), (c, mode, rt, k, ours)
                final = open(os.path.join(root, pipeline6.p_final(1, rt))).read()
                assert final.startswith(pipeline6.HEADER) and final.count("\n") == 1 + case["n_groups"] * len(ks)


def test_feature_level_on_reads_equals_experiment_4_on_the_same_texts(engine, oracle, tmp_path):
    """level="feature" (the reference's rule with its -r line commented out) is experiment 4's matrix with the read files as pivots."""
    from khoice_b200 import merge_lists, pipeline6, synth
    import make_golden_exp6 as G6
    case = G6.CASES[1]
    cfg, groups, reads = G6.inputs_of(case)
    root = str(tmp_path / "feature")
    synth.write_dataset_type6(cfg, root, n_reads=case["n_reads"])
    k = 21
    pipeline6.run_fused(root, case["n_groups"], [k], engine=engine, level="feature")
    G = case["n_groups"]
    for rt in G6.READ_TYPES:
        texts = [open(os.path.join(root, pipeline6.p_reads(rt, p + 1)), "rb").read() for p in range(G)]
        tables, inters = oracle.exp4(groups, texts, k)
        counts, masks = [], []
        for p in range(G):
            ids = oracle._row_ids([tables[p][0]] + [inters[p][d][0] for d in range(G)], k)
            m = np.zeros((tables[p][0].shape[0], 1), dtype=np.uint64)
            for d in range(G):
                m[np.isin(ids[0], ids[1 + d]), 0] |= np.uint64(1 << d)
            counts.append(tables[p][1]); masks.append(m)
        matrix, _ = merge_lists.confusion_from_masks(counts, masks, G)
        got = open(os.path.join(root, f"exp6_accuracies/{rt}/confusion_matrix/k_{k}_confusion_matrix.txt")).read()
        assert got == "".join(",".join(str(x) for x in row) + "\n" for row in matrix)
