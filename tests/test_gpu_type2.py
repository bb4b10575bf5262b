"""Experiment type 2 (pivot analysis, /root/reference/workflow/rules/exp_type_2.smk:297-553) on a GPU: the pivot
kernels, the across stage and the rule chain in fused / rule-by-rule mode against the CPU oracle."""
import filecmp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NB = 64


def _inputs(n_groups=3, per_group=4, length=30_000, seed=11):
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=n_groups, genomes_per_group=per_group, genome_len=length, seed=seed)
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, per_group)] for g in range(1, n_groups + 1)]
    pivots = [synth.make_genome(cfg, g, per_group) for g in range(1, n_groups + 1)]
    return cfg, groups, pivots


def _fold(two):
    """Oracle (sub_hist, inter_hist) -> the engine's single histogram (sub is row 1, inter rows >= 2)."""
    return two[0] + two[1]


@pytest.mark.parametrize("k", [5, 9, 16, 21, 31, 32, 33, 47, 64])
def test_pivot_within_and_across_match_oracle(engine, oracle, k):
    cfg, groups, pivots = _inputs()
    w_ref, a_ref = oracle.exp2(groups, pivots, k, nbins=NB)
    engine.group_sets_reset()
    packs = []
    try:
        for d in range(len(groups)):
            pk = engine.pack_group(groups[d] + [pivots[d]])
            packs.append(pk)
            hist, st = engine.pivot_group_from_packed(pk, k, nbins=NB, keep_sets=True)
            assert np.array_equal(hist, _fold(w_ref[d])), (k, d)
            assert w_ref[d][1][1] == 0 and w_ref[d][0][2:].sum() == 0
        info = engine.pivot_sets_info()
        assert info["n_pivots"] == len(groups)
        assert info["n_pivot_keys"] == sum(int(oracle.genome_set(p, k).shape[0]) for p in pivots)
        hists, _ = engine.pivot_across(nbins=NB)
        for d in range(len(groups)):
            assert np.array_equal(hists[d], _fold(a_ref[d])), (k, d)
    finally:
        for pk in packs:
            pk.free()
        engine.group_sets_reset()


def test_pivot_only_group_and_shared_pivot(engine, oracle):
    """A dataset whose rest of set is empty (everything is `subtract`), and a pivot identical to a rest genome
    (nothing is `subtract`)."""
    cfg, groups, pivots = _inputs(n_groups=2, per_group=3, length=8_000, seed=3)
    groups = [[], groups[1]]
    pivots = [pivots[0], groups[1][0]]
    k = 21
    w_ref, a_ref = oracle.exp2(groups, pivots, k, nbins=NB)
    engine.group_sets_reset()
    packs = []
    try:
        for d in range(2):
            pk = engine.pack_group(groups[d] + [pivots[d]])
            packs.append(pk)
            hist, _ = engine.pivot_group_from_packed(pk, k, nbins=NB)
            assert np.array_equal(hist, _fold(w_ref[d])), d
        assert w_ref[0][1].sum() == 0 and w_ref[1][0].sum() == 0
        hists, _ = engine.pivot_across(nbins=NB)
        for d in range(2):
            assert np.array_equal(hists[d], _fold(a_ref[d])), d
    finally:
        for pk in packs:
            pk.free()
        engine.group_sets_reset()


def test_mixing_plain_and_pivot_groups_is_rejected(engine):
    from khoice_b200.engine import KhbError
    cfg, groups, pivots = _inputs(n_groups=1, per_group=3, length=5_000)
    engine.group_sets_reset()
    engine.group_from_fasta(groups[0], 21)
    pk = engine.pack_group(groups[0] + [pivots[0]])
    try:
        with pytest.raises(KhbError):
            engine.pivot_group_from_packed(pk, 21, nbins=NB, keep_sets=True)
    finally:
        pk.free()
        engine.group_sets_reset()


@pytest.mark.parametrize("k", [13, 31, 40])
def test_sorted_lookup(engine, oracle, k):
    rng = np.random.default_rng(k)
    cfg, groups, pivots = _inputs(n_groups=1, per_group=3, length=20_000, seed=k)
    a = oracle.genome_set(pivots[0], k)
    b = oracle.genome_set(groups[0][0], k)
    idx = engine.sorted_lookup(a, b, k)
    ia, ib = oracle._row_ids([a, b], k)
    pos = np.searchsorted(ib, ia)
    hit = (pos < ib.shape[0]) & (ib[np.minimum(pos, ib.shape[0] - 1)] == ia)
    assert np.array_equal(idx >= 0, hit)
    assert np.array_equal(idx[hit], pos[hit])
    assert engine.sorted_lookup(a[:0], b, k).shape[0] == 0
    assert (engine.sorted_lookup(a, b[:0], k) == -1).all()


K_VALUES = ["9", "21", "31", "34"]


def test_type2_rule_chain_fused_and_rules(engine, oracle, tmp_path):
    from khoice_b200 import kmcdb, pipeline2, synth, tables
    cfg, groups, pivots = _inputs(n_groups=3, per_group=4, length=25_000, seed=21)
    roots = {}
    for mode in ("fused", "rules", "rules-subprocess"):
        roots[mode] = str(tmp_path / mode)
        synth.write_dataset_type2(cfg, roots[mode])
    pipeline2.run_fused(roots["fused"], cfg.n_groups, K_VALUES, engine=engine)
    pipeline2.run_rules(roots["rules"], cfg.n_groups, K_VALUES, engine=engine)
    pipeline2.run_rules(roots["rules-subprocess"], cfg.n_groups, K_VALUES[:1], subprocess_mode=True)
    for k in K_VALUES:
        w_ref, a_ref = oracle.exp2(groups, pivots, int(k), nbins=tables.HIST_ROWS)
        for mode in ("fused", "rules") + (("rules-subprocess",) if k in K_VALUES[:1] else ()):
            for num in range(1, cfg.n_groups + 1):
                for scope, ref, fn in (("within", w_ref, pipeline2.p_within), ("across", a_ref, pipeline2.p_across)):
                    for j, op in enumerate(pipeline2.OPS):
                        got = tables.read_histogram_file(os.path.join(roots[mode], fn(k, num, op) + ".hist.txt"))
                        assert len(got) == 5000
                        assert got == [int(x) for x in ref[num - 1][j][1:]], (mode, scope, k, num, op)
    for f in (pipeline2.P_WITHIN_CSV, pipeline2.P_ACROSS_CSV):
        assert filecmp.cmp(os.path.join(roots["fused"], f), os.path.join(roots["rules"], f), shallow=False), f
        assert open(os.path.join(roots["fused"], f)).read().count("\n") == 1 + cfg.n_groups * len(K_VALUES)
    # the rule-by-rule databases hold the oracle's sets
    k = 21
    pset = oracle.genome_set(pivots[0], k)
    ukeys, ucnt = oracle.union_sum([oracle.genome_set(g, k) for g in groups[0]], k)
    ikeys, icnt = oracle.simple_intersect_ocsum(pset, np.ones(pset.shape[0], np.uint32), ukeys, ucnt, k)
    skeys, _ = oracle.simple_kmers_subtract(pset, np.ones(pset.shape[0], np.uint32), ukeys, k)
    db = kmcdb.read_db(os.path.join(roots["rules"], pipeline2.p_within(str(k), 1, "intersect")))
    assert np.array_equal(db.keys, ikeys) and np.array_equal(db.counts, icnt)
    db = kmcdb.read_db(os.path.join(roots["rules"], pipeline2.p_within(str(k), 1, "subtract")))
    assert np.array_equal(db.keys, skeys) and (db.counts == 1).all()
    # resume: a second run does nothing; every declared output of the fused mode exists
    assert pipeline2.run_rules(roots["rules"], cfg.n_groups, K_VALUES, engine=engine)["jobs_run"] == 0
    for rule, outputs, _ in pipeline2._rule_jobs(roots["fused"], K_VALUES, cfg.n_groups):
        for o in outputs:
            assert os.path.exists(os.path.join(roots["fused"], o)), (rule, o)
