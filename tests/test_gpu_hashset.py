"""The sort-free group stage (csrc/hashset.cu): K2 fused with an open-addressing table of (k-mer, genome bit set)
records.  It must give exactly what the single-sort path, the KMC-shaped two-sort chain and the CPU oracle give
(reference call sites: exp_type_1.smk:156-191; pivot mode exp_type_2.smk:354-380), for every record width, across
table reuse, and when it hands a group back to the sort path."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import EDGE_FASTAS, random_fasta, sort_rows

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_groups(engine, groups, k, nbins=256):
    engine.group_sets_reset()
    hists, stats = [], []
    for grp in groups:
        h, st = engine.group_from_fasta(grp, k, nbins=nbins)
        hists.append(h)
        stats.append(st)
    sets = engine.group_sets_download()
    ha, sta = engine.across_groups(nbins=nbins)
    return hists, stats, sets, ha, sta


def _split_sorted(sets, stats):
    out, off = [], 0
    for st in stats:
        out.append(sort_rows(sets[off:off + st["distinct"]]))
        off += st["distinct"]
    return out


@pytest.mark.parametrize("k", [11, 16, 21, 27, 31])
@pytest.mark.parametrize("n_genomes", [3, 40, 70, 200])  # record widths 4, 4, 8 and 16 words
def test_hash_mode_equals_sort_modes_and_oracle(engine, oracle, k, n_genomes):
    from khoice_b200 import synth
    glen = 30_000 if n_genomes <= 40 else 8_000
    cfg = synth.SynthConfig(n_groups=2, genomes_per_group=n_genomes, genome_len=glen, seed=1234 + n_genomes)
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, n_genomes + 1)] for g in (1, 2)]
    groups[0][1] = groups[0][1] + EDGE_FASTAS[1] + EDGE_FASTAS[3] + EDGE_FASTAS[8]
    groups[1].append(b"")
    flat = [f for grp in groups for f in grp]
    gid = [i for i, grp in enumerate(groups) for _ in grp]
    w_ref, a_ref, st_ref = oracle.exp1(flat, gid, 2, k, nbins=256)
    results = {}
    try:
        for mode in ("hash", "single-sort"):
            engine.set_group_mode(mode)
            results[mode] = _run_groups(engine, groups, k)
    finally:
        engine.set_group_mode("auto")
    assert engine.hash_overflows == 0
    for mode, (hists, stats, sets, ha, sta) in results.items():
        for i in range(2):
            assert np.array_equal(hists[i], w_ref[i]), (mode, i)
        assert np.array_equal(ha, a_ref), mode
        assert sta["distinct"] == st_ref["distinct"]
        assert sum(st["genome_distinct"] for st in stats) == st_ref["sum_genome_distinct"], mode
        assert sum(st["distinct"] for st in stats) == st_ref["sum_group_distinct"], mode
    a, b = results["hash"], results["single-sort"]
    assert results["hash"][1][0]["passes_group"] == 0          # no radix pass at all
    assert results["single-sort"][1][0]["passes_group"] >= 1
    for x, y in zip(_split_sorted(a[2], a[1]), _split_sorted(b[2], b[1])):
        assert np.array_equal(x, y)


def test_table_is_clean_between_groups_of_different_sizes(engine, oracle):
    """The count kernel zeroes the records it reads; a larger, a smaller and again a larger group through the same table."""
    rng = np.random.default_rng(5)
    k = 23
    sizes = [(6, 120_000), (2, 3_000), (5, 60_000), (1, 10), (6, 120_000)]
    groups = [[random_fasta(rng, n, n_records=2) for _ in range(g)] for g, n in sizes]
    groups[4] = groups[0]
    try:
        engine.set_group_mode("hash")
        hists, stats, sets, ha, _ = _run_groups(engine, groups, k)
    finally:
        engine.set_group_mode("auto")
    flat = [f for grp in groups for f in grp]
    gid = [i for i, grp in enumerate(groups) for _ in grp]
    w_ref, a_ref, _ = oracle.exp1(flat, gid, len(groups), k, nbins=256)
    for i in range(len(groups)):
        assert np.array_equal(hists[i], w_ref[i]), i
    assert np.array_equal(hists[0], hists[4])
    assert np.array_equal(ha, a_ref)


def test_repeats_inside_a_genome_and_identical_genomes(engine, oracle):
    """Many windows of one genome hit the same record (the bit is set once), identical genomes share every record."""
    rep = b">rep\n" + b"ACGTTGCATTGACCAGTAGGATCCATGCAAGT" * 300 + b"\n"
    rng = np.random.default_rng(11)
    base = random_fasta(rng, 50_000)
    groups = [[rep, base, base, base + rep, rep + rep], [b">a\n" + b"A" * 5000 + b"\n", b">t\n" + b"T" * 5000 + b"\n"]]
    for k in (13, 31):
        try:
            engine.set_group_mode("hash")
            hists, stats, sets, ha, _ = _run_groups(engine, groups, k)
        finally:
            engine.set_group_mode("auto")
        flat = [f for grp in groups for f in grp]
        w_ref, a_ref, _ = oracle.exp1(flat, [0] * 5 + [1] * 2, 2, k, nbins=256)
        assert np.array_equal(hists[0], w_ref[0]) and np.array_equal(hists[1], w_ref[1])
        assert hists[1][2] == 1 and hists[1].sum() == 1     # A^k and T^k are one canonical k-mer, in both genomes
        assert np.array_equal(ha, a_ref)


SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, %(root)r)
from khoice_b200 import synth
from khoice_b200.engine import Engine
from oracle import oracle as O
cfg = synth.SynthConfig(n_groups=2, genomes_per_group=4, genome_len=50_000, seed=99)
groups = [[synth.make_genome(cfg, g, i) for i in range(1, 5)] for g in (1, 2)]
flat = [f for g in groups for f in g]
eng = Engine(0)
for k in (15, 31):
    w_ref, a_ref, _ = O.exp1(flat, [0] * 4 + [1] * 4, 2, k, nbins=64)
    eng.group_sets_reset()
    for d in range(2):
        hist, st = eng.group_from_fasta(groups[d], k, nbins=64)
        assert np.array_equal(hist, w_ref[d]), ("within", k, d)
        assert st["passes_group"] >= 1, st        # the sort path finished the group
    hist, _ = eng.across_groups(nbins=64)
    assert np.array_equal(hist, a_ref), ("across", k)
assert eng.hash_overflows == 4, eng.hash_overflows
eng.close()
print("fallback ok")
"""


def test_probe_limit_hands_the_group_back_to_the_sort_path():
    """KHB_HASH_MAX_PROBE=0: the first collision raises the overflow flag; the group is redone by sorting, the dirty
    table is cleared before its next use, results stay exact."""
    env = dict(os.environ, KHB_HASH_MAX_PROBE="0", KHB_GROUP_MODE="hash")
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "fallback ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
