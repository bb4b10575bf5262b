"""The minimizer-bin group stage (csrc/bins.cu, the default group path for 17 <= k <= 63, k != 32): super-k-mer records partitioned by
minimizer, every bin counted by one CTA in a shared-memory table.  It must give exactly what the single-sort path and the CPU
oracle give (reference call sites: /root/reference/workflow/rules/exp_type_1.smk:156-191 within a group, :233-259 across groups)
-- histograms, group sets, per-genome totals -- for every table variant (<= 64 genomes, chunks of 64 genomes), for records
split at their length limit, and when a bin outgrows its table (hash classes) or its region (the group is redone by sorting)."""
import os

import numpy as np
import pytest

from helpers import EDGE_FASTAS, random_fasta, sort_rows

pytestmark = pytest.mark.gpu


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


def _both_modes(engine, groups, k, nbins=256):
    res = {}
    try:
        for mode in ("bins", "single-sort"):
            engine.set_group_mode(mode)
            res[mode] = _run_groups(engine, groups, k, nbins)
    finally:
        engine.set_group_mode("auto")
    return res


def _check(res, w_ref, a_ref, st_ref, n_groups):
    for mode, (hists, stats, sets, ha, sta) in res.items():
        for i in range(n_groups):
            assert np.array_equal(hists[i], w_ref[i]), (mode, i)
        assert np.array_equal(ha, a_ref), mode
        assert sta["distinct"] == st_ref["distinct"], mode
        assert sum(st["genome_distinct"] for st in stats) == st_ref["sum_genome_distinct"], mode
        assert sum(st["distinct"] for st in stats) == st_ref["sum_group_distinct"], mode
    a, b = res["bins"], res["single-sort"]
    for x, y in zip(_split_sorted(a[2], a[1]), _split_sorted(b[2], b[1])):
        assert np.array_equal(x, y)


def test_across_stage_bin_by_bin_and_by_sort(oracle, monkeypatch):
    """The opt-in across-group stage over the segments the bins left in the store (KHB_ACROSS_MODE=bins, no sort) against the prefix sort
    of the same store and the oracle, for both key widths; tiny tables force hash classes; a group from another path in between makes
    the stage sort.  The switch is read once per process, so this test runs in a process of its own."""
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent("""
        import os, sys
        import numpy as np
        sys.path.insert(0, os.getcwd())
        from khoice_b200 import synth
        from khoice_b200.engine import Engine
        from oracle import oracle as O
        engine = Engine(0)
        def run_groups(groups, k):
            engine.group_sets_reset()
            hists = [engine.group_from_fasta(g, k, nbins=64)[0] for g in groups]
            ha, sta = engine.across_groups(nbins=64)
            return hists, ha, sta
        for k in (31, 45):
            cfg = synth.SynthConfig(n_groups=5, genomes_per_group=4, genome_len=25_000, seed=77 + k)
            groups = [[synth.make_genome(cfg, g, i) for i in range(1, 5)] for g in range(1, 6)]
            flat = [f for grp in groups for f in grp]
            gid = [i for i, grp in enumerate(groups) for _ in grp]
            w_ref, a_ref, st_ref = O.exp1(flat, gid, 5, k, nbins=64)
            for slots in (None, "8"):
                if slots:
                    os.environ["KHB_ACROSS_SLOTS_LOG2"] = slots
                before = engine.bins_counters
                hists, ha, sta = run_groups(groups, k)
                after = engine.bins_counters
                assert after["across_by_bins"] == before["across_by_bins"] + 1 and after["across_by_sort"] == before["across_by_sort"], (before, after)
                assert np.array_equal(ha, a_ref) and sta["distinct"] == st_ref["distinct"]
                assert all(np.array_equal(hists[i], w_ref[i]) for i in range(5))
            del os.environ["KHB_ACROSS_SLOTS_LOG2"]
            engine.group_sets_reset()
            for i, grp in enumerate(groups):
                engine.set_group_mode("single-sort" if i == 2 else "auto")
                engine.group_from_fasta(grp, k, nbins=64)
            engine.set_group_mode("auto")
            before = engine.bins_counters
            ha, sta = engine.across_groups(nbins=64)
            assert engine.bins_counters["across_by_sort"] == before["across_by_sort"] + 1
            assert np.array_equal(ha, a_ref) and sta["distinct"] == st_ref["distinct"]
        print("ok")
    """)
    env = dict(os.environ, KHB_ACROSS_MODE="bins")
    r = subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("k", [17, 18, 21, 24, 27, 30, 31, 33, 40, 47, 56, 63])   # 64-bit words up to 31, 128-bit from 33
@pytest.mark.parametrize("n_genomes", [3, 40, 70, 200])  # one chunk of genome bits, one, two and four chunks
def test_bins_mode_equals_sort_mode_and_oracle(engine, oracle, k, n_genomes):
    from khoice_b200 import synth
    glen = 30_000 if n_genomes <= 40 else 8_000
    cfg = synth.SynthConfig(n_groups=2, genomes_per_group=n_genomes, genome_len=glen, seed=4321 + n_genomes)
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, n_genomes + 1)] for g in (1, 2)]
    groups[0][1] = groups[0][1] + EDGE_FASTAS[1] + EDGE_FASTAS[3] + EDGE_FASTAS[8]
    groups[1].append(b"")
    flat = [f for grp in groups for f in grp]
    gid = [i for i, grp in enumerate(groups) for _ in grp]
    w_ref, a_ref, st_ref = oracle.exp1(flat, gid, 2, k, nbins=256)
    before = engine.bins_counters
    res = _both_modes(engine, groups, k)
    _check(res, w_ref, a_ref, st_ref, 2)
    assert res["bins"][1][0]["passes_group"] == 0            # no radix pass in the group stage
    assert res["single-sort"][1][0]["passes_group"] >= 1
    assert engine.bins_counters["fallbacks"] == before["fallbacks"]


@pytest.mark.parametrize("k", [19, 45])
def test_edge_inputs(engine, oracle, k):
    """Every edge FASTA of the suite as its own genome; empty genomes; a group without any k-mer."""
    groups = [list(EDGE_FASTAS), [b"", b">only header\n", b"ACGT\n"], [EDGE_FASTAS[8]] * 3]
    flat = [f for grp in groups for f in grp]
    gid = [i for i, grp in enumerate(groups) for _ in grp]
    w_ref, a_ref, st_ref = oracle.exp1(flat, gid, 3, k, nbins=64)
    res = _both_modes(engine, groups, k, nbins=64)
    _check(res, w_ref, a_ref, st_ref, 3)


@pytest.mark.parametrize("k", [25, 51])
def test_long_runs_are_split_into_several_records(engine, oracle, k):
    """Low-complexity sequence: thousands of consecutive windows share one minimizer, so a run is cut into records of at most
    65 - k (<= 32) windows and into pieces at tile boundaries; homopolymers and short tandem repeats also put one k-mer into a
    bin thousands of times."""
    rng = np.random.default_rng(11)
    parts = [b">a\n" + b"A" * 30_000 + b"\n", b">b\n" + b"AC" * 9_000 + b"\n", b">c\n" + b"ACGGT" * 5_000 + b"\n",
             random_fasta(rng, 20_000), b">d\n" + b"T" * 10_000 + b"G" + b"T" * 10_000 + b"\n"]
    groups = [[b"".join(parts), parts[0] + parts[3], random_fasta(rng, 5_000)], [parts[1] + parts[2], parts[4]]]
    flat = [f for grp in groups for f in grp]
    gid = [i for i, grp in enumerate(groups) for _ in grp]
    w_ref, a_ref, st_ref = oracle.exp1(flat, gid, 2, k, nbins=64)
    res = _both_modes(engine, groups, k, nbins=64)
    _check(res, w_ref, a_ref, st_ref, 2)


@pytest.mark.parametrize("k", [23, 37])
def test_full_tables_are_redone_in_hash_classes(engine, oracle, monkeypatch, k):
    """Tiny tables (256 slots) and a planner told that records hold next to no distinct k-mers: every bin is tried in one pass,
    overflows its table and is redone by mb_bigbin_kernel, class by class.  Then the same tables with the planner's own estimate:
    the bins are counted in several hash classes inside the streaming kernel.  Same results either way."""
    rng = np.random.default_rng(12)
    groups = [[random_fasta(rng, 150_000, n_records=2) for _ in range(3)], [random_fasta(rng, 40_000) for _ in range(70)]]
    flat = [f for grp in groups for f in grp]
    gid = [i for i, grp in enumerate(groups) for _ in grp]
    w_ref, a_ref, st_ref = oracle.exp1(flat, gid, 2, k, nbins=128)
    monkeypatch.setenv("KHB_BINS_SLOTS_LOG2", "8")
    monkeypatch.setenv("KHB_BINS_RHO_PCT", "5")
    before = engine.bins_counters
    res = _both_modes(engine, groups, k, nbins=128)
    _check(res, w_ref, a_ref, st_ref, 2)
    after = engine.bins_counters
    assert after["big_bins"] > before["big_bins"]
    assert after["fallbacks"] == before["fallbacks"]
    monkeypatch.delenv("KHB_BINS_RHO_PCT")
    res = _both_modes(engine, groups, k, nbins=128)
    _check(res, w_ref, a_ref, st_ref, 2)
    assert engine.bins_counters["fallbacks"] == before["fallbacks"]


def test_region_overflow_is_partitioned_again_with_exact_sizes(engine, oracle, monkeypatch):
    """Bin regions sized at 30 % of the expected records: the partition raises its flag and the group is partitioned a second time into
    regions of exactly the sizes the first attempt counted -- still without a sort."""
    rng = np.random.default_rng(13)
    k = 29
    groups = [[random_fasta(rng, 100_000) for _ in range(4)]]
    w_ref, a_ref, st_ref = oracle.exp1(groups[0], [0] * 4, 1, k, nbins=32)
    monkeypatch.setenv("KHB_BINS_SLACK_PCT", "30")
    before = engine.bins_counters
    try:
        engine.set_group_mode("bins")
        hists, stats, sets, ha, sta = _run_groups(engine, groups, k, nbins=32)
    finally:
        engine.set_group_mode("auto")
    assert engine.bins_counters["repartitions"] == before["repartitions"] + 1
    assert engine.bins_counters["fallbacks"] == before["fallbacks"]
    assert stats[0]["passes_group"] == 0
    assert np.array_equal(hists[0], w_ref[0])
    assert np.array_equal(ha, a_ref)
    assert stats[0]["distinct"] == st_ref["sum_group_distinct"]


def test_tables_are_clean_between_groups(engine, oracle):
    """A larger, a tiny, and again the larger group through the same context: nothing of a group may survive in the scratch."""
    rng = np.random.default_rng(14)
    k = 21
    big = [random_fasta(rng, 120_000, n_records=2) for _ in range(6)]
    groups = [big, [random_fasta(rng, 30)], [random_fasta(rng, 60_000) for _ in range(5)], big]
    flat = [f for grp in groups for f in grp]
    gid = [i for i, grp in enumerate(groups) for _ in grp]
    w_ref, a_ref, st_ref = oracle.exp1(flat, gid, 4, k, nbins=32)
    res = _both_modes(engine, groups, k, nbins=32)
    _check(res, w_ref, a_ref, st_ref, 4)
    assert np.array_equal(res["bins"][0][0], res["bins"][0][3])


@pytest.mark.parametrize("k", [31, 47])
def test_config2_sized_group(engine, oracle, k):
    """One config-2 group at full size (50 x 5 Mbp) through the default mode against the CPU oracle."""
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=1, genomes_per_group=50, genome_len=5_000_000, seed=20240131)
    grp = [synth.make_genome(cfg, 1, i) for i in range(1, 51)]
    w_ref, a_ref, st_ref = oracle.exp1(grp, [0] * 50, 1, k, nbins=5000)
    engine.group_sets_reset()
    before = engine.bins_counters
    h, st = engine.group_from_fasta(grp, k, nbins=5000)
    assert st["passes_group"] == 0
    assert engine.bins_counters["fallbacks"] == before["fallbacks"]
    assert np.array_equal(h, w_ref[0])
    assert st["distinct"] == st_ref["sum_group_distinct"]
    ha, sta = engine.across_groups(nbins=5000)
    assert np.array_equal(ha, a_ref)


@pytest.mark.parametrize("k,n_genomes,n_bins", [(31, 5, 97), (17, 3, 16), (24, 70, 64), (47, 4, 33), (63, 66, 50)])
def test_partition_against_its_numpy_statement(engine, k, n_genomes, n_bins):
    """The first pass alone: per region (bin, chunk of 64 genomes) the records and the windows the kernel wrote, against oracle/superkmer.py
    (minimizer hash, bin function, runs of one minimum, cuts at 4096-window tiles and at 32 windows) -- bit for bit."""
    from khoice_b200 import synth
    from oracle import superkmer as S
    cfg = synth.SynthConfig(n_groups=1, genomes_per_group=n_genomes, genome_len=12_000, seed=900 + k)
    genomes = [synth.make_genome(cfg, 1, i) for i in range(1, n_genomes + 1)]
    genomes[1] = genomes[1] + EDGE_FASTAS[1] + EDGE_FASTAS[3] + EDGE_FASTAS[8] + b">poly\n" + b"A" * 9000 + b"\n"
    genomes[-1] = b""
    rec, win, first = engine.bins_partition(genomes, k, n_bins)
    rec_ref, win_ref = S.region_counts(genomes, first, k, n_bins)
    assert np.array_equal(win.astype(np.int64), win_ref)
    assert np.array_equal(rec.astype(np.int64), rec_ref)
    assert int(win.sum()) > 0
