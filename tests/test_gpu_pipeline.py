"""Fused per-group / across-group stages against the CPU oracle (bit-exact histograms and sets)."""
import numpy as np
import pytest

from helpers import EDGE_FASTAS, random_fasta, sort_rows

pytestmark = pytest.mark.gpu


def _oracle_group_sets(oracle, groups, k):
    sets = []
    for genomes in groups:
        keys, _ = oracle.union_sum([oracle.genome_set(g, k) for g in genomes], k)
        sets.append(keys)
    return sets


@pytest.mark.parametrize("k", [3, 5, 13, 21, 31, 32, 33, 47, 63])
def test_small_groups_match_oracle(engine, oracle, k):
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=3, genomes_per_group=4, genome_len=60_000, seed=77)
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, 5)] for g in range(1, 4)]
    groups[1].append(EDGE_FASTAS[1] + EDGE_FASTAS[3])   # ragged extra genome with N runs / IUPAC
    groups[2].append(b"")                               # an empty file is a legal (empty) genome
    flat = [g for grp in groups for g in grp]
    gid = [i for i, grp in enumerate(groups) for _ in grp]
    w_ref, a_ref, st_ref = oracle.exp1(flat, gid, len(groups), k)
    engine.group_sets_reset()
    tot_bases, sizes = 0, []
    for i, grp in enumerate(groups):
        hist, st = engine.group_from_fasta(grp, k)
        assert np.array_equal(hist, w_ref[i]), f"group {i} within-group histogram"
        tot_bases += st["bases"]
        sizes.append(st["distinct"])
    assert tot_bases == st_ref["symbols"]
    info = engine.group_sets_info()
    assert info["n_groups"] == len(groups) and info["k"] == k
    assert info["n_keys"] == st_ref["sum_group_distinct"]
    sets = engine.group_sets_download()      # canonical values, group after group, prefix order inside a group
    ref_sets = _oracle_group_sets(oracle, groups, k)
    off = 0
    for i, ref in enumerate(ref_sets):
        assert sizes[i] == ref.shape[0]
        assert np.array_equal(sort_rows(sets[off:off + sizes[i]]), ref), f"group {i} distinct k-mer set"
        off += sizes[i]
    hist, st = engine.across_groups()
    assert np.array_equal(hist, a_ref)
    assert st["distinct"] == st_ref["distinct"]


def test_staged_equals_host_path(engine, oracle):
    rng = np.random.default_rng(9)
    files = [random_fasta(rng, 50_000) for _ in range(5)]
    engine.group_sets_reset()
    h1, s1 = engine.group_from_fasta(files, 31, keep_set=False)
    st = engine.stage_fasta(files)
    h2, s2 = engine.group_from_staged(st, 31, keep_set=False)
    assert np.array_equal(h1, h2) and s1["distinct"] == s2["distinct"]
    # sub-range of a staged buffer = that subset of genomes
    h3, _ = engine.group_from_staged(st, 31, keep_set=False, first=1, count=3)
    h4, _ = engine.group_from_fasta(files[1:4], 31, keep_set=False)
    assert np.array_equal(h3, h4)


def test_config1_full_size_matches_oracle(engine, oracle):
    """BASELINE config 1: 2 groups x 5 genomes x 5 Mbp, k=31 -- oracle finishes in seconds."""
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=2, genomes_per_group=5, genome_len=5_000_000)
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, 6)] for g in range(1, 3)]
    flat = [g for grp in groups for g in grp]
    gid = [i for i, grp in enumerate(groups) for _ in grp]
    w_ref, a_ref, st_ref = oracle.exp1(flat, gid, 2, 31)
    engine.group_sets_reset()
    for i, grp in enumerate(groups):
        hist, st = engine.group_from_fasta(grp, 31)
        assert np.array_equal(hist, w_ref[i])
    hist, st = engine.across_groups()
    assert np.array_equal(hist, a_ref)
    # size-independent properties: every distinct k-mer is counted once; occupancy <= members
    assert int(hist.sum()) == st_ref["distinct"] and not hist[3:].any()


def test_prefetch_double_buffering_gives_identical_results(engine):
    rng = np.random.default_rng(12)
    groups = [[random_fasta(rng, 30_000 + 1000 * j) for j in range(3)] for _ in range(4)]
    ref = []
    engine.group_sets_reset()
    for grp in groups:
        ref.append(engine.group_from_fasta(grp, 31)[0])
    a_ref = engine.across_groups()[0]
    engine.group_sets_reset()
    views = [[np.frombuffer(f, dtype=np.uint8) for f in grp] for grp in groups]
    engine.prefetch_fasta(views[0])
    got = []
    for i, grp in enumerate(views):
        if i + 1 < len(views):
            engine.prefetch_fasta(views[i + 1])       # deferred until the pending buffer has been taken over
        got.append(engine.group_from_fasta(grp, 31)[0])
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    assert np.array_equal(engine.across_groups()[0], a_ref)
    # a prefetch that is never consumed must not leak into the next call
    engine.group_sets_reset()
    engine.prefetch_fasta(views[2])
    assert np.array_equal(engine.group_from_fasta(views[1], 31, keep_set=False)[0], ref[1])


def test_pack_once_sweep_k_equals_per_k_pipeline(engine):
    """khb_pack_group + khb_group_from_packed (pack once, sweep k) == khb_group_from_fasta for every k."""
    rng = np.random.default_rng(21)
    groups = [[random_fasta(rng, 40_000 + 777 * j, p_n=0.003) for j in range(3)] + [EDGE_FASTAS[2]] for _ in range(2)]
    packed = [engine.pack_group(g) for g in groups]
    info = packed[0].info()
    assert info["bases"] > 100_000 and info["device_bytes"] < info["n_symbols"]  # 3/8 byte per symbol
    for k in (9, 31, 45, 64):
        engine.group_sets_reset()
        ref = [engine.group_from_fasta(g, k)[0] for g in groups]
        a_ref = engine.across_groups()[0]
        engine.group_sets_reset()
        got = [engine.group_from_packed(p, k)[0] for p in packed]
        a_got = engine.across_groups()[0]
        for a, b in zip(ref, got):
            assert np.array_equal(a, b), k
        assert np.array_equal(a_ref, a_got), k
    for p in packed:
        p.free()


@pytest.mark.parametrize("k", [4, 5, 6, 7, 8, 9, 10])
def test_small_k_direct_address_path_and_its_switch_over(engine, oracle, k):
    """k small enough for the presence table (4^k x ceil(N/32) words <= windows / 4) skips the sort (presence.cu); the
    sweep crosses the switch-over, and both sides must give the oracle's histograms, set sizes and sets."""
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=2, genomes_per_group=3, genome_len=200_000, seed=31)
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, 4)] for g in (1, 2)]
    flat = [f for grp in groups for f in grp]
    w_ref, a_ref, st_ref = oracle.exp1(flat, [0, 0, 0, 1, 1, 1], 2, k, nbins=32)
    engine.group_sets_reset()
    total = 0
    used_table = []
    for d, grp in enumerate(groups):
        hist, st = engine.group_from_fasta(grp, k, nbins=32)
        assert np.array_equal(hist, w_ref[d]), (k, d)
        total += st["genome_distinct"]
        used_table.append(st["passes_group"] == 0)
    assert total == st_ref["sum_genome_distinct"]
    assert all(used_table) == (k <= 8), (k, used_table)     # 600 k windows: 4^8 words x 4 <= windows < 4^9 words x 4
    got = engine.group_sets_download()
    ref = np.concatenate([oracle.union_sum([oracle.genome_set(g, k) for g in grp], k)[0] for grp in groups])
    assert np.array_equal(np.sort(got), np.sort(ref))
    hist, st = engine.across_groups(nbins=32)
    assert np.array_equal(hist, a_ref) and st["distinct"] == st_ref["distinct"]
    engine.group_sets_reset()


def test_two_contexts_in_one_process(engine, oracle):
    """include/khoice_b200.h promises re-entrancy across contexts: a second context (on another GPU when the box has one --
    kernel attributes such as the dynamic shared-memory opt-in are per DEVICE -- else on the same GPU) computes the same
    histograms, 64- and 128-bit keys, while the first context stays usable."""
    from khoice_b200.engine import Engine
    rng = np.random.default_rng(77)
    files = [random_fasta(rng, 60_000) for _ in range(4)]
    second = None
    for dev in (1, 0):
        try:
            second = Engine(dev)
            break
        except Exception:
            continue
    assert second is not None
    try:
        for k in (31, 47):
            w_ref, _, _ = oracle.exp1(files, [0] * 4, 1, k)
            engine.group_sets_reset()
            second.group_sets_reset()
            h2, _ = second.group_from_fasta(files, k, keep_set=False)
            h1, _ = engine.group_from_fasta(files, k, keep_set=False)
            assert np.array_equal(h1, w_ref[0]) and np.array_equal(h2, w_ref[0])
    finally:
        second.close()
