"""Property tests on the GPU: random ragged FASTA-like text (any bytes from a hostile alphabet, any k) must give the
oracle's histograms through the fused path, and the size-independent invariants of the domain must hold."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu

TEXT = st.text(alphabet="ACGTacgtNnRY>\n\r ;-", min_size=0, max_size=400)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(st.lists(st.lists(TEXT, min_size=1, max_size=4), min_size=1, max_size=3), st.integers(min_value=1, max_value=64))
def test_random_text_groups_match_oracle(engine, oracle, groups, k):
    bgroups = [[t.encode() for t in grp] for grp in groups]
    flat = [g for grp in bgroups for g in grp]
    gid = [i for i, grp in enumerate(bgroups) for _ in grp]
    w_ref, a_ref, st_ref = oracle.exp1(flat, gid, len(bgroups), k)
    engine.group_sets_reset()
    for i, grp in enumerate(bgroups):
        hist, stats = engine.group_from_fasta(grp, k)
        assert np.array_equal(hist, w_ref[i])
    hist, stats = engine.across_groups()
    assert np.array_equal(hist, a_ref)


@pytest.mark.parametrize("k", list(range(7, 32, 2)))
def test_k_sweep_config3_shape(engine, oracle, k):
    """BASELINE config 3 (k = 7, 9, ..., 31) on a small 3 x 4 set: 64-bit packing and prefix planning per k."""
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=3, genomes_per_group=4, genome_len=50_000, seed=31)
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, 5)] for g in range(1, 4)]
    flat = [f for grp in groups for f in grp]
    w_ref, a_ref, _ = oracle.exp1(flat, [0] * 4 + [1] * 4 + [2] * 4, 3, k)
    engine.group_sets_reset()
    for i, grp in enumerate(groups):
        hist, st_ = engine.group_from_fasta(grp, k)
        assert np.array_equal(hist, w_ref[i])
        assert int(hist.sum()) == st_["distinct"] <= 4 ** k          # every distinct k-mer counted once
        assert not hist[5:].any()                                     # occupancy <= members of the group
    hist, st_ = engine.across_groups()
    assert np.array_equal(hist, a_ref) and not hist[4:].any()


def test_domain_invariants_full_size(engine):
    """Size-independent properties at BASELINE scale (one config-2 group, 50 x 5 Mbp): histogram mass = distinct count,
    idempotence (a group unioned with itself), reverse-complement invariance of a genome's k-mer set.  The oracle comparison
    at the same size is tests/test_gpu_fullsize.py."""
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=1, genomes_per_group=50, genome_len=5_000_000)
    genomes = [synth.make_genome(cfg, 1, i) for i in range(1, 51)]
    engine.group_sets_reset()
    h1, s1 = engine.group_from_fasta(genomes, 31, keep_set=False)
    assert int(h1.sum()) == s1["distinct"] and not h1[51:].any() and h1[50] > 0
    assert s1["genome_distinct"] <= s1["windows"] and s1["distinct"] <= s1["genome_distinct"]
    # every genome listed twice: same distinct set, every occupancy doubles
    h2, s2 = engine.group_from_fasta(genomes[:10] + genomes[:10], 31, keep_set=False)
    h10, s10 = engine.group_from_fasta(genomes[:10], 31, keep_set=False)
    assert s2["distinct"] == s10["distinct"]
    assert np.array_equal(h2[2::2][:10], h10[1:11]) and not h2[1::2].any()
    # reverse complement of a genome has the same canonical k-mer set
    seq = b"".join(l for l in genomes[0].split(b"\n") if not l.startswith(b">"))
    rc = seq[::-1].translate(bytes.maketrans(b"ACGTacgt", b"TGCAtgca"))
    ha, sa = engine.group_from_fasta([b">f\n" + seq + b"\n", b">r\n" + rc + b"\n"], 31, keep_set=False)
    assert ha[1] == 0 and int(ha[2]) == sa["distinct"]
