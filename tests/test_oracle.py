"""The oracle pinned against everything the reference offers for this path (CPU, no GPU needed):
golden canonical k-mers produced by the reference's get_canonical_kmer, the inline invariants of
SURVEY.md section 4, the hand-derived known-answer example of SURVEY.md section 3.6, and an independent
pure-Python restatement on random ragged inputs."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from helpers import EDGE_FASTAS, as_py, random_fasta
from oracle import pyoracle as P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_canonical_matches_reference_vectors(oracle):
    vec = json.load(open(os.path.join(GOLDEN, "canonical.json")))["vectors"]
    assert len(vec) > 200
    for kmer, canon in vec:
        k = len(kmer)
        assert P.canonical(kmer) == canon
        keys, nsym = oracle.kmers(b">t\n" + kmer.encode() + b"\n", k)
        assert nsym == k and keys.shape[0] == 1
        assert as_py(keys)[0] == P.encode(canon), (kmer, canon)
        # lower case is accepted and means the same k-mer
        keys2, _ = oracle.kmers(kmer.lower().encode(), k)
        assert np.array_equal(keys, keys2)


def test_worked_example_survey_3_6(oracle):
    groups = [[b">a\nACGTACGTNACGGT\n", b">b\nacgttgca\n"], [b">c\nGGGCCCAT\n", b">d\nTTACGNNAC\n"]]
    expect_sets = [["ACC", "ACG", "CCG", "GTA"], ["AAC", "ACG", "CAA", "GCA"], ["ATG", "CCA", "CCC", "GCC"], ["ACG", "GTA", "TAA"]]
    flat = [g for grp in groups for g in grp]
    for g, exp in zip(flat, expect_sets):
        assert P.genome_set(g, 3) == exp
        assert [P.decode(v, 3) for v in as_py(oracle.genome_set(g, 3))] == exp
    within, across, stats = oracle.exp1(flat, [0, 0, 1, 1], 2, 3)
    assert list(within[0][1:4]) == [6, 1, 0] and list(within[1][1:4]) == [7, 0, 0]
    assert list(across[1:4]) == [10, 2, 0]
    assert stats["distinct"] == 12 and stats["sum_group_distinct"] == 14


@pytest.mark.parametrize("k", [1, 2, 4, 7, 16, 31, 32, 33, 48, 64])
def test_c_oracle_equals_python_oracle_on_edge_inputs(oracle, k):
    rng = np.random.default_rng(k)
    files = list(EDGE_FASTAS) + [random_fasta(rng, 3000, p_n=0.01), random_fasta(rng, 2500, crlf=True, line=17)]
    for f in files:
        keys, nsym = oracle.kmers(f, k)
        ref = P.kmers(f, k)
        assert nsym == P.n_symbols(f)
        assert as_py(keys) == [P.encode(x) for x in ref]
        assert as_py(oracle.genome_set(f, k)) == [P.encode(x) for x in P.genome_set(f, k)]


@settings(max_examples=60, deadline=None)
@given(st.lists(st.lists(st.text(alphabet="ACGTacgtNnRY>\n\r ;", min_size=0, max_size=120), min_size=1, max_size=4), min_size=1, max_size=3),
       st.integers(min_value=1, max_value=40))
def test_exp1_property_random_text(groups, k):
    from oracle import oracle as O
    bgroups = [[t.encode() for t in grp] for grp in groups]
    flat = [g for grp in bgroups for g in grp]
    gid = [i for i, grp in enumerate(bgroups) for _ in grp]
    w, a, _ = O.exp1(flat, gid, len(bgroups), k)
    wr, ar, _ = P.exp1(bgroups, k)
    assert [list(map(int, r)) for r in w] == wr and list(map(int, a)) == ar


def test_reference_invariants(oracle):
    """max multiplicity <= members; a set's histogram is all count 1; counters saturate."""
    rng = np.random.default_rng(3)
    base = random_fasta(rng, 20000, p_n=0.001)
    genomes = [base, base, base[:len(base) // 2], random_fasta(rng, 15000)]
    sets = [oracle.genome_set(g, 21) for g in genomes]
    keys, counts = oracle.union_sum(sets, 21)
    assert counts.max() <= len(genomes) and counts.min() >= 1
    assert np.all(keys[1:] > keys[:-1])
    h = oracle.histogram(counts)
    assert h.sum() == keys.shape[0] and h[0] == 0
    k2, c2 = oracle.union_sum([keys], 21)
    assert (c2 == 1).all() and np.array_equal(k2, keys)
    k3, c3 = oracle.union_sum([sets[0]] * 9, 21, cs=5)
    assert (c3 == 5).all()
    # canonical: a sequence and its reverse complement have the same k-mer set
    seq = b"ACGTTGCATGCCGATAGGCTAGCTAGGATCGATCGGGATATTTAGCGC"
    rc = seq[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))
    for k in (5, 31, 33):
        assert np.array_equal(oracle.genome_set(seq, k), oracle.genome_set(rc, k))


def test_generator_counts_and_determinism(oracle):
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=2, genomes_per_group=2, genome_len=30_000, seed=11)
    g = synth.make_genome(cfg, 1, 1)
    assert g == synth.make_genome(cfg, 1, 1) and g != synth.make_genome(cfg, 1, 2)
    assert synth.count_bases(g) == P.n_symbols(g) == oracle.kmers(g, 31)[1]
    assert b">" in g and b"N" in g and any(c in g for c in b"acgt")
    assert max(len(l) for l in g.split(b"\n") if not l.startswith(b">")) <= 80


@pytest.mark.parametrize("k", [4, 11, 31, 33, 47])
def test_exp2_c_oracle_equals_python_oracle(oracle, k):
    """Experiment type 2 (exp_type_2.smk:297-508): the numpy/C statement against the set/dict statement, and the
    invariants the reference asserts (exp_type_2.smk:184-185): the intersect histogram has no row 1, the subtract
    histogram has nothing but row 1."""
    from khoice_b200 import synth
    from oracle import pyoracle
    cfg = synth.SynthConfig(n_groups=3, genomes_per_group=4, genome_len=2500, seed=17)
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, 4)] for g in (1, 2, 3)]
    pivots = [synth.make_genome(cfg, g, 4) for g in (1, 2, 3)]
    groups[2] = []  # a dataset with an empty rest of set
    w, a = oracle.exp2(groups, pivots, k, nbins=30)
    pw, pa = pyoracle.exp2(groups, pivots, k, nbins=30)
    for d in range(3):
        for ref, got in ((pw, w), (pa, a)):
            assert list(got[d, 0]) == ref[d][0] and list(got[d, 1]) == ref[d][1], (k, d)
            assert got[d, 1, 1] == 0 and got[d, 0, 2:].sum() == 0
        n_pivot = len(pyoracle.genome_set(pivots[d], k))
        assert int(w[d].sum()) == n_pivot and int(a[d].sum()) == n_pivot
    assert w[2, 1].sum() == 0  # nothing to intersect with
