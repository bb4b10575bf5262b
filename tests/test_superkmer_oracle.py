"""The oracle of the NEXT design step (oracle/superkmer.py: minimizer bins and super-k-mers) against the
existing oracle: the bin partition loses and duplicates nothing, is strand- and genome-independent, and counting bin by bin
gives the step_4 histogram of the whole group."""
import numpy as np
import pytest

from helpers import EDGE_FASTAS, random_fasta, symbol_stream


@pytest.mark.parametrize("k,m,lb", [(31, 11, 10), (21, 9, 6), (15, 15, 4), (32, 7, 12), (5, 3, 3), (47, 13, 9), (64, 32, 5)])
def test_superkmers_tile_the_valid_windows_exactly(oracle, k, m, lb):
    from oracle import superkmer as S
    rng = np.random.default_rng(k)
    texts = list(EDGE_FASTAS) + [random_fasta(rng, 20_000, p_n=0.004), random_fasta(rng, 3_000, crlf=True)]
    for t in texts:
        code, valid = S.symbol_stream(t)
        rc, rv = symbol_stream(t)                                   # the pure-Python statement of K1 (tests/helpers.py)
        assert code.tolist() == rc and valid.astype(int).tolist() == rv
        ok, kmer, bins = S.window_bins(t, k, m, lb)
        keys, nsym = oracle.kmers(t, k)                             # valid windows in stream order (R1-R5)
        assert np.array_equal(kmer[ok], keys)
        sk = S.superkmers(t, k, m, lb)
        covered = np.zeros(ok.size, dtype=bool)
        for b, s, n in sk:
            assert n >= 1 and not covered[s:s + n].any() and ok[s:s + n].all() and (bins[s:s + n] == b).all()
            covered[s:s + n] = True
        assert np.array_equal(covered, ok)
        for (b0, s0, n0), (b1, s1, n1) in zip(sk, sk[1:]):          # maximal: neighbours differ in bin or are not adjacent
            assert s1 > s0 and (s0 + n0 != s1 or b0 != b1)


def test_bin_depends_on_the_kmer_only(oracle):
    """Same k-mer -> same bin, in any genome and on either strand (the reverse complement of a text has the same bins)."""
    from oracle import superkmer as S
    rng = np.random.default_rng(3)
    seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=5000).tobytes()
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    fwd, rev = b">a\n" + seq + b"\n", b">b\n" + seq.translate(comp)[::-1] + b"\n"
    k, m, lb = 31, 11, 12
    ok1, k1, b1 = S.window_bins(fwd, k, m, lb)
    ok2, k2, b2 = S.window_bins(rev, k, m, lb)
    d1 = dict(zip(k1[ok1].tolist(), b1[ok1].tolist()))
    d2 = dict(zip(k2[ok2].tolist(), b2[ok2].tolist()))
    assert d1 == d2 and len(set(d1.values())) > 100


@pytest.mark.parametrize("k,m,lb", [(31, 11, 8), (13, 7, 5), (40, 11, 7)])
def test_counting_bin_by_bin_gives_the_group_histogram(oracle, k, m, lb):
    from khoice_b200 import synth
    from oracle import superkmer as S
    cfg = synth.SynthConfig(n_groups=1, genomes_per_group=6, genome_len=20_000, seed=17)
    genomes = [synth.make_genome(cfg, 1, i) for i in range(1, 7)] + [EDGE_FASTAS[1] + EDGE_FASTAS[3]]
    hist, sets = S.binned_group_histogram(genomes, k, m, lb, nbins=64)
    w_ref, _, st = oracle.exp1(genomes, [0] * len(genomes), 1, k, nbins=64)
    assert np.array_equal(hist, w_ref[0])
    allk = np.concatenate(list(sets.values()))
    n_unique = np.unique(allk).size if k <= 32 else np.unique(allk, axis=0).shape[0]
    assert allk.shape[0] == n_unique == st["sum_group_distinct"]      # bins are disjoint


@pytest.mark.parametrize("k,m,lb", [(31, 11, 8), (35, 11, 6)])
def test_across_group_stage_bin_by_bin(oracle, k, m, lb):
    """A k-mer lands in the same bin in every group, so the step_8 histogram is the sum of per-bin counts over the groups' sets."""
    from khoice_b200 import synth
    from oracle import superkmer as S
    cfg = synth.SynthConfig(n_groups=4, genomes_per_group=3, genome_len=15_000, seed=29)
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, 4)] for g in range(1, 5)]
    per_group = []
    for gi, genomes in enumerate(groups):
        hist, sets = S.binned_group_histogram(genomes, k, m, lb, nbins=64)
        per_group.append(sets)
    flat = [t for g in groups for t in g]
    gid = [gi for gi, g in enumerate(groups) for _ in g]
    w_ref, a_ref, st = oracle.exp1(flat, gid, 4, k, nbins=64)
    across, distinct = S.binned_across_histogram(per_group, k, nbins=64)
    assert np.array_equal(across, a_ref) and distinct == st["distinct"]
