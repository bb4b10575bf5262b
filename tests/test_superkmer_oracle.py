"""The numpy statement of the minimizer-bin partition (oracle/superkmer.py; the product kernels are khoice_b200/csrc/bins.cu) against the
exp-1 oracle: the partition loses and duplicates nothing, is strand- and genome-independent, and counting bin by bin gives the
step_4 histogram of the whole group and the step_8 histogram across groups."""
import numpy as np
import pytest

from helpers import EDGE_FASTAS, random_fasta, symbol_stream


@pytest.mark.parametrize("k,m,nb", [(31, 13, 1000), (21, 9, 61), (15, 15, 16), (32, 7, 4096), (5, 3, 7), (47, 13, 513), (64, 13, 32)])
def test_superkmers_tile_the_valid_windows_exactly(oracle, k, m, nb):
    from oracle import superkmer as S
    rng = np.random.default_rng(k)
    texts = list(EDGE_FASTAS) + [random_fasta(rng, 20_000, p_n=0.004), random_fasta(rng, 3_000, crlf=True)]
    for t in texts:
        code, valid = S.symbol_stream(t)
        rc, rv = symbol_stream(t)                                   # the pure-Python statement of K1 (tests/helpers.py)
        assert code.tolist() == rc and valid.astype(int).tolist() == rv
        ok, kmer, mh, bins = S.window_bins(t, k, nb, m)
        keys, nsym = oracle.kmers(t, k)                             # valid windows in stream order (R1-R5)
        assert np.array_equal(kmer[ok], keys)
        sk = S.superkmers(t, k, nb, m)
        covered = np.zeros(ok.size, dtype=bool)
        for b, s, n in sk:
            assert n >= 1 and not covered[s:s + n].any() and ok[s:s + n].all() and (bins[s:s + n] == b).all()
            covered[s:s + n] = True
        assert np.array_equal(covered, ok)
        for (b0, s0, n0), (b1, s1, n1) in zip(sk, sk[1:]):          # maximal: neighbours differ in minimizer or are not adjacent
            assert s1 > s0 and (s0 + n0 != s1 or mh[s0] != mh[s1])
        pieces = S.records(t, k, nb, first_symbol=1234, m=m)          # the kernel's records: the same windows, cut at tiles and at 32
        cov2 = np.zeros(ok.size, dtype=bool)
        for b, s0, n0 in pieces:
            assert 1 <= n0 <= S.CAPW and not cov2[s0:s0 + n0].any() and (s0 + 1234) // S.TILE == (s0 + n0 - 1 + 1234) // S.TILE
            cov2[s0:s0 + n0] = True
        assert np.array_equal(cov2, ok)


def test_bin_depends_on_the_kmer_only(oracle):
    """Same k-mer -> same bin, in any genome and on either strand (the reverse complement of a text has the same bins)."""
    from oracle import superkmer as S
    rng = np.random.default_rng(3)
    seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=5000).tobytes()
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    fwd, rev = b">a\n" + seq + b"\n", b">b\n" + seq.translate(comp)[::-1] + b"\n"
    k, nb = 31, 4096
    ok1, k1, _, b1 = S.window_bins(fwd, k, nb)
    ok2, k2, _, b2 = S.window_bins(rev, k, nb)
    d1 = dict(zip(k1[ok1].tolist(), b1[ok1].tolist()))
    d2 = dict(zip(k2[ok2].tolist(), b2[ok2].tolist()))
    assert d1 == d2 and len(set(d1.values())) > 100


@pytest.mark.parametrize("k,nb", [(31, 256), (17, 31), (40, 100)])
def test_counting_bin_by_bin_gives_the_group_histogram(oracle, k, nb):
    from khoice_b200 import synth
    from oracle import superkmer as S
    cfg = synth.SynthConfig(n_groups=1, genomes_per_group=6, genome_len=20_000, seed=17)
    genomes = [synth.make_genome(cfg, 1, i) for i in range(1, 7)] + [EDGE_FASTAS[1] + EDGE_FASTAS[3]]
    hist, sets = S.binned_group_histogram(genomes, k, nb, nbins=64)
    w_ref, _, st = oracle.exp1(genomes, [0] * len(genomes), 1, k, nbins=64)
    assert np.array_equal(hist, w_ref[0])
    allk = np.concatenate(list(sets.values()))
    n_unique = np.unique(allk).size if k <= 32 else np.unique(allk, axis=0).shape[0]
    assert allk.shape[0] == n_unique == st["sum_group_distinct"]      # bins are disjoint


@pytest.mark.parametrize("k,nb", [(31, 256), (35, 64)])
def test_across_group_stage_bin_by_bin(oracle, k, nb):
    """A k-mer lands in the same bin in every group, so the step_8 histogram is the sum of per-bin counts over the groups' sets."""
    from khoice_b200 import synth
    from oracle import superkmer as S
    cfg = synth.SynthConfig(n_groups=4, genomes_per_group=3, genome_len=15_000, seed=29)
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, 4)] for g in range(1, 5)]
    per_group = []
    for gi, genomes in enumerate(groups):
        hist, sets = S.binned_group_histogram(genomes, k, nb, nbins=64)
        per_group.append(sets)
    flat = [t for g in groups for t in g]
    gid = [gi for gi, g in enumerate(groups) for _ in g]
    w_ref, a_ref, st = oracle.exp1(flat, gid, 4, k, nbins=64)
    across, distinct = S.binned_across_histogram(per_group, k, nbins=64)
    assert np.array_equal(across, a_ref) and distinct == st["distinct"]
