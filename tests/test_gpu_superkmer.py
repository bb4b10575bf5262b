"""EXPERIMENTAL kernel of the next design step (csrc/superkmer.cu): the count pass of the minimizer
partition against its oracle (oracle/superkmer.py) -- per bin the number of windows and of super-k-mers, bit for bit."""
import numpy as np
import pytest

from helpers import EDGE_FASTAS, random_fasta

pytestmark = pytest.mark.gpu

TILE = 4352   # KHB_SUPERKMER_TILE


@pytest.mark.parametrize("k,m,lb", [(31, 11, 12), (21, 9, 8), (15, 15, 6), (32, 7, 10), (5, 3, 3)])
def test_bin_counts_match_the_oracle(engine, k, m, lb):
    from khoice_b200 import synth
    from oracle import superkmer as S
    rng = np.random.default_rng(k + m)
    cfg = synth.SynthConfig(n_groups=1, genomes_per_group=2, genome_len=60_000, seed=5)
    texts = [b"".join(EDGE_FASTAS), random_fasta(rng, 30_000, p_n=0.003), synth.make_genome(cfg, 1, 1) + synth.make_genome(cfg, 1, 2)]
    for t in texts:
        staged = engine.stage_fasta([t])
        packed = engine.pack_fasta(staged)
        try:
            win, sk, _ = engine.superkmer_count(packed, k, m, lb)
        finally:
            for b in (staged.buf, packed["codes"], packed["valid"]):
                b.free()
        # the staged text carries filler behind the file (one more break symbol): it adds no window
        ref_w = np.zeros(1 << lb, dtype=np.int64)
        ref_s = np.zeros(1 << lb, dtype=np.int64)
        cap = 65 - k                                  # a record holds k - 1 + len <= 64 symbols: longer runs are several records
        for b, s, n in S.superkmers(t, k, m, lb, tile=TILE):
            ref_w[b] += n
            ref_s[b] += -(-n // cap)
        assert np.array_equal(win.astype(np.int64), ref_w), (k, m, lb)
        assert np.array_equal(sk.astype(np.int64), ref_s), (k, m, lb)
        ok, _, _ = S.window_bins(t, k, m, lb)
        assert int(win.sum()) == int(ok.sum())


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("k,m,lb", [(31, 11, 10), (21, 9, 8), (13, 7, 6), (32, 12, 9)])
def test_group_stage_through_minimizer_bins_equals_the_product_path(engine, oracle, k, m, lb, compact):
    """EXPERIMENT: count pass + scatter + one CTA per bin with a shared-memory table == the single-sort path == the oracle."""
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=1, genomes_per_group=7, genome_len=40_000, seed=23)
    genomes = [synth.make_genome(cfg, 1, i) for i in range(1, 8)] + [EDGE_FASTAS[1] + EDGE_FASTAS[3], b""]
    hist, st = engine.superkmer_group(genomes, k, m, lb, nbins=64, compact=compact)   # compact: 24-byte super-k-mer records in the bins
    assert st["overflowed_bins"] == 0
    engine.group_sets_reset()
    ref, rst = engine.group_from_fasta(genomes, k, nbins=64, keep_set=False)
    w_ref, _, ost = oracle.exp1(genomes, [0] * len(genomes), 1, k, nbins=64)
    assert np.array_equal(ref, w_ref[0])
    assert np.array_equal(hist, w_ref[0]), (hist[:10], w_ref[0][:10])
    assert st["distinct"] == rst["distinct"] == ost["sum_group_distinct"]
    assert st["genome_distinct"] == ost["sum_genome_distinct"]
