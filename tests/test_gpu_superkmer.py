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
        for b, s, n in S.superkmers(t, k, m, lb, tile=TILE):
            ref_w[b] += n
            ref_s[b] += 1
        assert np.array_equal(win.astype(np.int64), ref_w), (k, m, lb)
        assert np.array_equal(sk.astype(np.int64), ref_s), (k, m, lb)
        ok, _, _ = S.window_bins(t, k, m, lb)
        assert int(win.sum()) == int(ok.sum())
