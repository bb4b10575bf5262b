"""The mixed-prefix-run machinery of the single-sort path (pairs_kernel bitmap -> mixed_collect_kernel ->
mixed_runs_kernel, including its quadratic overflow fallback) under prefixes that are far too short: KHB_PREFIX_SLACK
removes prefix bits, so that most prefix runs hold several (or hundreds of) distinct keys.  The variable is read once
per process, hence the sub-process.  KHB_GROUP_MODE=single-sort keeps the default hash path (hashset.cu) out of the way."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import os
import sys
import numpy as np
sys.path.insert(0, %(root)r)
from khoice_b200 import synth
from khoice_b200.engine import Engine
from oracle import oracle as O
cfg = synth.SynthConfig(n_groups=2, genomes_per_group=4, genome_len=20_000, seed=77)
groups = [[synth.make_genome(cfg, g, i) for i in range(1, 4)] for g in (1, 2)]
pivots = [synth.make_genome(cfg, g, 4) for g in (1, 2)]
# a k-mer that repeats many times inside one genome, so that long runs meet mixed prefixes
groups[0][0] += b">rep\n" + b"ACGTTGCATTGACCAGTAGGATCCATGCAAGT" * 40 + b"\n"
eng = Engine(0)
for k in (7, 15, 31, 47):
    flat = [f for g in groups for f in g]
    w_ref, a_ref, st_ref = O.exp1(flat, [0, 0, 0, 1, 1, 1], 2, k, nbins=64)
    eng.group_sets_reset()
    total = 0
    for d in range(2):
        hist, st = eng.group_from_fasta(groups[d], k, nbins=64)
        assert np.array_equal(hist, w_ref[d]), ("within", k, d)
        if int(os.environ["KHB_PREFIX_SLACK"]) < 0:
            assert st["passes_group"] <= 2, st   # 0 on the direct-address path (small k)
        total += st["genome_distinct"]
    assert total == st_ref["sum_genome_distinct"], (total, st_ref)
    hist, st = eng.across_groups(nbins=64)
    assert np.array_equal(hist, a_ref), ("across", k)
    # the retained sets are exactly the oracle's group sets
    got = eng.group_sets_download()
    w2, a2 = O.exp2(groups, pivots, k, nbins=64)
    eng.group_sets_reset()
    packs = []
    for d in range(2):
        pk = eng.pack_group(groups[d] + [pivots[d]])
        packs.append(pk)
        hist, _ = eng.pivot_group_from_packed(pk, k, nbins=64)
        assert np.array_equal(hist, w2[d][0] + w2[d][1]), ("pivot within", k, d)
    hists, _ = eng.pivot_across(nbins=64)
    for d in range(2):
        assert np.array_equal(hists[d], a2[d][0] + a2[d][1]), ("pivot across", k, d)
    for pk in packs:
        pk.free()
eng.close()
print("mixed ok")
"""


@pytest.mark.parametrize("slack,smallk_mb", [("-6", "1024"), ("-10", "1024"), ("-13", "0"), ("0", "0")])
def test_short_prefixes_give_the_same_answers(slack, smallk_mb, oracle):
    """smallk_mb = 0 switches the direct-address path for small k off, so k = 7 goes through the sort path too (runs of
    thousands of equal keys)."""
    env = dict(os.environ, KHB_PREFIX_SLACK=slack, KHB_SMALLK_TABLE_MB=smallk_mb, KHB_GROUP_MODE="single-sort")
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mixed ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
