"""The CUDA path against the oracle at BASELINE.json's FULL group sizes (the oracle needs seconds per group on the GPU box's
host cores): a config-2 set of two 50 x 5 Mbp groups at k = 31, one config-4 group (100 genomes, 128-bit words) at k = 47 and
k = 63, one config-5 group (200 genomes, 10^9 windows in one group stage) at k = 31 -- in every group mode that can run them.
Reference path: /root/reference/workflow/rules/exp_type_1.smk:156-259."""
import numpy as np
import pytest

from helpers import synth_genomes

pytestmark = pytest.mark.gpu

MODES = ["auto", "single-sort"]


@pytest.fixture(scope="module")
def threads(oracle):
    return oracle.set_num_threads(oracle.host_cores())


@pytest.fixture()
def mode(engine, request):
    engine.set_group_mode(request.param)
    yield request.param
    engine.set_group_mode("auto")


def _check_groups(engine, oracle, groups, k):
    flat = [f for grp in groups for f in grp]
    gid = [i for i, grp in enumerate(groups) for _ in grp]
    w_ref, a_ref, st_ref = oracle.exp1(flat, gid, len(groups), k)
    engine.group_sets_reset()
    distinct = 0
    for i, grp in enumerate(groups):
        hist, st = engine.group_from_fasta(grp, k)
        assert np.array_equal(hist, w_ref[i]), (k, i, np.flatnonzero(hist != w_ref[i])[:8])
        assert int(hist.sum()) == st["distinct"] and not hist[len(grp) + 1:].any()
        distinct += st["distinct"]
    hist, st = engine.across_groups()
    assert np.array_equal(hist, a_ref), (k, np.flatnonzero(hist != a_ref)[:8])
    assert distinct == st_ref["sum_group_distinct"] and st["distinct"] == st_ref["distinct"]
    engine.group_sets_reset()


@pytest.mark.parametrize("mode", MODES, indirect=True)
def test_config2_two_full_groups_k31(engine, oracle, threads, mode):
    groups = [synth_genomes(50, group=g) for g in (1, 2)]
    _check_groups(engine, oracle, groups, 31)


@pytest.mark.parametrize("mode", MODES, indirect=True)
@pytest.mark.parametrize("k", [47, 63])
def test_config4_full_group_128bit(engine, oracle, threads, mode, k):
    _check_groups(engine, oracle, [synth_genomes(100, group=3)], k)


@pytest.mark.parametrize("mode", MODES, indirect=True)
def test_config5_full_group_200_genomes_k31(engine, oracle, threads, mode):
    """10^9 windows in one group stage (32-bit tile / offset limits of the sort, table and bin sizing of the bin path)."""
    _check_groups(engine, oracle, [synth_genomes(200, group=7)], 31)
