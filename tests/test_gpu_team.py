"""ONE group sharded over the members of a team (csrc/team.cu, csrc/bins.cu: mb_partition_kernel<KW, true>, dist.TeamSharder) on ONE
GPU: two processes share cuda:0, each packs and partitions its slice of the group's genomes, stores the super-k-mer records into the
record buffer of the bin's owner through CUDA IPC, and counts the bins it owns; the control plane runs over gloo.  The members' partial
histograms must add up to the oracle's step_4 histogram, and the across-group stage over the keys the members emitted must give the
oracle's step_8 histogram -- for 64- and 128-bit k-mers, slices of more than 64 genomes, every kind of source, forced region overflow
(retry with larger regions) and forced table overflow (hash classes, the big-bin pass)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch
import torch.distributed as dist
from khoice_b200 import synth, dist as kd
from khoice_b200.engine import Engine
from oracle import oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
eng = Engine(0)
ad = kd.CudaAdapter(eng, torch.device("cuda", 0))
ts = kd.TeamSharder(eng, world, rank)

def summed(h):
    t = torch.from_numpy(np.ascontiguousarray(h).astype(np.int64)); dist.all_reduce(t); return t.numpy().astype(np.uint64)

def run(G, N, L, ks, seed, rounds=2, check_across=True, force=None):
    cfg = synth.SynthConfig(n_groups=G, genomes_per_group=N, genome_len=L, seed=seed)
    groups = {g: [synth.make_genome(cfg, g, i) for i in range(1, N + 1)] for g in range(1, G + 1)}
    flat = [f for g in range(1, G + 1) for f in groups[g]]
    gid = [g - 1 for g in range(1, G + 1) for _ in range(N)]
    slices = kd.genome_slices(N, world)
    sizes = [hi - lo for lo, hi in slices]
    lo, hi = slices[rank]
    NB = max(64, N + 2)                         # histogram rows: every possible count has one
    for k in ks:
        w_ref, a_ref, st_ref = O.exp1(flat, gid, G, k, nbins=NB)
        ex = kd.AcrossExchanger(ad, k, G, nbins=NB, mode="peer", region_keys=max(400_000, 2 * N * L)) if check_across else None
        for rnd in range(rounds):
            eng.group_sets_reset()
            if ex: ex.begin()
            for g in range(1, G + 1):
                mine = groups[g][lo:hi]
                n_sym = sum(len(f) for f in groups[g])
                which = (rnd + g) %% 3
                src = mine if which == 0 else eng.stage_fasta(mine) if which == 1 else eng.pack_group(mine)
                if force:
                    force(ts, k, N, sizes)
                h, st = ts.run_group(src, k, N, sizes, n_sym, nbins=NB, keep_set=check_across)
                if ex: ex.after_group()
                assert np.array_equal(summed(h), w_ref[g - 1]), (k, g, rnd, summed(h)[:8], w_ref[g - 1][:8])
                assert int(summed(np.array([st["distinct"]]))[0]) == int(w_ref[g - 1][1:].sum()), (k, g)
                if which == 2: src.free()
            if ex:
                hist, info = ex.finish()
                assert info["exchange"] == "peer", info
                assert np.array_equal(hist, a_ref), (k, rnd, rank)
        if ex: ex.close()

# 64- and 128-bit k-mers, uneven slices (3 + 2 genomes), all three kinds of source, across-group stage over the emitted keys
run(3, 5, 40_000, (31, 21, 47, 63), seed=11)
assert ts.retries == 0
# slices of more than 64 genomes: chunks 0-1 belong to member 0, 2-3 to member 1
run(2, 130, 3_000, (31, 47), seed=12, rounds=1)
# larger genomes: many tiles per slice, bins with several lumps
run(2, 6, 600_000, (31,), seed=13, rounds=1)
# a region that is too small: every member partitions again into larger regions
r0 = ts.retries
def tiny_regions(ts, k, N, sizes):
    ts.hints[(k, N, kd.chunk_layout(sizes)[1])] = {"cap": 2}
run(2, 5, 40_000, (31, 47), seed=14, rounds=1, force=tiny_regions)
assert ts.retries > r0, (ts.retries, r0)
# areas of the receive buffers that are too small for a sender's share: every member partitions again, the buffers grow
r0, s0 = ts.retries, ts.setups
def small_areas(ts, k, N, sizes):
    key = (k, N, kd.chunk_layout(sizes)[1])
    if "area_pct" not in ts.hints.get(key, {}):
        ts.hints[key] = dict(ts.hints.get(key, {}), area_pct=30)
run(2, 6, 600_000, (31,), seed=17, rounds=1, force=small_areas)
assert ts.retries > r0, (ts.retries, r0)
# tables that are too small for their bins: hash classes and the big-bin pass
os.environ["KHB_BINS_SLOTS_LOG2"] = "8"
os.environ["KHB_BINS_RHO_PCT"] = "2"
b0 = eng.bins_counters["big_bins"]
run(2, 5, 40_000, (31, 47), seed=15, rounds=1)
assert eng.bins_counters["big_bins"] > b0
del os.environ["KHB_BINS_SLOTS_LOG2"], os.environ["KHB_BINS_RHO_PCT"]
# without a retained set (histograms only)
run(1, 4, 20_000, (25,), seed=16, rounds=1, check_across=False)
ts.close()
eng.close()
dist.barrier()
dist.destroy_process_group()
print("team ok", rank)
"""


def _spawn(world, extra_env=None):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), **(extra_env or {}))
        procs.append(subprocess.Popen([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    # a member that fails leaves the others waiting in a collective: stop everybody as soon as one exits with an error
    import time
    deadline = time.time() + 600
    while any(p.poll() is None for p in procs):
        if any(p.poll() not in (None, 0) for p in procs) or time.time() > deadline:
            time.sleep(2)
            for q in procs:
                if q.poll() is None:
                    q.kill()
            break
        time.sleep(0.2)
    outs = [p.communicate() for p in procs]
    for rank, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"team ok {rank}" in out, f"member {rank} rc={p.returncode}\n" + out[-2000:] + err[-4000:]


@pytest.mark.parametrize("world,fuse", [(2, "1"), (2, "0"), (3, "1")])
def test_members_of_a_team_on_one_gpu(world, fuse):
    _spawn(world, {"KHB_PEER_FUSE": fuse})


PIPELINE_SCRIPT = r"""
import os, sys
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from khoice_b200 import pipeline
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rep = pipeline.run_fused_distributed(%(work)r, %(groups)d, %(ks)r, exchange="peer", team=%(team)d)
assert rep["team_size"] == 2, rep["team_size"]
print("pipeline ok", os.environ["RANK"])
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world,team", [(2, 2), (2, 0)])
def test_work_root_driver_with_a_team_on_one_gpu(engine, tmp_path, world, team):
    """pipeline.run_fused_distributed(team=2) with the product adapter (two processes on cuda:0, gloo control plane): every group is
    sharded over the two ranks, and the step_4 / step_8 files and both CSVs equal the single-process run's.  team=0 picks the team
    size itself: 3 groups on 2 ranks do not deal evenly, so it shards them too."""
    import filecmp
    from khoice_b200 import pipeline, synth
    cfg = synth.SynthConfig(n_groups=3, genomes_per_group=5, genome_len=30_000, seed=78)
    ks = ["31", "40", "21"]           # the key width changes twice
    work, ref = str(tmp_path / "dist"), str(tmp_path / "single")
    synth.write_dataset(cfg, work)
    synth.write_dataset(cfg, ref)
    pipeline.run_fused(ref, cfg.n_groups, ks, engine=engine)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = PIPELINE_SCRIPT % {"root": ROOT, "work": work, "groups": cfg.n_groups, "ks": ks, "team": team}
    procs = [subprocess.Popen([sys.executable, "-c", script], env=dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1",
                                                                     MASTER_PORT=str(port)), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    import time
    deadline = time.time() + 600
    while any(p.poll() is None for p in procs):
        if any(p.poll() not in (None, 0) for p in procs) or time.time() > deadline:
            time.sleep(2)
            for q in procs:
                if q.poll() is None:
                    q.kill()
            break
        time.sleep(0.2)
    outs = [p.communicate() for p in procs]
    for r, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"pipeline ok {r}" in out, out[-2000:] + err[-4000:]
    for k in ks:
        for num in range(1, cfg.n_groups + 1):
            assert filecmp.cmp(os.path.join(work, pipeline.p_step4(k, num)), os.path.join(ref, pipeline.p_step4(k, num)), shallow=False), (k, num)
        assert filecmp.cmp(os.path.join(work, pipeline.p_step8(k)), os.path.join(ref, pipeline.p_step8(k)), shallow=False), k
    for f in (pipeline.P_STEP5, pipeline.P_STEP9) + pipeline.P_FINAL:
        assert filecmp.cmp(os.path.join(work, f), os.path.join(ref, f), shallow=False), f
