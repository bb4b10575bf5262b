"""The multi-GPU exchange over peer memory (csrc/peer.cu + dist.AcrossExchanger) on ONE GPU: two processes share cuda:0,
map each other's receive buffer through CUDA IPC and push their groups' k-mers to the hash-range owner; the control plane
(IPC handles, count table, histogram all-reduce) runs over gloo.  The histograms must equal the oracle's, i.e. must not
depend on the number of ranks, for 8- and 16-byte keys; a region that is too small must raise the flag."""
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
G, N = 5, 3
cfg = synth.SynthConfig(n_groups=G, genomes_per_group=N, genome_len=30_000, seed=321)
groups = {g: [synth.make_genome(cfg, g, i) for i in range(1, N + 1)] for g in range(1, G + 1)}
flat = [f for g in range(1, G + 1) for f in groups[g]]
gid = [g - 1 for g in range(1, G + 1) for _ in range(N)]
eng = Engine(0)
ad = kd.CudaAdapter(eng, torch.device("cuda", 0))
mine = kd.groups_of_rank(G, rank, world)
for k in (31, 21, 47):
    w_ref, a_ref, st_ref = O.exp1(flat, gid, G, k, nbins=64)
    ex = kd.AcrossExchanger(ad, k, G, nbins=64, mode="peer", region_keys=200_000)
    assert ex.ready and eng.peer_region_keys == 200_000
    eng.profile_enable(True)
    for rnd in range(3):                       # the regions are reused round after round
        eng.group_sets_reset()
        ex.begin()
        for g in mine:
            h, st = eng.group_from_fasta(groups[g], k, nbins=64)
            assert np.array_equal(h, w_ref[g - 1]), (k, g)
            ex.after_group()
        hist, info = ex.finish()
        assert info["exchange"] == "peer"
        assert np.array_equal(hist, a_ref), (k, rnd, rank)
        tot = torch.tensor([info["local_distinct"]]); dist.all_reduce(tot)
        assert int(tot.item()) == st_ref["distinct"]           # every k-mer has exactly one owner
    assert ex.rounds_peer == 3 and ex.rounds_nccl == 0
    # the minimizer-bin stage stores its keys into the owners' regions from its own end-of-bin pass: no push kernel runs
    # (KHB_PEER_FUSE=0 and the push on its own stream keep the separate pass)
    pushes = eng.profile_read()["partition"]["launches"]
    fused = os.environ.get("KHB_PEER_FUSE", "1") != "0" and os.environ.get("KHB_PEER_ASYNC", "0") == "0"
    assert (pushes == 0) == fused, (k, pushes, fused)
    eng.profile_enable(False)
    ex.close()
# a region that cannot hold a rank's share raises the overflow flag (the driver then redoes the round over NCCL)
ex = kd.AcrossExchanger(ad, 31, G, nbins=64, mode="peer", region_keys=64)
eng.group_sets_reset(); ex.begin()
for g in mine:
    eng.group_from_fasta(groups[g], 31, nbins=64); ex.after_group()
counts, ovf = eng.peer_counts(world)
assert ovf and counts.max() > 64, (counts, ovf)
dist.barrier()
ex.close()
eng.close()
dist.destroy_process_group()
print("peer ok", rank)
"""


@pytest.mark.parametrize("async_push,fuse", [("0", "1"), ("0", "0"), ("1", "1")])
def test_two_ranks_on_one_gpu_push_over_ipc(async_push, fuse):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), KHB_PEER_ASYNC=async_push, KHB_PEER_FUSE=fuse)
        procs.append(subprocess.Popen([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=600))
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for rank, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"peer ok {rank}" in out, out[-2000:] + err[-4000:]


PIPELINE_SCRIPT = r"""
import os, sys
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from khoice_b200 import pipeline
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rep = pipeline.run_fused_distributed(%(work)r, %(groups)d, %(ks)r, exchange=%(exchange)r)
routes = [s["exchange"] for s in rep["stages"] if s.get("group") == "across"]
print("pipeline ok", os.environ["RANK"], ",".join(routes))
dist.destroy_process_group()
"""


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_distributed_work_root_driver_on_two_ranks_sharing_one_gpu(engine, tmp_path, exchange):
    """pipeline.run_fused_distributed with the product adapter (two processes on cuda:0, gloo control plane): the step_4 /
    step_8 files and both CSVs equal the single-process run's."""
    import filecmp
    from khoice_b200 import pipeline, synth
    cfg = synth.SynthConfig(n_groups=5, genomes_per_group=3, genome_len=30_000, seed=77)
    ks = ["31", "40", "21", "13"]     # the key width changes twice, in both directions: one exchanger at a time (a ctx holds one peer exchange)
    work, ref = str(tmp_path / "dist"), str(tmp_path / "single")
    synth.write_dataset(cfg, work)
    synth.write_dataset(cfg, ref)
    pipeline.run_fused(ref, cfg.n_groups, ks, engine=engine)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = PIPELINE_SCRIPT % {"root": ROOT, "work": work, "groups": cfg.n_groups, "ks": ks, "exchange": exchange}
    procs = [subprocess.Popen([sys.executable, "-c", script], env=dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1",
                                                                     MASTER_PORT=str(port)), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    for r, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"pipeline ok {r}" in out, out[-2000:] + err[-4000:]
    if exchange == "peer":
        assert "peer" in outs[0][0]          # after the sizing round the push is the route
    for k in ks:
        for num in range(1, cfg.n_groups + 1):
            assert filecmp.cmp(os.path.join(work, pipeline.p_step4(k, num)), os.path.join(ref, pipeline.p_step4(k, num)), shallow=False), (k, num)
        assert filecmp.cmp(os.path.join(work, pipeline.p_step8(k)), os.path.join(ref, pipeline.p_step8(k)), shallow=False), k
    for f in (pipeline.P_STEP5, pipeline.P_STEP9) + pipeline.P_FINAL:
        assert filecmp.cmp(os.path.join(work, f), os.path.join(ref, f), shallow=False), f
