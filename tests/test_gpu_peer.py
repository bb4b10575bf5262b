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


@pytest.mark.parametrize("async_push", ["0", "1"])
def test_two_ranks_on_one_gpu_push_over_ipc(async_push):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), KHB_PEER_ASYNC=async_push)
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
