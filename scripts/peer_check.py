"""torchrun check of the across-group exchange on N real GPUs (NCCL control plane): the peer-memory push and the NCCL
all-to-all must give the same histograms; a deliberately tiny region must trigger the fallback + regrow.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/peer_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from khoice_b200 import synth, dist as kd
from khoice_b200.engine import Engine

rank, world, local = kd.init_from_env("nccl")
G, N, L = 4 * world, 10, 1_000_000
cfg = synth.SynthConfig(n_groups=G, genomes_per_group=N, genome_len=L, seed=99)
mine = kd.groups_of_rank(G, rank, world)
groups = {g: [synth.make_genome(cfg, g, i) for i in range(1, N + 1)] for g in mine}
eng = Engine(local)
ad = kd.CudaAdapter(eng, torch.device("cuda", local))
for k in (31, 47):
    res = {}
    for mode, region in (("nccl", None), ("peer", None), ("peer", 1024)):
        ex = kd.AcrossExchanger(ad, k, G, mode=mode, region_keys=region)
        hs = []
        for rnd in range(3):
            eng.group_sets_reset(); ex.begin()
            torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
            for g in mine:
                eng.group_from_fasta(groups[g], k); ex.after_group()
            h, info = ex.finish()
            torch.cuda.synchronize(); dt = time.time() - t0
            hs.append(h)
            if rank == 0:
                print(f"k={k} mode={mode} region={region} round {rnd}: {info['exchange']} {dt * 1e3:.1f} ms, region_keys now {eng.peer_region_keys}", flush=True)
        res[(mode, region)] = hs
        if mode == "peer" and region is None:
            assert ex.rounds_nccl == 1 and ex.rounds_peer == 2, (ex.rounds_nccl, ex.rounds_peer)
        if region == 1024:
            assert ex.rounds_nccl == 1 and ex.rounds_peer == 2, (ex.rounds_nccl, ex.rounds_peer)   # round 0 overflows, falls back, regrows
        ex.close()
    ref = res[("nccl", None)][0]
    for key, hs in res.items():
        for h in hs:
            assert np.array_equal(h, ref), (k, key)
if rank == 0:
    print("peer_check ok")
dist.barrier()
eng.close()
dist.destroy_process_group()
