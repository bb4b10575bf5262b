"""One-off validation at config-5 group size: 1 group x 200 synthetic 5 Mbp genomes (1e9 windows in one sort segment),
k=31, GPU histogram vs the CPU oracle.  usage: python scripts/validate_big_group.py [genomes] [k]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from khoice_b200 import synth
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 200
k = int(sys.argv[2]) if len(sys.argv) > 2 else 31
cfg = synth.SynthConfig(n_groups=1, genomes_per_group=ng, genome_len=5_000_000)
t0 = time.time()
from multiprocessing import get_context
def gen(i):
    return synth.make_genome(cfg, 1, i)
with get_context("fork").Pool(16) as pool:
    genomes = pool.map(gen, range(1, ng + 1))
print(f"generated {ng} genomes in {time.time() - t0:.1f}s, {sum(map(len, genomes)) / 1e9:.2f} GB", flush=True)
from khoice_b200.engine import Engine
from oracle import oracle as O
eng = Engine(0)
t0 = time.time()
hist, st = eng.group_from_fasta(genomes, k)
ha, sa = eng.across_groups()
print(f"gpu: {time.time() - t0:.2f}s wall, device {st['ms_total']:.1f} ms, windows {st['windows']}, S_G {st['genome_distinct']}, D_G {st['distinct']}, passes {st['passes_genome']}+{st['passes_group']}", flush=True)
t0 = time.time()
w, a, so = O.exp1(genomes, [0] * ng, 1, k)
print(f"oracle: {time.time() - t0:.1f}s on {O.num_threads()} threads", flush=True)
assert np.array_equal(hist, w[0]), "within-group histogram differs"
assert np.array_equal(ha, a), "across histogram differs"
assert st["distinct"] == so["sum_group_distinct"]
print("OK: histograms identical; nonzero bins:", {int(i): int(hist[i]) for i in np.flatnonzero(hist)[:8]}, "...")
