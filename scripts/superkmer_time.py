"""Time of the EXPERIMENTAL minimizer-partition count pass (csrc/superkmer.cu) on one config-2 group (50 x 5 Mbp), k = 31."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiprocessing as mp
from khoice_b200 import synth
from khoice_b200.engine import Engine
def gen(i):
    return synth.make_genome(synth.SynthConfig(n_groups=1, genomes_per_group=50, genome_len=5_000_000), 1, i)
with mp.get_context("fork").Pool(16) as pool:
    files = pool.map(gen, range(1, 51))
eng = Engine(0)
st = eng.stage_fasta(files)
pk = eng.pack_fasta(st)
for m, lb in ((11, 16), (11, 14), (9, 16)):
    for _ in range(2):
        win, sk, ms = eng.superkmer_count(pk, 31, m, lb)
    print(f"m={m} bins=2^{lb}: {ms:.2f} ms for {pk['n_symbols']} symbols; windows {int(win.sum())}, super-k-mers {int(sk.sum())} ({win.sum() / max(sk.sum(), 1):.1f} windows each), "
          f"largest bin {int(win.max())} windows = {win.max() / win.mean():.1f} x mean", flush=True)
eng.close()
