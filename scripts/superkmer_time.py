"""Times of the EXPERIMENTAL minimizer-bin group stage (csrc/superkmer.cu) on one config-2 group (50 x 5 Mbp), k = 31,
next to the product path's time for the same group."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiprocessing as mp
import numpy as np
from khoice_b200 import synth
from khoice_b200.engine import Engine
def gen(i):
    return synth.make_genome(synth.SynthConfig(n_groups=1, genomes_per_group=50, genome_len=5_000_000), 1, i)
with mp.get_context("fork").Pool(16) as pool:
    files = pool.map(gen, range(1, 51))
eng = Engine(0)
for _ in range(2):
    eng.group_sets_reset()
    ref, rst = eng.group_from_fasta(files, 31, nbins=64, keep_set=False)
print(f"product path: extract {rst['ms_extract']:.2f} + sort {rst['ms_sort1']:.2f} + count {rst['ms_count']:.2f} ms (pack {rst['ms_pack']:.2f}); distinct {rst['distinct']}", flush=True)
for m, lb, compact in ((11, 16, False), (11, 16, True), (13, 16, True)):
    for _ in range(2):
        hist, st = eng.superkmer_group(files, 31, m, lb, nbins=64, compact=compact)
    print(f"m={m} bins=2^{lb} {'super-k-mer records' if compact else 'expanded k-mers'}: count pass {st['ms_count']:.2f} ms (incl. host scan), scatter {st['ms_scatter']:.2f} ms, bins {st['ms_bins']:.2f} ms; distinct {st['distinct']}, "
          f"overflowed bins {st['overflowed_bins']}, histogram equal to the product path: {bool(np.array_equal(hist, ref))}", flush=True)
os.environ["KHB_SUPERKMER_ONEPASS"] = "1"
for _ in range(2):
    hist, st = eng.superkmer_group(files, 31, 11, 16, nbins=64, compact=True)
print(f"ONE pass (fixed bin regions, no count pass): setup {st['ms_count']:.2f} ms, partition {st['ms_scatter']:.2f} ms, bins {st['ms_bins']:.2f} ms; distinct {st['distinct']}, "
      f"overflow {st['overflowed_bins']}, histogram equal to the product path: {bool(np.array_equal(hist, ref))}", flush=True)
eng.close()
