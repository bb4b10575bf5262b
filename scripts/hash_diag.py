"""Timing experiments for hash_insert_kernel (KHB_HASH_DIAG variants give wrong results by design): one config-2 group."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from khoice_b200 import synth
from khoice_b200.engine import Engine
import multiprocessing as mp
def gen(i):
    cfg = synth.SynthConfig(n_groups=1, genomes_per_group=50, genome_len=5_000_000)
    return synth.make_genome(cfg, 1, i)
with mp.get_context("fork").Pool(16) as pool:
    files = pool.map(gen, range(1, 51))
eng = Engine(0)
eng.set_group_mode("hash")
st = eng.stage_fasta(files)
for _ in range(2):
    eng.group_sets_reset(); eng.group_from_staged(st, 31)
eng.profile_enable(True)
for _ in range(3):
    eng.group_sets_reset(); h, s = eng.group_from_staged(st, 31)
p = eng.profile_read()
print(os.environ.get("KHB_HASH_DIAG", "0"), os.environ.get("KHB_HASH_ORDER", "interleave"), "insert ms/launch", p["hash_insert"]["ms"] / 3, "count ms/launch", p["hash_count"]["ms"] / 3,
      "distinct", s["distinct"], "genome_distinct", s["genome_distinct"], "windows", s["windows"])
eng.close()
