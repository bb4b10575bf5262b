"""Informational end-to-end number of SURVEY.md section 8d (ii): from `.fna.gz` files on disk (host inflate included) to the
step_5 / step_9 CSVs, through the fused host driver (khoice_b200/pipeline.py).  KMC's own time includes the inflate too.
usage: python scripts/bench_from_disk.py [groups] [genomes] [k,k,...] [work_root]"""
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiprocessing import get_context  # noqa: E402

from khoice_b200 import pipeline, synth  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ks = sys.argv[3].split(",") if len(sys.argv) > 3 else ["31"]
root = sys.argv[4] if len(sys.argv) > 4 else tempfile.mkdtemp(prefix="khb_disk_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
cfg = synth.SynthConfig(n_groups=G, genomes_per_group=N, genome_len=5_000_000)


def write_group(g):
    import gzip
    d = os.path.join(root, "data", f"dataset_{g}")
    os.makedirs(d, exist_ok=True)
    bases = 0
    for i in range(1, N + 1):
        t = synth.make_genome(cfg, g, i)
        bases += synth.count_bases(t)
        with gzip.open(os.path.join(d, synth.genome_name(g, i) + ".fna.gz"), "wb", compresslevel=1) as fd:
            fd.write(t)
    return bases


t0 = time.time()
with get_context("fork").Pool(min(16, G)) as pool:
    bases = sum(pool.map(write_group, range(1, G + 1)))
gz = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(os.path.join(root, "data")) for f in fs)
print(f"wrote {G} x {N} genomes to {root}: {bases / 1e9:.3f} Gbases, {gz / 1e9:.2f} GB of .fna.gz, {time.time() - t0:.1f} s", flush=True)
from khoice_b200.engine import Engine  # noqa: E402
eng = Engine(0)
for rep in range(2):  # the second run has warm scratch allocations and a warm page cache, like a k sweep in progress
    for d in os.listdir(root):
        if d != "data":
            shutil.rmtree(os.path.join(root, d), ignore_errors=True)
    t0 = time.time()
    r = pipeline.run_fused(root, G, ks, engine=eng, stubs=True)
    dt = time.time() - t0
    print(f"run {rep}: {dt:.2f} s wall for k = {','.join(ks)} -> {bases * len(ks) / dt / 1e9:.2f} Gbase-k/s from .fna.gz to CSVs "
          f"({os.cpu_count()} host cores, inflate threads {os.environ.get('KHB_INFLATE_THREADS', 'default')})", flush=True)
eng.close()
shutil.rmtree(root, ignore_errors=True)
