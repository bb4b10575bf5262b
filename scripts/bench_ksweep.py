"""BASELINE config 3: k sweep (k = 7, 9, ..., 31) over the 10 x 50 x 5 Mbp set with the 2-bit stream packed ONCE and kept
in HBM (khb_pack_group / khb_group_from_packed).  Prints per-k device time and the sweep total.
usage: python scripts/bench_ksweep.py [groups] [genomes] [kmin] [kmax]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (generate_groups)
from khoice_b200 import synth  # noqa: E402
from khoice_b200.engine import Engine  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
kmin = int(sys.argv[3]) if len(sys.argv) > 3 else 7
kmax = int(sys.argv[4]) if len(sys.argv) > 4 else 31
cfg = synth.SynthConfig(n_groups=G, genomes_per_group=N, genome_len=5_000_000)
groups = bench.generate_groups(cfg, list(range(1, G + 1)), min(16, os.cpu_count() or 1))
eng = Engine(0)
t0 = time.time()
packed = {g: eng.pack_group(groups[g]) for g in sorted(groups)}
eng.sync()
t_pack = time.time() - t0
bases = sum(pk.info()["bases"] for pk in packed.values())
print(f"packed {G} groups x {N} genomes: {bases / 1e9:.3f} Gbases, {sum(pk.info()['device_bytes'] for pk in packed.values()) / 1e9:.2f} GB resident, {t_pack:.2f} s (incl. H2D from pageable memory)")
total = 0.0
for rep in range(2):  # first sweep warms the scratch allocations up
    total = 0.0
    for k in range(kmin, kmax + 1, 2):
        t0 = time.time()
        eng.group_sets_reset()
        ms = 0.0
        for g in sorted(packed):
            _, st = eng.group_from_packed(packed[g], k)
            ms += st["ms_total"]
        _, st = eng.across_groups()
        ms += st["ms_total"]
        wall = time.time() - t0
        total += wall
        if rep == 1:
            print(f"k={k:2d}: device {ms:8.2f} ms, wall {wall * 1e3:8.2f} ms -> {bases / wall / 1e9:6.2f} Gbases/s (passes per group {st['passes_group']})")
print(f"sweep of {(kmax - kmin) // 2 + 1} k values: {total:.3f} s wall = {bases * ((kmax - kmin) // 2 + 1) / total / 1e9:.2f} Gbase-k/s")
eng.close()
