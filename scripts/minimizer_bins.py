"""Feasibility numbers for the design DESIGN.md section 7 names as the next step: partition the windows of a group by
MINIMIZER into bins that one CTA can count in shared memory.  Host-side experiment (numpy): for g genomes of one synthetic
group, the canonical-m-mer minimizer of every k-mer window (hashed, so that poly-A does not dominate), bins = minimizer
hash mod B.  Prints bin-size skew (windows per bin), super-k-mer length, and distinct k-mers per bin.
usage: python scripts/minimizer_bins.py [genomes=8] [k=31] [m=11] [log2_bins=14]"""
import os
import sys

import numpy as np
from scipy.ndimage import minimum_filter1d

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from khoice_b200 import synth  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
k = int(sys.argv[2]) if len(sys.argv) > 2 else 31
m = int(sys.argv[3]) if len(sys.argv) > 3 else 11
LB = int(sys.argv[4]) if len(sys.argv) > 4 else 14
B = 1 << LB
cfg = synth.SynthConfig(n_groups=1, genomes_per_group=G, genome_len=5_000_000)
MASK = np.uint64((1 << (2 * m)) - 1)


def mix(x):
    x = (x ^ (x >> np.uint64(15))) * np.uint64(0x9E3779B97F4A7C15)
    return x ^ (x >> np.uint64(29))


bins_all, keys_all, slens = [], [], []
for g in range(1, G + 1):
    seq = synth._clean_sequence(synth.make_genome(cfg, 1, g))          # ACGT only (N runs dropped: fine for statistics)
    code = np.searchsorted(np.frombuffer(b"ACGT", np.uint8), seq).astype(np.uint64)
    n = code.size
    fwd = np.zeros(n - m + 1, dtype=np.uint64)
    rc = np.zeros(n - m + 1, dtype=np.uint64)
    for j in range(m):
        fwd = (fwd << np.uint64(2)) | code[j:n - m + 1 + j]
        rc = rc | ((np.uint64(3) - code[j:n - m + 1 + j]) << np.uint64(2 * j))
    h = mix(np.minimum(fwd, rc) & MASK)
    w = k - m + 1                                                       # m-mers per k-mer window
    mn = minimum_filter1d(h, size=w, mode="nearest", origin=0)
    # window i covers m-mers [i, i + w): minimum_filter1d is centred -> shift
    half = w // 2
    win_min = mn[half:half + (n - k + 1)]
    b = (((win_min * np.uint64(0xD6E8FEB86659FD93)) >> np.uint64(64 - LB))).astype(np.int64)   # the minimum itself is biased: re-mix, take top bits
    bins_all.append(b)
    change = np.flatnonzero(np.diff(win_min) != 0)
    slens.append(np.diff(np.concatenate([[0], change + 1, [win_min.size]])))
    # canonical k-mer hash stand-in for distinct counting: 64-bit hash of (fwd/rc k-mer) via rolling polynomial is overkill;
    # use the pair (minimizer, position-independent 64-bit hash of the k-mer) computed from two 32-base halves
    kf = np.zeros(n - k + 1, dtype=np.uint64)
    kr = np.zeros(n - k + 1, dtype=np.uint64)
    for j in range(k):
        kf = (kf << np.uint64(2)) | code[j:n - k + 1 + j]
        kr = kr | ((np.uint64(3) - code[j:n - k + 1 + j]) << np.uint64(2 * j))
    keys_all.append(np.minimum(kf, kr))
bins = np.concatenate(bins_all)
keys = np.concatenate(keys_all)
cnt = np.bincount(bins, minlength=B)
sl = np.concatenate(slens)
print(f"{G} genomes x 5 Mbp, k={k}, m={m}, {B} bins: {bins.size} windows")
print(f"windows per bin: mean {cnt.mean():.0f}, median {np.median(cnt):.0f}, p99 {np.percentile(cnt, 99):.0f}, max {cnt.max()} ({cnt.max() / cnt.mean():.1f} x mean)")
print(f"super-k-mers: {sl.size} ({bins.size / sl.size:.1f} windows each on average; {(2 * (sl + k - 1)).sum() / 8 / bins.size:.2f} bytes per window at 2 bits per base)")
order = np.argsort(bins, kind="stable")
sb, sk = bins[order], keys[order]
pair_new = np.ones(sb.size, dtype=bool)
o2 = np.lexsort((sk, sb))
sb2, sk2 = sb[o2], sk[o2]
new = np.ones(sb2.size, dtype=bool)
new[1:] = (sb2[1:] != sb2[:-1]) | (sk2[1:] != sk2[:-1])
dcnt = np.bincount(sb2[new], minlength=B)
print(f"distinct k-mers per bin: mean {dcnt.mean():.0f}, p99 {np.percentile(dcnt, 99):.0f}, max {dcnt.max()}; overall {int(new.sum())} distinct of {bins.size} windows")
scale = 50 / G
print(f"scaled to 50 genomes: ~{cnt.mean() * scale:.0f} windows per bin (max ~{cnt.max() * scale:.0f}); a 16-byte (key, genome bits) table entry per distinct "
      f"k-mer: mean ~{dcnt.mean() * (1 + (50 - G) * 0.1 / (1 + (G - 1) * 0.1)) * 16 / 1024:.0f} KiB per bin")
