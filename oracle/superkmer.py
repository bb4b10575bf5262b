"""TEST INFRASTRUCTURE (like everything under oracle/): the numpy statement of the FIRST pass of the minimizer-bin group stage
(khoice_b200/csrc/bins.cu) -- partition the k-mer windows of a genome by minimizer into bins, as super-k-mer records -- so that the
CUDA partition can be checked on its own, bit for bit, below the level of the histograms.  Nothing under khoice_b200/ uses this.

Semantics of the windows are the oracle's (rules R1-R5 of SURVEY.md 8c, the reference's `kmc -fm` call sites
/root/reference/workflow/rules/exp_type_1.smk:156-163): symbols `ACGTacgt` are valid, anything else breaks the window,
windows never span `>` records, the k-mer value is base 4 with the first base most significant, canonical = min(k-mer,
reverse complement).  On top of that, for a window with canonical m-mers c_0 .. c_{k-m} (m = 13, 13 <= k <= 64):

    minimizer hash   H = min_j  mix32(c_j)          32-bit: x *= 0x9E3779B1; x ^= x >> 15; x *= 0x85EBCA77; x ^= x >> 13;
                                                    x *= 0xC2B2AE3D; x ^= x >> 16; x |= 1   (0 is kept for "no m-mer")
    bin              b = (B * n_bins) >> 32          B = H * 0xD6E8FEB9; B ^= B >> 16; B *= 0x7FEB352D   (mod 2^32)
    super-k-mer      a maximal run of CONSECUTIVE windows (symbol positions i, i+1, ...) that are all valid and share H,
    record           ... cut at every multiple of 4096 window starts (the kernel's tiles) and into pieces of <= 32 windows

The bin is a function of the window's sequence content up to strand -- the same k-mer lands in the same bin in every
genome and on both strands -- which is what lets a bin be counted independently of all others."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

U = np.uint64
M = 13
TILE = 4096
CAPW = 32
_M32 = U(0xFFFFFFFF)


def mix32(x: np.ndarray) -> np.ndarray:
    x = (x.astype(U) * U(0x9E3779B1)) & _M32
    x ^= x >> U(15)
    x = (x * U(0x85EBCA77)) & _M32
    x ^= x >> U(13)
    x = (x * U(0xC2B2AE3D)) & _M32
    x ^= x >> U(16)
    return x | U(1)


def bin_of(minhash: np.ndarray, n_bins: int) -> np.ndarray:
    b = (minhash.astype(U) * U(0xD6E8FEB9)) & _M32
    b ^= b >> U(16)
    b = (b * U(0x7FEB352D)) & _M32
    return ((b * U(n_bins)) >> U(32)).astype(np.int64)


def symbol_stream(fasta: bytes) -> Tuple[np.ndarray, np.ndarray]:
    """(codes uint8 [n_sym], valid bool [n_sym]): one symbol per sequence character outside header lines (newlines and
    carriage returns skipped) and one INVALID break symbol per '>' (so that windows never span records) -- the layout K1 packs."""
    a = np.frombuffer(bytes(fasta), dtype=np.uint8)
    n = a.size
    if n == 0:
        return np.zeros(0, np.uint8), np.zeros(0, bool)
    is_nl = a == 10
    # header state: a '>' opens a header only outside a header; scan events in order
    ev_pos = np.flatnonzero((a == 62) | is_nl)
    in_hdr = np.zeros(n, dtype=bool)
    opens = np.zeros(n, dtype=bool)
    state = False
    start = 0
    for p in ev_pos:                       # events are few (one per line); the per-byte work below is vectorised
        if state:
            if a[p] == 10:
                in_hdr[start:p + 1] = True
                state = False
        elif a[p] == 62:
            state, start = True, p
            opens[p] = True
    if state:
        in_hdr[start:] = True
    keep = opens | (~in_hdr & ~is_nl & (a != 13))
    sym = a[keep]
    code = np.full(sym.size, 255, dtype=np.uint8)
    for ch, v in zip(b"ACGTacgt", (0, 1, 2, 3, 0, 1, 2, 3)):
        code[sym == ch] = v
    code[opens[keep]] = 255
    valid = code != 255
    code[~valid] = 0
    return code, valid


def _canonical_values(code: np.ndarray, length: int) -> np.ndarray:
    """canonical value of the `length`-mer starting at every symbol position (garbage where the stretch is not valid)."""
    n = code.size - length + 1
    if n <= 0:
        return np.zeros(0, U)
    c = code.astype(U)
    fwd = np.zeros(n, dtype=U)
    rc = np.zeros(n, dtype=U)
    for j in range(length):
        fwd = (fwd << U(2)) | c[j:n + j]
        rc = rc | ((U(3) - c[j:n + j]) << U(2 * j))
    return np.minimum(fwd, rc)


def _all_valid(valid: np.ndarray, length: int) -> np.ndarray:
    n = valid.size - length + 1
    if n <= 0:
        return np.zeros(0, bool)
    cs = np.concatenate([[0], np.cumsum(valid.astype(np.int64))])
    return (cs[length:length + n] - cs[:n]) == length


def window_bins(fasta: bytes, k: int, n_bins: int, m: int = M):
    """Per symbol position i (a window start): (ok bool, canonical k-mer, minimizer hash uint64, bin int64); entries with ok False
    are undefined.  The k-mer column is uint64 [n] for k <= 32 and uint64 [n, 2] (lo, hi) for 33 <= k <= 64 (the oracle's own
    layout, taken from oracle.kmers, which lists the valid windows in stream order)."""
    if not (1 <= m <= min(k, 15) and k <= 64 and n_bins >= 1):
        raise ValueError("1 <= m <= min(k, 15), k <= 64 and n_bins >= 1")
    code, valid = symbol_stream(fasta)
    ok = _all_valid(valid, k)
    n = ok.size
    if n == 0:
        return ok, np.zeros((0,) if k <= 32 else (0, 2), U), np.zeros(0, U), np.zeros(0, np.int64)
    if k <= 32:
        kmer = _canonical_values(code, k)
    else:
        from . import oracle as O
        keys, _ = O.kmers(fasta, k)
        kmer = np.zeros((n, 2), dtype=U)
        kmer[ok] = keys
    hm = mix32(_canonical_values(code, m))        # per m-mer position; only positions inside valid windows are ever used
    w = k - m + 1
    best = hm[:n].copy()
    for j in range(1, w):
        best = np.minimum(best, hm[j:n + j])
    return ok, kmer, best, bin_of(best, n_bins)


def records(fasta: bytes, k: int, n_bins: int, first_symbol: int = 0, m: int = M, tile: int = TILE, capw: int = CAPW) -> List[Tuple[int, int, int]]:
    """[(bin, first symbol position, number of windows)] in stream order: the records the partition kernel writes for this genome when
    its symbols start at position `first_symbol` of the group's stream (runs are cut where first_symbol + position is a multiple of
    `tile`, and into pieces of at most `capw` windows)."""
    ok, _, mh, bins = window_bins(fasta, k, n_bins, m)
    idx = np.flatnonzero(ok)
    if idx.size == 0:
        return []
    brk = np.ones(idx.size, dtype=bool)
    brk[1:] = (np.diff(idx) != 1) | (mh[idx][1:] != mh[idx][:-1])
    if tile > 0:
        brk |= ((idx + first_symbol) % tile) == 0
    starts = idx[brk]
    ends = np.concatenate([idx[np.flatnonzero(brk)[1:] - 1], idx[-1:]])
    out = []
    for s0, e0 in zip(starts.tolist(), ends.tolist()):
        n = e0 - s0 + 1
        for p in range(0, n, capw):
            out.append((int(bins[s0]), s0 + p, min(capw, n - p)))
    return out


def region_counts(genomes, first_symbols, k: int, n_bins: int):
    """(records, windows), each int64 [n_bins, ceil(len(genomes) / 64)]: what the partition kernel leaves in every region (one region per
    bin and chunk of 64 genomes) for a group whose genome g starts at symbol first_symbols[g] of the group's stream."""
    chunks = (len(genomes) + 63) // 64
    rec = np.zeros((n_bins, max(chunks, 1)), dtype=np.int64)
    win = np.zeros((n_bins, max(chunks, 1)), dtype=np.int64)
    for g, text in enumerate(genomes):
        for b, _, n in records(text, k, n_bins, int(first_symbols[g])):
            rec[b, g // 64] += 1
            win[b, g // 64] += n
    return rec, win


def superkmers(fasta: bytes, k: int, n_bins: int, m: int = M) -> List[Tuple[int, int, int]]:
    """[(bin, first symbol position, number of windows)]: the maximal runs themselves (no tile cuts, no length limit)."""
    return records(fasta, k, n_bins, 0, m, tile=0, capw=1 << 30)


def binned_group_histogram(genomes, k: int, n_bins: int, nbins: int = 5000, m: int = M):
    """The group stage computed BIN BY BIN: hist[c] = number of distinct k-mers found in exactly c genomes, and the group's
    distinct k-mers per bin.  Must equal oracle.exp1's within-group histogram (tests/test_superkmer_oracle.py)."""
    per_bin = {}
    uniq = (lambda a: np.unique(a)) if k <= 32 else (lambda a: np.unique(a, axis=0))
    for g, text in enumerate(genomes):
        ok, kmer, _, bins = window_bins(text, k, n_bins, m)
        for b in np.unique(bins[ok]):
            per_bin.setdefault(int(b), []).append((g, uniq(kmer[ok & (bins == b)])))
    hist = np.zeros(nbins + 1, dtype=np.uint64)
    sets = {}
    for b, parts in per_bin.items():
        cat = np.concatenate([p[1] for p in parts])                                        # each part is one genome's distinct keys
        keys, cnt = np.unique(cat, return_counts=True) if k <= 32 else np.unique(cat, axis=0, return_counts=True)
        np.add.at(hist, np.minimum(cnt, nbins), 1)
        sets[b] = keys
    return hist, sets


def binned_across_histogram(group_bin_sets, k: int, nbins: int = 5000):
    """The across-group stage computed BIN BY BIN: group_bin_sets[g] = {bin: distinct k-mers of group g in that bin} (the second
    result of binned_group_histogram).  Because the bin is a function of the k-mer alone, every copy of a k-mer sits in the same
    bin in every group: hist[c] = number of k-mers found in exactly c groups needs no exchange between bins."""
    hist = np.zeros(nbins + 1, dtype=np.uint64)
    total = 0
    for b in sorted(set().union(*[set(s) for s in group_bin_sets])):
        parts = [s[b] for s in group_bin_sets if b in s]
        cat = np.concatenate(parts)
        _, cnt = np.unique(cat, return_counts=True) if k <= 32 else np.unique(cat, axis=0, return_counts=True)
        np.add.at(hist, np.minimum(cnt, nbins), 1)
        total += cnt.size
    return hist, total
