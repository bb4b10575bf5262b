"""Pure-Python, string-level restatement of the exp-1 k-mer semantics (TEST INFRASTRUCTURE).

Deliberately naive: it works on Python strings, exactly as the reference's own in-tree statement of
a canonical k-mer does (/root/reference/src/merge_lists.py:60-73), so that it shares no code path
with the C oracle (oracle/kmer_oracle.c) or the CUDA kernels.  Small inputs only.
PARITY UNPINNED (no KMC binary, no reference fixtures) -- see oracle/kmer_oracle.c.
"""
from __future__ import annotations

from collections import Counter

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def canonical(kmer: str) -> str:
    """Lexicographic min of the k-mer and its reverse complement, A<C<G<T (merge_lists.py:60-73)."""
    rc = "".join(_COMP[ch] for ch in reversed(kmer))
    return kmer if kmer < rc else rc


def records(fasta: bytes):
    """R1, R2, R4: split FASTA text into per-record sequence strings (headers dropped, \\n and \\r
    skipped).  Text before the first '>' is a record of its own."""
    recs, cur, in_header = [], [], False
    for b in fasta:
        ch = chr(b)
        if in_header:
            if ch == "\n":
                in_header = False
            continue
        if ch == ">":
            recs.append("".join(cur))
            cur = []
            in_header = True
            continue
        if ch in "\n\r":
            continue
        cur.append(ch)
    recs.append("".join(cur))
    return recs


def kmers(fasta: bytes, k: int):
    """Canonical k-mers (upper-case strings) of every valid window, input order (R3, R5)."""
    out = []
    for rec in records(fasta):
        s = rec.upper()
        for i in range(len(s) - k + 1):
            w = s[i:i + k]
            if all(ch in "ACGT" for ch in w):
                out.append(canonical(w))
    return out


def n_symbols(fasta: bytes) -> int:
    return sum(len(r) for r in records(fasta))


def genome_set(fasta: bytes, k: int):
    """R6: distinct canonical k-mers, sorted (lexicographic == numeric with A<C<G<T)."""
    return sorted(set(kmers(fasta, k)))


def union_sum(sets, cs: int = 5000):
    """R7: union with counter sum, saturating at cs. Returns sorted [(kmer, count)]."""
    c = Counter()
    for s in sets:
        c.update(s)
    return [(x, min(c[x], cs)) for x in sorted(c)]


def histogram(table, nbins: int = 5000):
    """R8: list h with h[c] = number of k-mers whose counter is c (h[0] unused)."""
    h = [0] * (nbins + 1)
    for _, c in table:
        if c <= nbins:
            h[c] += 1
    return h


def encode(kmer: str) -> int:
    """Base-4 value, first base most significant (R5)."""
    v = 0
    for ch in kmer:
        v = (v << 2) | "ACGT".index(ch)
    return v


def decode(v: int, k: int) -> str:
    return "".join("ACGT"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def exp1(groups, k: int, cs: int = 5000, nbins: int = 5000):
    """groups: list of lists of FASTA bytes. Returns (within hists, across hist, group tables)."""
    within, group_sets, tables = [], [], []
    for genomes in groups:
        t = union_sum([genome_set(g, k) for g in genomes], cs)
        tables.append(t)
        within.append(histogram(t, nbins))
        group_sets.append([x for x, _ in t])
    across_table = union_sum(group_sets, cs)
    return within, histogram(across_table, nbins), tables


def exp2(groups, pivots, k: int, cs: int = 5000, nbins: int = 5000):
    """Experiment type 2 (exp_type_2.smk:297-508) with Python sets and dicts.  groups: rest_of_set FASTA texts per
    dataset; pivots: one FASTA text per dataset.  Returns (within, across): per dataset (sub_hist, inter_hist) where
    sub = `pivot kmers_subtract union` (all counters 1) and inter = `pivot intersect union -ocsum` (1 + union counter)."""
    unions, psets = [], []
    for genomes in groups:
        unions.append(dict(union_sum([genome_set(g, k) for g in genomes], cs)))
    for p in pivots:
        psets.append(genome_set(p, k))

    def versus(pset, table):
        sub = [(x, 1) for x in pset if x not in table]
        inter = [(x, min(1 + table[x], cs)) for x in pset if x in table]
        return histogram(sub, nbins), histogram(inter, nbins)

    within = [versus(psets[d], unions[d]) for d in range(len(groups))]
    across = []
    for d in range(len(groups)):
        others = dict(union_sum([sorted(unions[i]) for i in range(len(groups)) if i != d], cs))
        across.append(versus(psets[d], others))
    return within, across
