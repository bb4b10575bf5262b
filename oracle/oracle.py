"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- see oracle/kmer_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  PARITY UNPINNED: KMC 3.2.1 is neither in /root/reference nor installed; see the C header.

K-mer words: k <= 32 -> numpy uint64 [n];  k <= 64 -> numpy uint64 [n, 2] with column 0 = low word,
column 1 = high word (the in-memory layout of a little-endian 128-bit integer).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkmer_oracle.so")
_lib = None

CS_DEFAULT = 5000  # `-cs5000`, /root/reference/workflow/rules/exp_type_1.smk:61,84
NBINS_DEFAULT = 5000  # R8: histogram rows 1..5000


def build(force: bool = False) -> str:
    """Compile oracle/libkmer_oracle.so with the committed Makefile."""
    src_mtime = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("kmer_oracle.c", "ko_body.inc"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_mtime:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ko_kmers.restype = C.c_int64
        L.ko_kmers.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.ko_sort_unique.restype = C.c_int64
        L.ko_sort_unique.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        L.ko_union_sum.restype = C.c_int64
        L.ko_union_sum.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_uint32]
        L.ko_histogram.restype = None
        L.ko_histogram.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ko_exp1.restype = C.c_int
        L.ko_exp1.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32,
                              C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ko_num_threads.restype = C.c_int
        L.ko_set_num_threads.restype = None
        L.ko_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def key_shape(n: int, k: int):
    return (n,) if k <= 32 else (n, 2)


def _as_u8(fasta) -> np.ndarray:
    if isinstance(fasta, (bytes, bytearray, memoryview)):
        return np.frombuffer(fasta, dtype=np.uint8)
    return np.ascontiguousarray(fasta, dtype=np.uint8)


def kmers(fasta, k: int):
    """All canonical k-mers of valid windows, input order (rules R1-R5). Returns (keys, n_symbols)."""
    buf = _as_u8(fasta)
    cap = buf.size + 1
    out = np.empty(key_shape(cap, k), dtype=np.uint64)
    nsym = C.c_uint64(0)
    n = lib().ko_kmers(buf.ctypes.data, buf.size, k, out.ctypes.data, cap, C.byref(nsym))
    if n < 0:
        raise ValueError(f"ko_kmers failed: {n}")
    return out[:n].copy(), int(nsym.value)


def sort_unique(keys: np.ndarray, k: int) -> np.ndarray:
    """Sorted distinct k-mers (rule R6)."""
    a = np.ascontiguousarray(keys, dtype=np.uint64).copy()
    n = lib().ko_sort_unique(a.ctypes.data, a.shape[0], k)
    if n < 0:
        raise MemoryError
    return a[:n].copy()


def genome_set(fasta, k: int) -> np.ndarray:
    """step_1 + step_2 of the reference: distinct canonical k-mers of one genome, sorted."""
    return sort_unique(kmers(fasta, k)[0], k)


def union_sum(sets, k: int, cs: int = CS_DEFAULT):
    """`kmc_tools complex (set1 + ... + setN) -cs{cs}` (rule R7). Returns (sorted keys, uint32 counts)."""
    sets = [np.ascontiguousarray(s, dtype=np.uint64) for s in sets]
    if len(sets) == 0 or sum(s.shape[0] for s in sets) == 0:
        return np.empty(key_shape(0, k), np.uint64), np.empty(0, np.uint32)
    cat = np.concatenate(sets, axis=0).copy()
    counts = np.empty(cat.shape[0], dtype=np.uint32)
    n = lib().ko_union_sum(cat.ctypes.data, cat.shape[0], k, counts.ctypes.data, cs)
    if n < 0:
        raise MemoryError
    return cat[:n].copy(), counts[:n].copy()


def histogram(counts: np.ndarray, nbins: int = NBINS_DEFAULT) -> np.ndarray:
    """`kmc_tools transform X histogram` (rule R8): uint64[nbins+1], index c = #k-mers with counter c."""
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    h = np.zeros(nbins + 1, dtype=np.uint64)
    lib().ko_histogram(counts.ctypes.data, counts.size, h.ctypes.data, nbins)
    return h


def exp1(genomes, group_of, n_groups: int, k: int, cs: int = CS_DEFAULT, nbins: int = NBINS_DEFAULT):
    """Whole exp-1 arithmetic for one k on the host cores (OpenMP).

    genomes: list of bytes / uint8 arrays (FASTA text); group_of: list of 0-based group ids.
    Returns (within_hist [n_groups, nbins+1], across_hist [nbins+1], stats dict)."""
    bufs = [_as_u8(g) for g in genomes]
    n = len(bufs)
    ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data for b in bufs])
    lens = (C.c_size_t * max(n, 1))(*[b.size for b in bufs])
    grp = np.ascontiguousarray(group_of, dtype=np.int32)
    within = np.zeros((n_groups, nbins + 1), dtype=np.uint64)
    across = np.zeros(nbins + 1, dtype=np.uint64)
    stats = np.zeros(5, dtype=np.uint64)
    rc = lib().ko_exp1(ptrs, lens, grp.ctypes.data, n, n_groups, k, cs, nbins, within.ctypes.data,
                       across.ctypes.data, stats.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"ko_exp1 failed: {rc}")
    names = ("symbols", "valid_kmers", "sum_genome_distinct", "sum_group_distinct", "distinct")
    return within, across, dict(zip(names, (int(x) for x in stats)))


def _row_ids(arrays, k: int):
    """Map k-mer words of several arrays to dense integer ids that compare like the k-mers (k <= 32: the values
    themselves; k <= 64: ranks of the (hi, lo) pairs)."""
    if k <= 32:
        return [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1) for a in arrays]
    cat = np.concatenate([np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 2) for a in arrays], axis=0)
    if cat.shape[0] == 0:
        return [np.empty(0, np.uint64) for _ in arrays]
    order = np.lexsort((cat[:, 0], cat[:, 1]))  # by hi, then lo
    srt = cat[order]
    new = np.ones(srt.shape[0], dtype=bool)
    new[1:] = np.any(srt[1:] != srt[:-1], axis=1)
    ranks = np.empty(cat.shape[0], dtype=np.uint64)
    ranks[order] = (np.cumsum(new) - 1).astype(np.uint64)
    out, at = [], 0
    for a in arrays:
        n = np.asarray(a).reshape(-1, 2).shape[0]
        out.append(ranks[at:at + n])
        at += n
    return out


def simple_intersect_ocsum(a_keys, a_counts, b_keys, b_counts, k: int, cs: int = CS_DEFAULT):
    """`kmc_tools simple A B intersect O -ocsum` (/root/reference/workflow/rules/exp_type_2.smk:354-365, 470-481):
    k-mers present in both inputs, counter = c_A + c_B saturated at cs (rule R10, [KMC-ext]: the saturation value of
    the output is not visible in the reference; every counter on this path is far below either candidate).
    Inputs sorted and duplicate-free.  Returns (keys, uint32 counts)."""
    ia, ib = _row_ids([a_keys, b_keys], k)
    pos = np.searchsorted(ib, ia)
    pos_c = np.minimum(pos, max(ib.shape[0] - 1, 0))
    hit = (pos < ib.shape[0]) & (ib[pos_c] == ia) if ib.shape[0] else np.zeros(ia.shape[0], dtype=bool)
    cnt = np.minimum(np.asarray(a_counts, dtype=np.uint64)[hit] + np.asarray(b_counts, dtype=np.uint64)[pos_c[hit]], cs).astype(np.uint32)
    return np.asarray(a_keys)[hit], cnt


def simple_kmers_subtract(a_keys, a_counts, b_keys, k: int):
    """`kmc_tools simple A B kmers_subtract O` (exp_type_2.smk:367-380, 483-496): k-mers of A absent from B, with A's
    counters."""
    ia, ib = _row_ids([a_keys, b_keys], k)
    pos = np.searchsorted(ib, ia)
    pos_c = np.minimum(pos, max(ib.shape[0] - 1, 0))
    hit = (pos < ib.shape[0]) & (ib[pos_c] == ia) if ib.shape[0] else np.zeros(ia.shape[0], dtype=bool)
    return np.asarray(a_keys)[~hit], np.asarray(a_counts, dtype=np.uint32)[~hit]


def exp2(groups, pivots, k: int, cs: int = CS_DEFAULT, nbins: int = NBINS_DEFAULT):
    """Whole exp-2 arithmetic for one k (rule chain of exp_type_2.smk:297-508), built from the same primitives.

    groups: list of lists of FASTA texts (rest_of_set of every dataset); pivots: one FASTA text per dataset.
    Returns (within, across), each uint64 [n_datasets, 2, nbins+1]: [:, 0] = histogram of the kmers_subtract result,
    [:, 1] = histogram of the intersect -ocsum result."""
    n = len(groups)
    within = np.zeros((n, 2, nbins + 1), dtype=np.uint64)
    across = np.zeros((n, 2, nbins + 1), dtype=np.uint64)
    unions, psets = [], []
    for d in range(n):
        pset = genome_set(pivots[d], k)
        ones = np.ones(pset.shape[0], dtype=np.uint32)
        ukeys, ucnt = union_sum([genome_set(g, k) for g in groups[d]], k, cs)
        _, c = simple_intersect_ocsum(pset, ones, ukeys, ucnt, k, cs)
        within[d, 1] = histogram(c, nbins)
        _, c = simple_kmers_subtract(pset, ones, ukeys, k)
        within[d, 0] = histogram(c, nbins)
        unions.append(ukeys)
        psets.append(pset)
    for d in range(n):
        ones = np.ones(psets[d].shape[0], dtype=np.uint32)
        okeys, ocnt = union_sum([unions[i] for i in range(n) if i != d], k, cs)
        _, c = simple_intersect_ocsum(psets[d], ones, okeys, ocnt, k, cs)
        across[d, 1] = histogram(c, nbins)
        _, c = simple_kmers_subtract(psets[d], ones, okeys, k)
        across[d, 0] = histogram(c, nbins)
    return within, across


def kmer_counts(fasta, k: int, cs: int = 255):
    """`kmc -fm -k{k} -ci1` (exp_type_4.smk:146-153): distinct canonical k-mers, ascending, with their occurrence
    counts saturated at KMC's default -cs255.  Returns (keys, uint32 counts)."""
    keys, _ = kmers(fasta, k)
    if keys.shape[0] == 0:
        return keys, np.empty(0, np.uint32)
    if k <= 32:
        u, c = np.unique(keys, return_counts=True)
    else:
        order = np.lexsort((keys[:, 0], keys[:, 1]))
        srt = keys[order]
        new = np.ones(srt.shape[0], dtype=bool)
        new[1:] = np.any(srt[1:] != srt[:-1], axis=1)
        u = srt[new]
        starts = np.flatnonzero(new)
        c = np.diff(np.append(starts, srt.shape[0]))
    return u, np.minimum(c, cs).astype(np.uint32)


def exp4(groups, pivots, k: int, cs: int = CS_DEFAULT):
    """Experiment type 4's databases for one k (exp_type_4.smk:136-243): per pivot its counted k-mers, and for every
    (pivot, dataset) the result of `kmc_tools simple union_d pivot_p intersect -ocsum`.
    Returns (pivot_tables [(keys, counts)], intersections [[(keys, counts)] per pivot])."""
    unions = []
    for genomes in groups:
        ukeys, _ = union_sum([genome_set(g, k) for g in genomes], k, cs)
        unions.append(ukeys)
    tables, inters = [], []
    for p in pivots:
        pk, pc = kmer_counts(p, k)
        tables.append((pk, pc))
        row = []
        for ukeys in unions:
            ones = np.ones(ukeys.shape[0], dtype=np.uint32)
            row.append(simple_intersect_ocsum(ukeys, ones, pk, pc, k, cs))
        inters.append(row)
    return tables, inters


def num_threads() -> int:
    return int(lib().ko_num_threads())


def set_num_threads(n: int) -> int:
    """Use n host threads from now on (torchrun exports OMP_NUM_THREADS=1; the reference arm asks for all cores). Returns n."""
    lib().ko_set_num_threads(int(n))
    return num_threads()


def host_cores() -> int:
    """Cores this process may run on (affinity mask if the platform has one, else os.cpu_count())."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)
