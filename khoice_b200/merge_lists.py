"""Confusion matrix of experiment type 4 (feature level) -- the host-side table builder that stands in for the
reference's /root/reference/src/merge_lists.py (same command line, same three output files, same bytes).

The reference joins text dumps in a Python dictionary (merge_lists.py:14-33): for pivot p, every k-mer x with
occurrence count c(x) and the list M(x) of datasets whose rest-of-set union contains it,

    matrix[p][d]           += 1 / len(M(x)) * c(x)      for d in M(x)            (merge_lists.py:132-138)
    matrix[p][d]           += 1 / G * sum of c(x) over k-mers with M(x) empty    (merge_lists.py:140-144)
    matrix_with_ucol[p][d]    = the first sum alone                              (merge_lists.py:139, 147)

in the order of the pivot dump (ascending k-mers, `dump -s`).  Floating-point addition is not associative, so the
ORDER is part of the result: here every column is accumulated with ``numpy.cumsum`` over the k-mers in dump order,
which performs exactly the reference's sequence of additions (adding 0.0 for k-mers outside the column changes
nothing), and columns that never receive a term stay the integer 0 like the reference's (they print as ``0``).

Two front ends:
* ``confusion_from_masks`` -- counts + group-membership bitmasks straight from the GPU (Engine.group_membership);
* ``main`` / ``confusion_from_dumps`` -- the reference's command line over text dumps (rule-compatible mode).
Read level (``-r DIR``, merge_lists.py:149-181; what experiment type 6's rule passes, exp_type_6.smk:327-346): every line
without ``>`` of ``DIR/pivot_{p}.fa`` is a read; its k-mers vote ``1 / len(M(x))`` for every dataset of M(x), in read order,
and the read counts 1 for the dataset with the most votes -- ties (also: a read without any hit) are drawn with
``random.choice`` from Python's global generator, UNSEEDED in the reference, so tied reads make the reference's own output
vary from run to run.  This module makes the same calls in the same order (one ``random.choice`` per read, over the same
candidate array), so that under one ``random.seed`` both programs print the same bytes; the votes come either from the GPU
(``read_level_row`` over Engine.read_votes) or, in rule-compatible mode, from the text dumps (``votes_from_dump_index``).
Like the reference, a read that is long enough to have k-mers and contains anything but upper-case ACGT is an error.
"""
from __future__ import annotations

import argparse
import os
import random
import sys
from typing import List, Sequence

import numpy as np


def _column_sums(weights: np.ndarray, member: np.ndarray) -> list:
    """Per dataset column: sequential (left to right) sum of the weights of the k-mers that belong to it; the integer
    0 if there is none.  weights float64 [n]; member bool [n, G]."""
    out = []
    for d in range(member.shape[1]):
        col = member[:, d]
        if not col.any():
            out.append(0)
        else:
            out.append(float(np.cumsum(np.where(col, weights, 0.0))[-1]))
    return out


def confusion_rows(counts: np.ndarray, member: np.ndarray, num_datasets: int):
    """One pivot's rows of (confusion_matrix, confusion_matrix_with_unidentified).  counts int [n] in dump order,
    member bool [n, num_datasets]."""
    counts = np.asarray(counts, dtype=np.int64)
    n_matches = member.sum(axis=1)
    hit = n_matches > 0
    weights = np.zeros(counts.shape[0], dtype=np.float64)
    weights[hit] = (1.0 / n_matches[hit]) * counts[hit]          # 1 / len(matches) * count
    with_ucol = _column_sums(weights, member) + [0]
    unique_pivot_count = int(counts[~hit].sum())
    spread = 1 / num_datasets * unique_pivot_count
    regular = [v + spread for v in with_ucol[:num_datasets]] + [0]
    return regular, with_ucol


def confusion_from_masks(pivot_counts: Sequence[np.ndarray], pivot_masks: Sequence[np.ndarray], num_datasets: int):
    """pivot_counts[p]: uint32 [n_p] (ascending k-mer order); pivot_masks[p]: uint64 [n_p, words], bit d of word d//64
    set iff dataset d+1 holds the k-mer.  Returns (matrix, matrix_with_ucol), lists of rows."""
    matrix, matrix_u = [], []
    for counts, masks in zip(pivot_counts, pivot_masks):
        masks = np.asarray(masks, dtype=np.uint64).reshape(len(counts), -1)
        member = np.zeros((len(counts), num_datasets), dtype=bool)
        for d in range(num_datasets):
            member[:, d] = (masks[:, d // 64] >> np.uint64(d % 64)) & np.uint64(1)
        r, u = confusion_rows(counts, member, num_datasets)
        matrix.append(r)
        matrix_u.append(u)
    return matrix, matrix_u


# ---- read level (experiment type 6) ---------------------------------------------------------------------
def split_reads(text: bytes) -> List[bytes]:
    """The reads of a ``pivot_{p}.fa`` file as the reference sees them (merge_lists.py:150-156): ``readlines()``, lines
    containing '>' are skipped, every other line -- stripped -- is one read (an empty line is a read without k-mers)."""
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return [ln.strip() for ln in lines if b">" not in ln]


def check_reads(reads: Sequence[bytes], k: int) -> None:
    """The reference raises KeyError (rev_comp_dict / the pivot dictionary) for a k-mer with anything but ACGT."""
    ok = frozenset(b"ACGT")
    for i, r in enumerate(reads):
        if len(r) >= k and not ok.issuperset(r):
            raise KeyError(f"read {i + 1} contains a symbol other than upper-case ACGT (the reference's read-level code fails on it)")


def read_level_row(votes: np.ndarray, num_datasets: int) -> list:
    """One pivot's confusion-matrix row from the per-read votes (float64 [n_reads, num_datasets]): argmax per read,
    ties drawn exactly like merge_lists.py:176-178 (one random.choice per read over the array of maximal indexes)."""
    row = [0] * (num_datasets + 1)
    for v in np.asarray(votes, dtype=np.float64).reshape(-1, num_datasets):
        max_indexes = np.where(v == max(v.tolist()))[0]
        row[int(random.choice(max_indexes))] += 1
    return row


def votes_from_dump_index(reads: Sequence[bytes], k: int, index: dict, member: np.ndarray, num_datasets: int) -> np.ndarray:
    """Rule-compatible mode: the reference's loop (merge_lists.py:157-174) over Python ints instead of strings.
    index: canonical k-mer value -> row of `member` (bool [n, num_datasets])."""
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    out = np.zeros((len(reads), num_datasets), dtype=np.float64)
    n_match = member.sum(axis=1)
    for r, read in enumerate(reads):
        votes = [0] * num_datasets
        for i in range(0, len(read) - k + 1):
            kmer = read[i:i + k]
            rc = kmer.translate(comp)[::-1]
            row = index[int(min(kmer, rc).translate(_TO_BASE4), 4)]          # KeyError like the reference's dictionary
            if n_match[row]:
                w = 1 / int(n_match[row])
                for d in np.flatnonzero(member[row]):
                    votes[d] += w
        out[r] = votes
    return out


def calculate_accuracy_values(confusion_matrix, num_datasets: int, k) -> List[list]:
    """[k, pivot, TP, TN, FP, FN] per pivot (merge_lists.py:35-51), accumulated in the reference's loop order."""
    out = []
    for pivot in range(num_datasets):
        tp = confusion_matrix[pivot][pivot]
        fp = fn = tn = 0
        for row in range(num_datasets):
            for column in range(num_datasets + 1):
                curr = confusion_matrix[row][column]
                if column == pivot and row != pivot:
                    fp += curr
                elif row == pivot and column != pivot:
                    fn += curr
                elif row != pivot:
                    tn += curr
        out.append([k, pivot, tp, tn, fp, fn])
    return out


def write_outputs(output_path: str, k, matrix, matrix_u, num_datasets: int) -> None:
    """The three files merge_lists.py writes (:183-207); ``output_path`` is a prefix ending in '/' like the rule's."""
    k = str(k)
    os.makedirs(output_path + "confusion_matrix", exist_ok=True)
    os.makedirs(output_path + "values", exist_ok=True)
    for name, m in (("_confusion_matrix.txt", matrix), ("_confusion_matrix_with_unidentified.txt", matrix_u)):
        with open(output_path + "confusion_matrix/k_" + k + name, "w") as fd:
            for row in m:
                fd.write(",".join(str(x) for x in row) + "\n")
    with open(output_path + "values/k_" + k + "_accuracy_values.csv", "w") as fd:
        a = calculate_accuracy_values(matrix, num_datasets, k)
        b = calculate_accuracy_values(matrix_u, num_datasets, k)
        for c1, c2 in zip(a, b):
            fd.write(",".join(str(x) for x in c1) + "," + ",".join(str(x) for x in c2[2:]) + "\n")


# ---- text dumps (rule-compatible mode) ------------------------------------------------------------------
_TO_BASE4 = bytes.maketrans(b"ACGT", b"0123")


def read_dump(path: str):
    """`kmc_tools transform X dump -s` text: "<kmer>\\t<count>" per line.  Returns (kmers as Python ints, counts)."""
    keys, counts = [], []
    with open(path, "rb") as fd:
        for line in fd:
            parts = line.split()
            if not parts:
                continue
            keys.append(int(parts[0].translate(_TO_BASE4), 4))
            counts.append(int(parts[1]))
    return keys, np.asarray(counts, dtype=np.int64)


def confusion_from_dumps(pivot_files: Sequence[str], intersect_files: Sequence[str], num_datasets: int, reads_prefix: str = None, k: int = 0):
    """reads_prefix (the -r argument, used as a string prefix like the reference does): read level."""
    matrix, matrix_u = [], []
    at = 0
    for p, pf in enumerate(pivot_files):
        keys, counts = read_dump(pf)
        index = {x: i for i, x in enumerate(keys)}
        member = np.zeros((len(keys), num_datasets), dtype=bool)
        for _ in range(num_datasets):
            d = at % num_datasets                                  # merge_lists.py:31
            for x in read_dump(intersect_files[at])[0]:
                i = index.get(x)
                if i is not None:
                    member[i, d] = True
            at += 1
        if reads_prefix is None:
            r, u = confusion_rows(counts, member, num_datasets)
        else:
            with open(f"{reads_prefix}pivot_{p + 1}.fa", "rb") as fd:
                reads = split_reads(fd.read())
            r = read_level_row(votes_from_dump_index(reads, k, index, member, num_datasets), num_datasets)
            u = list(r)
        matrix.append(r)
        matrix_u.append(u)
    return matrix, matrix_u


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="merge k-mer lists into a confusion matrix (experiment types 4 and 6)")
    ap.add_argument("-n", "--num", dest="num_datasets", required=True, type=int)
    ap.add_argument("-p", "--pivot_list", dest="pivot_filelist", required=True)
    ap.add_argument("-i", "--intersect_list", dest="intersect_list", required=True)
    ap.add_argument("-o", "--output_path", dest="output_path", required=True)
    ap.add_argument("-k", "--k_value", dest="k", required=True)
    ap.add_argument("-r", "--read-level", dest="read_level", nargs=1)
    a = ap.parse_args(argv)
    if a.num_datasets <= 0:
        print("Error: The number of datasets needs to be positive integer.")
        return 1
    lists = []
    for p in (a.pivot_filelist, a.intersect_list):
        if not os.path.isfile(p):
            print("Error: One of the provided files is not valid: " + p)
            return 1
        with open(p) as fd:
            lists.append([x.strip() for x in fd.readlines()])
    for f in lists[0] + lists[1]:
        if not os.path.isfile(f):
            print(f"Error: At least one of the file paths in the file lists is not valid ({f})")
            return 1
    try:
        matrix, matrix_u = confusion_from_dumps(lists[0], lists[1], a.num_datasets, a.read_level[0] if a.read_level is not None else None, int(a.k))
    except KeyError as e:
        print(f"Error: {e}", file=sys.stderr)
        return 1
    write_outputs(a.output_path, a.k, matrix, matrix_u, a.num_datasets)
    return 0


if __name__ == "__main__":
    sys.exit(main())
