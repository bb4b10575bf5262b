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
The read-level mode of the reference (``-r``, merge_lists.py:149-181) draws ties with ``random.choice`` and needs
simulated reads; it is not part of experiment type 4's rule (exp_type_4.smk:284-290) and is not provided.
"""
from __future__ import annotations

import argparse
import os
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


def confusion_from_dumps(pivot_files: Sequence[str], intersect_files: Sequence[str], num_datasets: int):
    matrix, matrix_u = [], []
    at = 0
    for pf in pivot_files:
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
        r, u = confusion_rows(counts, member, num_datasets)
        matrix.append(r)
        matrix_u.append(u)
    return matrix, matrix_u


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="merge k-mer lists into a confusion matrix (experiment type 4, feature level)")
    ap.add_argument("-n", "--num", dest="num_datasets", required=True, type=int)
    ap.add_argument("-p", "--pivot_list", dest="pivot_filelist", required=True)
    ap.add_argument("-i", "--intersect_list", dest="intersect_list", required=True)
    ap.add_argument("-o", "--output_path", dest="output_path", required=True)
    ap.add_argument("-k", "--k_value", dest="k", required=True)
    ap.add_argument("-r", "--read-level", dest="read_level", nargs=1)
    a = ap.parse_args(argv)
    if a.read_level is not None:
        print("Error: read-level analysis is not provided by khoice-b200 (see module docstring).", file=sys.stderr)
        return 1
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
    matrix, matrix_u = confusion_from_dumps(lists[0], lists[1], a.num_datasets)
    write_outputs(a.output_path, a.k, matrix, matrix_u, a.num_datasets)
    return 0


if __name__ == "__main__":
    sys.exit(main())
