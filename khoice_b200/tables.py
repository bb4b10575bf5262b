"""step_5 / step_9 table builders and the histogram text format (host side, FP64 Python semantics).

Mirrors, name for name, the reference's helpers for this path:

* ``summarize_histogram_type1``      <- /root/reference/workflow/rules/exp_type_1.smk:115-150
* ``within_group_union_analysis``    <- rule of the same name, exp_type_1.smk:193-231  (step_5 CSV)
* ``across_group_union_analysis``    <- rule of the same name, exp_type_1.smk:262-297  (step_9 CSV)
* ``get_num_of_dataset_members``     <- exp_type_1.smk:107-113
* histogram text ``"<occ>\\t<n>\\n"`` <- what ``kmc_tools transform X histogram`` writes and what
  the rules parse with ``int(record.split()[1])`` (exp_type_1.smk:210-212, 279-281)

The reference environment is Python 3.10 (workflow/envs/khoice_exps.yaml:177) where ``sum`` over
floats is a plain left-to-right accumulation; newer interpreters compensate.  ``_lsum`` pins the
3.10 behaviour so the CSV bytes do not depend on the interpreter running this package.
"""
from __future__ import annotations

import os
from typing import Callable, Iterable, List, Sequence

WITHIN_HEADER = ("group_num,k,percent_1_occ,percent_25_or_less,percent_25_to_75,"
                 "percent_75_or_more,unique_stat,unique_stat_norm,delta_frac,delta_frac_norm\n")
ACROSS_HEADER = ("group_num,k,percent_1_occ,percent_2_to_5,percent_5_to_20,percent_20_more,"
                 "unique_stat,unique_stat_norm,delta_frac,delta_frac_norm\n")

HIST_ROWS = 5000  # rows written per histogram file: occurrences 1..5000, mirroring `-cs5000`


def _lsum(values: Iterable):
    """Left-to-right accumulation starting from int 0 (CPython 3.10 ``sum``)."""
    acc = 0
    for v in values:
        acc = acc + v
    return acc


def summarize_histogram_type1(hist_counts: Sequence[int], num_dataset_members: int,
                              across_group_analysis: bool, k: int) -> List[float]:
    """Seven metrics of one occurrence histogram (``hist_counts[i]`` = #k-mers seen in i+1 members).

    [percent_1_occ, low bin, middle bin, high bin, unique_stat, unique_stat_norm, delta_frac];
    bin edges are 25 % / 75 % of the member count (at least 1) within a group, and the fixed
    indices 5 / 20 across groups."""
    total = _lsum(hist_counts)
    if across_group_analysis:
        lo, hi = 5, 20
    else:
        lo = max(int(0.25 * num_dataset_members), 1)
        hi = max(int(0.75 * num_dataset_members), 1)
    n = len(hist_counts)

    def share(a: int, b: int) -> float:
        return round(_lsum(hist_counts[i] for i in range(a, b)) / total, 3)

    out = [round(hist_counts[0] / total, 3), share(1, lo), share(lo, hi), share(hi, n)]
    if not abs(_lsum(out) - 1) < 0.05:
        raise AssertionError("Issue occurred with histogram summarization")
    out.append(round(_lsum((i + 1) * (hist_counts[i] / total) for i in range(n)), 4))
    out.append(round(_lsum(((i + 1) / num_dataset_members) * (hist_counts[i] / total) for i in range(n)), 4))
    out.append(round(total / k, 4))
    return out


def get_num_of_dataset_members(dataset_num, data_root: str = "data") -> int:
    """Number of ``*.fna.gz`` files in ``data/dataset_{n}`` (exp_type_1.smk:107-113)."""
    return sum(1 for f in os.listdir(os.path.join(data_root, f"dataset_{dataset_num}")) if f.endswith(".fna.gz"))


def write_histogram_file(path: str, hist: Sequence[int], rows: int = HIST_ROWS) -> None:
    """``hist[c]`` for c = 1..rows as ``"<c>\\t<n>\\n"`` lines; ``hist`` is indexed by occurrence
    (index 0 unused) and is zero-extended.  Written atomically (temp file + rename)."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "w") as fd:
        fd.write("".join(f"{c}\t{int(hist[c]) if c < len(hist) else 0}\n" for c in range(1, rows + 1)))
    os.replace(tmp, path)


def read_histogram_file(path: str) -> List[int]:
    """Second column of every line; row i is occurrence i+1 (exp_type_1.smk:210-212)."""
    with open(path, "r") as fd:
        return [int(record.split()[1]) for record in fd.readlines()]


def _emit(path: str, header: str, rows: List[list]) -> None:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "w") as fd:
        fd.write(header)
        for row in rows:
            fd.write(",".join(str(x) for x in row) + "\n")
    os.replace(tmp, path)


def within_group_union_analysis(input_files: Sequence[str], output_csv: str, num_datasets: int,
                                members_of: Callable[[str], int] = get_num_of_dataset_members) -> None:
    """step_5 builder.  ``input_files`` are the ``step_4/k_{k}/dataset_{n}/dataset_{n}_k{k}_hist.txt``
    paths in the rule's expand order (k-major); k and n are parsed from the path like the rule does."""
    rows = []
    for f in input_files:
        parts = f.split("/")
        k = parts[-3][2:]
        dataset_num = parts[-2].split("_")[1]
        hist = read_histogram_file(f)
        rows.append([f"group_{dataset_num}", k] + summarize_histogram_type1(hist, members_of(dataset_num), False, int(k)))
    for n in range(1, num_datasets + 1):
        mine = [r for r in rows if r[0] == f"group_{n}"]
        top = max(r[8] for r in mine)
        for r in mine:
            r.append(round(r[8] / top, 4))
    _emit(output_csv, WITHIN_HEADER, rows)


def across_group_union_analysis(input_files: Sequence[str], output_csv: str, num_datasets: int) -> None:
    """step_9 builder.  ``input_files`` are the ``step_8/k_{k}/all_datasets_k{k}_hist.txt`` paths."""
    rows = []
    for f in input_files:
        k = f.split("/")[-2][2:]
        hist = read_histogram_file(f)
        rows.append(["full_group", k] + summarize_histogram_type1(hist, num_datasets, True, int(k)))
    top = max(r[8] for r in rows)
    for r in rows:
        r.append(round(r[8] / top, 4))
    _emit(output_csv, ACROSS_HEADER, rows)
