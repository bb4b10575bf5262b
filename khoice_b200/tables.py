"""step_5 / step_9 table builders and the histogram text format (host side, FP64 Python semantics).

Mirrors, name for name, the reference's helpers for this path:

* ``summarize_histogram_type1``      <- /root/reference/workflow/rules/exp_type_1.smk:115-150
* ``within_group_union_analysis``    <- rule of the same name, exp_type_1.smk:193-231  (step_5 CSV)
* ``across_group_union_analysis``    <- rule of the same name, exp_type_1.smk:262-297  (step_9 CSV)
* ``get_num_of_dataset_members``     <- exp_type_1.smk:107-113
* histogram text ``"<occ>\\t<n>\\n"`` <- what ``kmc_tools transform X histogram`` writes and what
  the rules parse with ``int(record.split()[1])`` (exp_type_1.smk:210-212, 279-281)

Experiment type 2 (next row N1 of SURVEY.md section 8f):

* ``summarize_histogram_type2``          <- /root/reference/workflow/rules/exp_type_2.smk:171-216
* ``within_group_analysis_exp_type2``    <- rule of the same name, exp_type_2.smk:398-436
* ``across_group_analysis_exp_type2``    <- rule of the same name, exp_type_2.smk:515-553
* ``get_num_of_dataset_members_exp2``    <- exp_type_2.smk:121-127

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


# ---- experiment type 2 -------------------------------------------------------------------------------------
WITHIN_HEADER_T2 = WITHIN_HEADER
ACROSS_HEADER_T2 = ("group_num,k,percent_1_occ,percent_2_to_3,percent_4_to_8,percent_9_more,"
                    "unique_stat,unique_stat_norm,delta_frac,delta_frac_norm\n")


def summarize_histogram_type2(sub_counts: Sequence[int], inter_counts: Sequence[int], num_genomes_in_dataset: int,
                              across_group_analysis: bool, k: int) -> List[float]:
    """Seven metrics of a pivot genome against a union: ``sub_counts`` is the histogram of the pivot k-mers absent
    from the union (all counters 1), ``inter_counts`` that of the shared k-mers with counter 1 + union counter
    (row i = counter i+1).  Bin edges: 25 % / 75 % of the member count within a group, indices 3 / 8 across groups."""
    if inter_counts[0] != 0:
        raise AssertionError("intersection counts should have 0 unique kmers")
    if _lsum(sub_counts[1:]) != 0:
        raise AssertionError("all of kmers in sub_counts should be unique")
    total = _lsum(sub_counts) + _lsum(inter_counts)
    if across_group_analysis:
        lo, hi = 3, 8
    else:
        lo = max(int(0.25 * num_genomes_in_dataset), 1)
        hi = max(int(0.75 * num_genomes_in_dataset), 1)
    n = len(inter_counts)

    def share(a: int, b: int) -> float:
        return round(_lsum(inter_counts[i] for i in range(a, b)) / total, 3)

    out = [round(sub_counts[0] / total, 3), share(1, lo), share(lo, hi), share(hi, n)]
    if not abs(_lsum(out) - 1) < 0.05:
        raise AssertionError("Issue occurred with histogram summarization")
    stat = 1 * sub_counts[0] / total
    stat += _lsum((i + 1) * (inter_counts[i] / total) for i in range(1, n))
    out.append(round(stat, 4))
    norm = (1 / num_genomes_in_dataset) * sub_counts[0] / total
    norm += _lsum(((i + 1) / num_genomes_in_dataset) * (inter_counts[i] / total) for i in range(1, n))
    out.append(round(norm, 4))
    out.append(round(total / k, 4))
    return out


def get_num_of_dataset_members_exp2(dataset_num, input_root: str = "input_type_2") -> int:
    """Number of ``*.fna.gz`` files in ``input_type_2/rest_of_set/dataset_{n}`` (exp_type_2.smk:121-127)."""
    d = os.path.join(input_root, "rest_of_set", f"dataset_{dataset_num}")
    return sum(1 for f in os.listdir(d) if f.endswith(".fna.gz"))


def _pivot_analysis(input_files: Sequence[str], output_csv: str, num_datasets: int, header: str, across: bool,
                    members_of: Callable[[str], int]) -> None:
    """Shared body of the two type-2 table rules.  ``input_files`` alternate subtract / intersect histograms in the
    order of get_{within,across}_group_histogram_files (dataset-major, then k; exp_type_2.smk:153-169); dataset and k
    are parsed from path components 2 and 1 like the rules do."""
    rows = []
    for i in range(0, len(input_files), 2):
        parts = input_files[i].split("/")
        dataset_num = parts[-3].split("_")[1]
        k = parts[-4].split("_")[1]
        sub = read_histogram_file(input_files[i])
        inter = read_histogram_file(input_files[i + 1])
        n_members = num_datasets if across else members_of(dataset_num)
        rows.append([f"group_{dataset_num}", k] + summarize_histogram_type2(sub, inter, n_members, across, int(k)))
    for n in range(1, num_datasets + 1):
        mine = [r for r in rows if r[0] == f"group_{n}"]
        top = max(r[8] for r in mine)
        for r in mine:
            r.append(round(r[8] / top, 4))
    _emit(output_csv, header, rows)


def within_group_analysis_exp_type2(input_files: Sequence[str], output_csv: str, num_datasets: int,
                                    members_of: Callable[[str], int] = get_num_of_dataset_members_exp2) -> None:
    """``within_dataset_analysis_type_2/within_dataset_analysis.csv`` (exp_type_2.smk:398-436)."""
    _pivot_analysis(input_files, output_csv, num_datasets, WITHIN_HEADER_T2, False, members_of)


def across_group_analysis_exp_type2(input_files: Sequence[str], output_csv: str, num_datasets: int) -> None:
    """``across_dataset_analysis_type_2/across_dataset_analysis.csv`` (exp_type_2.smk:515-553)."""
    _pivot_analysis(input_files, output_csv, num_datasets, ACROSS_HEADER_T2, True, lambda n: num_datasets)
