"""Seeded synthetic genome generator for the exp-type-1 path.

The reference has no generator -- it downloads RefSeq genomes from NCBI
(/root/reference/src/download_genomes.py:46-122) into ``database_N/dataset_M/`` -- and there is no
network here, so BASELINE.json's configs are defined on synthetic genomes.  This module produces the
inputs in the layout the rules expect: ``{work_root}/data/dataset_{n}/{genome}.fna.gz``
(/root/reference/workflow/rules/exp_type_1.smk:44-47,158).

Model: a uniform-ACGT root; per 10 kb block a sharing class decides whether a
group takes the block from the root (core), from its super-clade (25 groups), from its clade (4 groups)
or from its own private sequence, so that the across-group bins 1 / 2-5 / 6-20 / 21+ are all
populated; a genome descends from its group ancestor through one of four sub-clades (1 %
substitutions) plus 0.2-0.5 % private substitutions and 0.02 % indels, so that the within-group bins
are populated; then N-runs, rare IUPAC codes, soft-masked lower case, 1-4 records, 80-column lines.

Every genome is a pure function of (seed, group, genome): ranks and worker processes can generate
their own shard without coordination.  seed(group, genome) = base + 1000*group + genome (1-based).
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass
from functools import lru_cache

import numpy as np

BASE_SEED = 20240131
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_IUPAC = np.frombuffer(b"RYKMSWBDHVN", dtype=np.uint8)
BLOCK = 10_000


@dataclass(frozen=True)
class SynthConfig:
    n_groups: int = 2
    genomes_per_group: int = 5
    genome_len: int = 5_000_000
    seed: int = BASE_SEED
    line_width: int = 80

    @property
    def n_genomes(self) -> int:
        return self.n_groups * self.genomes_per_group


def _rng(*key) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(list(int(x) for x in key)))


def _random_seq(rng: np.random.Generator, n: int) -> np.ndarray:
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def _substitute(seq: np.ndarray, rng: np.random.Generator, rate: float) -> np.ndarray:
    """Replace ~rate of the positions by a different base (shift by 1..3 in ACGT order)."""
    n = seq.size
    m = int(round(n * rate))
    if m == 0:
        return seq
    pos = rng.integers(0, n, size=m)
    shift = rng.integers(1, 4, size=m, dtype=np.uint8)
    code = np.searchsorted(_ACGT, seq[pos]).astype(np.uint8)
    out = seq.copy()
    out[pos] = _ACGT[(code + shift) & 3]
    return out


@lru_cache(maxsize=2)
def _root(seed: int, length: int):
    rng = _rng(seed)
    seq = _random_seq(rng, length)
    nblocks = (length + BLOCK - 1) // BLOCK
    # sharing class per block: 0 core, 1 super-clade, 2 clade, 3 private
    cls = rng.choice(4, size=nblocks, p=[0.20, 0.15, 0.20, 0.45]).astype(np.uint8)
    return seq, np.repeat(cls, BLOCK)[:length]


@lru_cache(maxsize=4)
def _group_ancestor(seed: int, length: int, group: int) -> np.ndarray:
    root, cls = _root(seed, length)
    layers = (
        root,
        _random_seq(_rng(seed, 1, (group - 1) // 25), length),
        _random_seq(_rng(seed, 2, (group - 1) // 4), length),
        _random_seq(_rng(seed, 3, group), length),
    )
    anc = np.choose(cls, layers)
    return _substitute(anc, _rng(seed, 4, group), 0.001)


@lru_cache(maxsize=8)
def _subclade(seed: int, length: int, group: int, sub: int) -> np.ndarray:
    return _substitute(_group_ancestor(seed, length, group), _rng(seed, 5, group, sub), 0.01)


def genome_name(group: int, genome: int) -> str:
    return f"syn_g{group:03d}_{genome:04d}"


def make_genome(cfg: SynthConfig, group: int, genome: int) -> bytes:
    """FASTA text of genome ``genome`` (1-based) of group ``group`` (1-based)."""
    rng = _rng(cfg.seed + 1000 * group + genome)
    anc_len = int(cfg.genome_len * 1.02)
    seq = _subclade(cfg.seed, anc_len, group, (genome - 1) % 4)
    seq = _substitute(seq, rng, float(rng.uniform(0.002, 0.005)))
    # indels: 0.02 % of positions, half deletions, half single/multi-base insertions
    n_ev = int(round(seq.size * 0.0002))
    if n_ev:
        dele = rng.integers(0, seq.size, size=n_ev // 2)
        seq = np.delete(seq, dele)
        ins_at = np.sort(rng.integers(0, seq.size, size=n_ev - n_ev // 2))
        seq = np.insert(seq, ins_at, _random_seq(rng, ins_at.size))
    # length 5 Mbp +- 2 %
    seq = seq[: int(seq.size * (1.0 - float(rng.uniform(0.0, 0.04))))].copy()
    n = seq.size
    # soft-masked lower case: ~1 % in a few stretches
    n_mask = max(1, n // 500_000)
    for s in rng.integers(0, max(1, n - 5000), size=n_mask):
        seq[s:s + 5000] |= 0x20
    # N runs: ~0.01 % of positions, run length 1..500
    n_runs = max(1, int(n * 0.0001 / 250))
    for s, ln in zip(rng.integers(0, n, size=n_runs), rng.integers(1, 501, size=n_runs)):
        seq[s:s + ln] = ord("N")
    # IUPAC ambiguity codes at 1e-5
    n_amb = max(1, int(n * 1e-5))
    seq[rng.integers(0, n, size=n_amb)] = _IUPAC[rng.integers(0, _IUPAC.size, size=n_amb)]
    # 1-4 records, plus (first genome of each group) one contig shorter than any k used
    n_rec = int(rng.integers(1, 5))
    cuts = [0] + sorted(int(x) for x in rng.integers(1, max(2, n), size=n_rec - 1)) + [n]
    name = genome_name(group, genome)
    parts = []
    for i in range(n_rec):
        parts.append(f">{name}_c{i + 1} synthetic group={group} genome={genome}\n".encode())
        parts.append(_wrap(seq[cuts[i]:cuts[i + 1]], cfg.line_width))
    if genome == 1:
        parts.append(f">{name}_short\n".encode())
        parts.append(b"ACGTT\n")
    return b"".join(parts)


def _wrap(seq: np.ndarray, width: int) -> bytes:
    n = seq.size
    if n == 0:
        return b""
    full = n // width
    body = np.empty((full, width + 1), dtype=np.uint8)
    body[:, :width] = seq[: full * width].reshape(full, width)
    body[:, width] = 10
    tail = seq[full * width:]
    out = body.tobytes()
    if tail.size:
        out += tail.tobytes() + b"\n"
    return out


def count_bases(fasta: bytes) -> int:
    """Number of sequence symbols (everything outside header lines except \\n / \\r) -- the unit of
    the Gbases/s metric."""
    a = np.frombuffer(fasta, dtype=np.uint8)
    n = a.size
    is_gt = a == ord(">")
    is_nl = a == 10
    # header state = last event among {'>', '\n'} is '>'
    ev = np.where(is_gt, 2, np.where(is_nl, 1, 0)).astype(np.int8)
    idx = np.where(ev > 0, np.arange(n), -1)
    last = np.maximum.accumulate(idx)
    in_hdr = np.zeros(n, dtype=bool)
    has = last >= 0
    in_hdr[has] = ev[last[has]] == 2
    seq = ~in_hdr & ~is_nl & (a != 13)
    return int(seq.sum())


def write_dataset(cfg: SynthConfig, work_root: str, compresslevel: int = 1) -> None:
    """Materialise ``{work_root}/data/dataset_{n}/{genome}.fna.gz`` for every group and genome."""
    for g in range(1, cfg.n_groups + 1):
        d = os.path.join(work_root, "data", f"dataset_{g}")
        os.makedirs(d, exist_ok=True)
        for i in range(1, cfg.genomes_per_group + 1):
            with gzip.open(os.path.join(d, genome_name(g, i) + ".fna.gz"), "wb", compresslevel=compresslevel) as fd:
                fd.write(make_genome(cfg, g, i))


def write_dataset_type2(cfg: SynthConfig, work_root: str, compresslevel: int = 1) -> None:
    """Experiment type 2 layout (/root/reference/workflow/rules/exp_type_2.smk:31-48): the LAST genome of every
    group is the held-out pivot ``input_type_2/pivot/dataset_{n}/pivot_{n}.fna.gz``, the others go to
    ``input_type_2/rest_of_set/dataset_{n}/`` (next to a ``nonpivot_names.txt`` like the reference's database)."""
    for g in range(1, cfg.n_groups + 1):
        d = os.path.join(work_root, "input_type_2", "rest_of_set", f"dataset_{g}")
        os.makedirs(d, exist_ok=True)
        names = []
        for i in range(1, cfg.genomes_per_group):
            names.append(genome_name(g, i))
            with gzip.open(os.path.join(d, genome_name(g, i) + ".fna.gz"), "wb", compresslevel=compresslevel) as fd:
                fd.write(make_genome(cfg, g, i))
        with open(os.path.join(d, "nonpivot_names.txt"), "w") as fd:
            fd.write("\n".join(names) + "\n")
        p = os.path.join(work_root, "input_type_2", "pivot", f"dataset_{g}")
        os.makedirs(p, exist_ok=True)
        with gzip.open(os.path.join(p, f"pivot_{g}.fna.gz"), "wb", compresslevel=compresslevel) as fd:
            fd.write(make_genome(cfg, g, cfg.genomes_per_group))


def write_dataset_type4(cfg: SynthConfig, work_root: str, compresslevel: int = 1, out_pivot: bool = True) -> None:
    """Experiment type 4 layout (exp_type_4.smk:31-52): like type 2 but under ``input_type4/`` with all pivots in one
    directory; with out_pivot=False the pivot is also a member of its rest of set."""
    os.makedirs(os.path.join(work_root, "input_type4", "pivot"), exist_ok=True)
    for g in range(1, cfg.n_groups + 1):
        d = os.path.join(work_root, "input_type4", "rest_of_set", f"dataset_{g}")
        os.makedirs(d, exist_ok=True)
        for i in range(1, cfg.genomes_per_group):
            with gzip.open(os.path.join(d, genome_name(g, i) + ".fna.gz"), "wb", compresslevel=compresslevel) as fd:
                fd.write(make_genome(cfg, g, i))
        pivot = make_genome(cfg, g, cfg.genomes_per_group)
        with gzip.open(os.path.join(work_root, "input_type4", "pivot", f"pivot_{g}.fna.gz"), "wb", compresslevel=compresslevel) as fd:
            fd.write(pivot)
        if not out_pivot:
            with gzip.open(os.path.join(d, f"pivot_{g}.fna.gz"), "wb", compresslevel=compresslevel) as fd:
                fd.write(pivot)


# ---- experiment type 6: simulated reads of the pivot genomes -------------------------------------------------------
def _clean_sequence(fasta: bytes) -> np.ndarray:
    """Upper-case ACGT symbols of a FASTA text, records joined (what a read simulator samples from); every other
    symbol (N runs, IUPAC codes) is dropped."""
    a = np.frombuffer(fasta, dtype=np.uint8)
    n = a.size
    is_gt, is_nl = a == ord(">"), a == 10
    ev = np.where(is_gt, 2, np.where(is_nl, 1, 0)).astype(np.int8)
    idx = np.where(ev > 0, np.arange(n), -1)
    last = np.maximum.accumulate(idx)
    in_hdr = np.zeros(n, dtype=bool)
    has = last >= 0
    in_hdr[has] = ev[last[has]] == 2
    seq = (a[~in_hdr & ~is_nl] & 0xDF).astype(np.uint8)
    return seq[np.isin(seq, np.frombuffer(b"ACGT", np.uint8))]


_COMP = np.zeros(256, dtype=np.uint8)
_COMP[np.frombuffer(b"ACGT", np.uint8)] = np.frombuffer(b"TGCA", np.uint8)


def make_reads(cfg: SynthConfig, group: int, read_type: str, n_reads: int = 200) -> bytes:
    """Stand-in for the reference's read simulators (ART / PBSIM2 in prepare_data.smk; not available here): reads of the
    pivot genome of ``group`` -- the last genome of the group -- as the single-line FASTA the reference's
    ``pivot_{n}_subset.fa`` files are (src/merge_lists.py:149-156 treats every line without '>' as one read).
    illumina: 150 bp, 0.5 % substitutions; ont: 600-3000 bp, 6 % substitutions; both strands; upper-case ACGT only
    (the reference's read-level code raises KeyError on anything else)."""
    if read_type not in ("illumina", "ont"):
        raise ValueError(read_type)
    rng = _rng(cfg.seed + 7_000_000 + 1000 * group + (0 if read_type == "illumina" else 1))
    seq = _clean_sequence(make_genome(cfg, group, cfg.genomes_per_group))
    out = []
    for i in range(n_reads):
        ln = 150 if read_type == "illumina" else int(rng.integers(600, 3001))
        ln = min(ln, seq.size)
        s = int(rng.integers(0, seq.size - ln + 1))
        r = seq[s:s + ln].copy()
        r = _substitute(r, rng, 0.005 if read_type == "illumina" else 0.06)
        if rng.random() < 0.5:
            r = _COMP[r[::-1]]
        out.append(f">pivot_{group}_{read_type}_{i + 1}\n".encode() + r.tobytes() + b"\n")
    return b"".join(out)


def write_dataset_type6(cfg: SynthConfig, work_root: str, n_reads: int = 200, compresslevel: int = 1) -> None:
    """Experiment type 6 layout (exp_type_6.smk:31-57): ``exp6_input/rest_of_set/dataset_{n}/*.fna.gz``,
    ``exp6_input/pivot/pivot_{n}.fna.gz`` and the two read files ``exp6_input/pivot_reads_subset/{illumina,ont}/pivot_{n}.fa``."""
    os.makedirs(os.path.join(work_root, "exp6_input", "pivot"), exist_ok=True)
    for rt in ("illumina", "ont"):
        os.makedirs(os.path.join(work_root, "exp6_input", "pivot_reads_subset", rt), exist_ok=True)
    for g in range(1, cfg.n_groups + 1):
        d = os.path.join(work_root, "exp6_input", "rest_of_set", f"dataset_{g}")
        os.makedirs(d, exist_ok=True)
        for i in range(1, cfg.genomes_per_group):
            with gzip.open(os.path.join(d, genome_name(g, i) + ".fna.gz"), "wb", compresslevel=compresslevel) as fd:
                fd.write(make_genome(cfg, g, i))
        with gzip.open(os.path.join(work_root, "exp6_input", "pivot", f"pivot_{g}.fna.gz"), "wb", compresslevel=compresslevel) as fd:
            fd.write(make_genome(cfg, g, cfg.genomes_per_group))
        for rt in ("illumina", "ont"):
            with open(os.path.join(work_root, "exp6_input", "pivot_reads_subset", rt, f"pivot_{g}.fa"), "wb") as fd:
                fd.write(make_reads(cfg, g, rt, n_reads))
