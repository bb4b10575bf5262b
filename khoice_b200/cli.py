"""Command-line shims for the rule-compatible mode: drop-in stand-ins for the two KMC binaries as the
reference's rules invoke them (/root/reference/workflow/rules/exp_type_1.smk):

    kmc -fm -m64 -k{k} -ci1 {in.fna.gz} {out_prefix} tmp/                    (:163)
    kmc_tools transform {in_prefix} set_counts 1 {out_prefix}                (:173, :241)
    kmc_tools complex {ops.txt}                                              (:182, :250)
    kmc_tools transform {in_prefix} histogram {out.txt}                      (:191, :259)

and, for experiment type 2 (/root/reference/workflow/rules/exp_type_2.smk:354-380, 470-496):

    kmc_tools simple {A} {B} intersect {out_prefix} -ocsum
    kmc_tools simple {A} {B} kmers_subtract {out_prefix}

and for experiment type 4 (exp_type_4.smk:254-271):

    kmc_tools transform {in_prefix} dump -s {out.txt}

``khoice_b200/bin/kmc`` and ``khoice_b200/bin/kmc_tools`` exec this module, so putting that directory
first on PATH makes the UNMODIFIED reference rules run on the B200 engine.  Databases use this package's
own layout (khoice_b200/kmcdb.py).  Exit status: 0 ok, 1 on any error with partial outputs removed
(Snakemake deletes the outputs of a failed job; the shims never leave a half-written file either).

Only what the reference's exp-1 call sites use is implemented; anything else is rejected loudly.
Every arithmetic step runs on the GPU (no CPU fallback); a process pays one CUDA context start, which is
why the fused mode (khoice_b200/pipeline.py) is the fast path.
"""
from __future__ import annotations

import os
import re
import sys
from typing import List, Optional

import numpy as np

from . import ingest, kmcdb
from .engine import COUNTER_MAX, Engine
from .tables import HIST_ROWS, write_histogram_file

KMC_DEFAULT_CS = 255  # KMC's default counter saturation for `kmc` (no -cs on the reference's command line)

_engine: Optional[Engine] = None


def get_engine() -> Engine:
    global _engine
    if _engine is None:
        _engine = Engine(int(os.environ.get("KHB_DEVICE", "0")))
    return _engine


def set_engine(eng: Optional[Engine]) -> None:
    """Let an in-process caller (the mini rule runner, tests) share one context."""
    global _engine
    _engine = eng


def read_fasta(path: str) -> bytes:
    """FASTA text of a .fna.gz (multi-member gzip handled, rule R9) or plain file."""
    return ingest.read_fasta(path)


class UsageError(Exception):
    pass


# ---- kmc -------------------------------------------------------------------------------------------
def kmc_count(fasta_path: str, out_prefix: str, k: int, ci: int = 1, cs: int = KMC_DEFAULT_CS) -> None:
    """Per-genome canonical k-mer database with occurrence counters (saturating at cs, keeping >= ci)."""
    eng = get_engine()
    text = read_fasta(fasta_path)
    staged = eng.stage_fasta([text])
    packed = eng.pack_fasta(staged)
    n = packed["n_symbols"]
    keys = eng.extract_kmers(packed, k)
    srt = eng.sort_keys(keys, n, k)
    hist, runs, ok, oc = eng.count_runs(srt, n, k, nbins=HIST_ROWS, cs=cs, want_keys=True, want_counts=True)
    w = 1 if k <= 32 else 2
    hk = ok.download(np.uint64, runs * w).reshape((runs,) if w == 1 else (runs, 2))
    hc = oc.download(np.uint32, runs)
    if ci > 1:
        keep = hc >= ci
        hk, hc = hk[keep], hc[keep]
        hist[:ci] = 0
    kmcdb.write_db(out_prefix, k, hk, hc, hist, cs)


def kmc_main(argv: List[str]) -> int:
    k, ci, cs, fm, pos = 25, 2, KMC_DEFAULT_CS, False, []
    for a in argv:
        if a == "-fm":
            fm = True
        elif re.fullmatch(r"-k\d+", a):
            k = int(a[2:])
        elif re.fullmatch(r"-ci\d+", a):
            ci = int(a[3:])
        elif re.fullmatch(r"-cs\d+", a):
            cs = int(a[3:])
        elif re.fullmatch(r"-m\d+", a) or re.fullmatch(r"-t\d+", a) or a in ("-v", "-hp"):
            pass  # memory / thread hints of the CPU tool: irrelevant here
        elif a.startswith("-"):
            raise UsageError(f"kmc: option {a} is not supported by the khoice-b200 shim")
        else:
            pos.append(a)
    if len(pos) != 3:
        raise UsageError("usage: kmc [options] <input.fna[.gz]> <output_prefix> <tmp_dir>")
    if not fm:
        raise UsageError("kmc: only multi-line FASTA input (-fm) is supported (the reference passes -fm)")
    if not 1 <= k <= 64:
        raise UsageError(f"kmc: -k{k} outside 1..64")
    kmc_count(pos[0], pos[1], k, ci, cs)
    return 0


# ---- kmc_tools ---------------------------------------------------------------------------------------
def transform_set_counts(in_prefix: str, value: int, out_prefix: str) -> None:
    db = kmcdb.read_db(in_prefix, keys_only=True)   # the counters are overwritten: a set-only step_3 table of the fused group job will do
    hist = np.zeros(HIST_ROWS + 1, dtype=np.uint64)
    if value <= HIST_ROWS:
        hist[value] = db.keys.shape[0]
    kmcdb.write_db(out_prefix, db.k, db.keys, np.full(db.keys.shape[0], value, np.uint32), hist, max(db.counter_max, value))


def transform_dump(in_prefix: str, out_txt: str) -> None:
    """`kmc_tools transform X dump -s out.txt` (exp_type_4.smk:254-258, 267-271): sorted text dump."""
    db = kmcdb.read_db(in_prefix)
    kmcdb.write_text_dump(out_txt, db.keys, db.counts, db.k)


def transform_histogram(in_prefix: str, out_txt: str) -> None:
    db = kmcdb.read_db(in_prefix, header_only=True)
    write_histogram_file(out_txt, db.hist, HIST_ROWS)


def parse_complex(path: str):
    """The operation file the reference generates (exp_type_1.smk:52-61, 75-84):
    INPUT:/setN = prefix ... OUTPUT:/prefix = (set1 + set2 + ...) OUTPUT_PARAMS:/-cs5000."""
    section, inputs, output, expr, cs = None, {}, None, None, KMC_DEFAULT_CS
    with open(path) as fd:
        for raw in fd:
            line = raw.strip()
            if not line:
                continue
            if line in ("INPUT:", "OUTPUT:", "OUTPUT_PARAMS:"):
                section = line[:-1]
                continue
            if section == "INPUT":
                name, _, val = line.partition("=")
                inputs[name.strip()] = val.strip()
            elif section == "OUTPUT":
                out, _, e = line.partition("=")
                output, expr = out.strip(), e.strip()
            elif section == "OUTPUT_PARAMS":
                for tok in line.split():
                    if re.fullmatch(r"-cs\d+", tok):
                        cs = int(tok[3:])
                    elif re.fullmatch(r"-ci\d+", tok) and int(tok[3:]) <= 1:
                        pass
                    else:
                        raise UsageError(f"kmc_tools complex: output parameter {tok} is not supported")
            else:
                raise UsageError(f"kmc_tools complex: line outside a section: {line}")
    if output is None or expr is None:
        raise UsageError("kmc_tools complex: no OUTPUT section")
    body = expr.strip()
    while body.startswith("(") and body.endswith(")"):
        body = body[1:-1].strip()
    names = [t.strip() for t in body.split("+")]
    if not names or any(not re.fullmatch(r"\w+", t) for t in names):
        raise UsageError(f"kmc_tools complex: only unions `(a + b + ...)` are supported, got: {expr}")
    missing = [t for t in names if t not in inputs]
    if missing:
        raise UsageError(f"kmc_tools complex: undefined input(s) {missing}")
    return [inputs[t] for t in names], output, cs


def complex_union(ops_path: str) -> None:
    """Union with counter sum, saturating at -cs (the `+` of kmc_tools complex)."""
    in_prefixes, out_prefix, cs = parse_complex(ops_path)
    eng = get_engine()
    dbs = [kmcdb.read_db(p) for p in in_prefixes]
    k = dbs[0].k
    if any(d.k != k for d in dbs):
        raise UsageError("kmc_tools complex: inputs were built with different k")
    if any(d.counts.size and (d.counts != 1).any() for d in dbs):
        raise UsageError("kmc_tools complex: the khoice-b200 shim sums SET inputs only (all counters 1, i.e. after "
                         "`transform ... set_counts 1`, as every reference call site does)")
    keys = np.concatenate([d.keys for d in dbs], axis=0)
    n = keys.shape[0]
    w = 1 if k <= 32 else 2
    buf = eng.alloc((n + 4) * 8 * w)
    buf.upload(keys)
    srt = eng.sort_keys(buf, n, k)
    hist, runs, ok, oc = eng.count_runs(srt, n, k, nbins=HIST_ROWS, cs=cs, want_keys=True, want_counts=True)
    hk = ok.download(np.uint64, runs * w).reshape((runs,) if w == 1 else (runs, 2))
    hc = oc.download(np.uint32, runs)
    kmcdb.write_db(out_prefix, k, hk, hc, hist, cs)


def simple_op(a_prefix: str, b_prefix: str, op: str, out_prefix: str, ocsum: bool) -> None:
    """`kmc_tools simple A B intersect O -ocsum` (counter = c_A + c_B) and `kmc_tools simple A B kmers_subtract O`
    (k-mers of A absent from B, A's counters).  The join runs on the GPU (khb_sorted_lookup)."""
    if op not in ("intersect", "kmers_subtract"):
        raise UsageError(f"kmc_tools simple: operation {op} is not supported by the khoice-b200 shim")
    if op == "intersect" and not ocsum:
        raise UsageError("kmc_tools simple intersect: only the -ocsum counter mode is supported (the reference passes -ocsum)")
    A, B = kmcdb.read_db(a_prefix), kmcdb.read_db(b_prefix)
    if A.k != B.k:
        raise UsageError("kmc_tools simple: inputs were built with different k")
    idx = get_engine().sorted_lookup(A.keys, B.keys, A.k)
    cmax = max(A.counter_max, B.counter_max)
    if op == "intersect":
        keep = idx >= 0
        counts = np.minimum(A.counts[keep].astype(np.uint64) + B.counts[idx[keep]].astype(np.uint64), cmax).astype(np.uint32)
    else:
        keep = idx < 0
        counts = A.counts[keep]
    hist = np.bincount(np.minimum(counts, HIST_ROWS + 1), minlength=HIST_ROWS + 2)[:HIST_ROWS + 1].astype(np.uint64)
    hist[0] = 0
    kmcdb.write_db(out_prefix, A.k, A.keys[keep], counts, hist, cmax)


def kmc_tools_main(argv: List[str]) -> int:
    args = [a for a in argv if not re.fullmatch(r"-t\d+", a) and a not in ("-v", "-hp")]
    if not args:
        raise UsageError("usage: kmc_tools <transform|complex|simple> ...")
    if args[0] == "transform" and "dump" in args and "-s" not in args:
        raise UsageError("kmc_tools transform dump: only the sorted dump (-s) is supported (the reference passes -s)")
    if args[0] == "simple":
        flags = [a for a in args[1:] if a.startswith("-")]
        pos = [a for a in args[1:] if not a.startswith("-")]
        if len(pos) != 4 or any(f != "-ocsum" for f in flags):
            raise UsageError("usage: kmc_tools simple <A> <B> <intersect|kmers_subtract> <out> [-ocsum]")
        simple_op(pos[0], pos[1], pos[2], pos[3], "-ocsum" in flags)
        return 0
    if args[0] == "complex":
        if len(args) != 2:
            raise UsageError("usage: kmc_tools complex <operations_file>")
        complex_union(args[1])
        return 0
    if args[0] == "transform":
        if len(args) == 5 and args[2] == "set_counts":
            transform_set_counts(args[1], int(args[3]), args[4])
            return 0
        if len(args) == 4 and args[2] == "histogram":
            transform_histogram(args[1], args[3])
            return 0
        if len(args) == 5 and args[2] == "dump" and args[3] == "-s":
            transform_dump(args[1], args[4])
            return 0
        raise UsageError("kmc_tools transform: only `set_counts <v> <out>`, `histogram <out.txt>` and `dump -s <out.txt>` are supported")
    raise UsageError(f"kmc_tools {args[0]}: not supported by the khoice-b200 shim (exp type 1 uses transform and complex)")


# ---- rule-granular fused commands (used by khoice_b200/workflow/exp_type_1.smk) ------------------------------
def fused_group(k: int, genome_paths: List[str], hist_out: Optional[str], set_prefix: Optional[str], table_prefix: Optional[str] = None) -> None:
    """The per-genome rules and within_group_union for one (k, group) in ONE process (the drop-in's fused
    `within_group_union`): the step_3 table, written SET-ONLY (histogram + distinct k-mers, kmcdb.py) so that the reference's
    own `transform ... histogram` and `transform ... set_counts 1` strings produce step_4 and step_6 from it.  With --hist /
    --set the step_4 histogram and the step_6 database are written here as well (then the table is a header-only stub)."""
    eng = get_engine()
    texts = ingest.read_many(genome_paths)
    eng.group_sets_reset()
    hist, st = eng.group_from_fasta(texts, k, nbins=HIST_ROWS, keep_set=True)
    keys = eng.group_sets_download()
    eng.group_sets_reset()
    keys = np.sort(keys) if keys.ndim == 1 else keys[np.lexsort((keys[:, 0], keys[:, 1]))]
    if hist_out:
        write_histogram_file(hist_out, hist, HIST_ROWS)
    if set_prefix:
        one = np.zeros(HIST_ROWS + 1, dtype=np.uint64)
        one[1] = keys.shape[0]
        kmcdb.write_db(set_prefix, k, keys, np.ones(keys.shape[0], np.uint32), one, COUNTER_MAX)
    if table_prefix:
        if set_prefix:
            kmcdb.write_db(table_prefix, k, None, None, hist, COUNTER_MAX, n_keys=keys.shape[0])
        else:
            kmcdb.write_db(table_prefix, k, keys, None, hist, COUNTER_MAX)


def fused_stub(k: int, prefix: str) -> None:
    """The drop-in's fused `build_kmc_database_on_genome` / `transform_genome_to_set`: a header-only placeholder, so that every
    target of the reference's DAG still resolves while the group job reads the genomes itself."""
    kmcdb.write_db(prefix, k, None, None, np.zeros(HIST_ROWS + 1, dtype=np.uint64), KMC_DEFAULT_CS)


def fused_across(k: int, set_prefixes: List[str], hist_out: Optional[str], table_prefix: Optional[str] = None) -> None:
    """Rules across_group_union + across_group_union_histogram for one k over the step_6 databases."""
    eng = get_engine()
    eng.group_sets_reset()
    for p in set_prefixes:
        db = kmcdb.read_db(p)
        if db.k != k:
            raise UsageError(f"{p}: built with k={db.k}, not {k}")
        eng.group_sets_append_host(db.keys, k, 1)
    hist, st = eng.across_groups(nbins=HIST_ROWS)
    eng.group_sets_reset()
    if hist_out:
        write_histogram_file(hist_out, hist, HIST_ROWS)
    if table_prefix:
        kmcdb.write_db(table_prefix, k, None, None, hist, COUNTER_MAX, n_keys=int(st["distinct"]))


def khb_main(argv: List[str]) -> int:
    import argparse
    ap = argparse.ArgumentParser(prog="khb")
    sub = ap.add_subparsers(dest="cmd", required=True)
    g = sub.add_parser("group")
    g.add_argument("--k", type=int, required=True)
    g.add_argument("--hist", default=None)
    g.add_argument("--set", default=None, dest="set_prefix")
    g.add_argument("--table", default=None)
    g.add_argument("genomes", nargs="+")
    s = sub.add_parser("stub")
    s.add_argument("--k", type=int, required=True)
    s.add_argument("prefix")
    a = sub.add_parser("across")
    a.add_argument("--k", type=int, required=True)
    a.add_argument("--hist", default=None)
    a.add_argument("--table", default=None)
    a.add_argument("sets", nargs="+")
    ns = ap.parse_args(argv)
    if ns.cmd == "group":
        if not (ns.hist or ns.set_prefix or ns.table):
            raise UsageError("khb group: nothing to write (give --table, --hist and / or --set)")
        fused_group(ns.k, ns.genomes, ns.hist, ns.set_prefix, ns.table)
    elif ns.cmd == "stub":
        fused_stub(ns.k, ns.prefix)
    else:
        fused_across(ns.k, ns.sets, ns.hist, ns.table)
    return 0


def _outputs_of(tool: str, argv: List[str]) -> List[str]:
    """Best-effort list of files a failed invocation may have started (for cleanup)."""
    outs = []
    try:
        if tool == "kmc":
            pos = [a for a in argv if not a.startswith("-")]
            outs = [pos[1] + ".kmc_pre", pos[1] + ".kmc_suf"]
        elif argv and argv[0] == "transform":
            outs = [argv[-1]] if argv[2] in ("histogram", "dump") else [argv[-1] + ".kmc_pre", argv[-1] + ".kmc_suf"]
        elif argv and argv[0] == "complex":
            _, out, _ = parse_complex(argv[1])
            outs = [out + ".kmc_pre", out + ".kmc_suf"]
        elif argv and argv[0] == "simple":
            pos = [a for a in argv[1:] if not a.startswith("-")]
            outs = [pos[3] + ".kmc_pre", pos[3] + ".kmc_suf"]
    except Exception:
        pass
    return outs


def main(argv: Optional[List[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("kmc", "kmc_tools", "khb"):
        print("usage: python -m khoice_b200.cli <kmc|kmc_tools|khb> ...", file=sys.stderr)
        return 1
    tool, rest = argv[0], argv[1:]
    try:
        if tool == "khb":
            return khb_main(rest)
        return kmc_main(rest) if tool == "kmc" else kmc_tools_main(rest)
    except Exception as e:  # exit-code discipline at the process boundary
        print(f"{tool} (khoice-b200): {e}", file=sys.stderr)
        for f in _outputs_of(tool, rest):
            for cand in (f, f"{f}.tmp.{os.getpid()}"):
                if os.path.exists(cand):
                    os.remove(cand)
        return 1


if __name__ == "__main__":
    sys.exit(main())
