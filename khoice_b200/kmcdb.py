"""On-disk k-mer databases for the rule-compatible mode.

The reference's rules exchange KMC databases, pairs ``X.kmc_pre`` + ``X.kmc_suf``
(/root/reference/workflow/rules/exp_type_1.smk:160-161,170-171,179-180,238-239,247-248).  BASELINE.json's
north_star pins the rule names, their inputs and the CSV outputs, not the byte format of these
intermediates (KMC's prefix-LUT format is not documented in the reference tree), so this package keeps
the two file names but stores its own simple layout:

``X.kmc_pre``  64-byte header  + uint64[hist_rows+1] occurrence histogram of the counters
``X.kmc_suf``  n_keys k-mer words (8 or 16 bytes each, ascending) + n_keys uint32 counters

In fused mode most intermediates are never read back, so only the header (``stub`` flag 1) is written -- enough for
Snakemake's file DAG / resume semantics; reading such a stub as a k-mer set is an error.  The fused group job writes its
step_3 table SET-ONLY (``stub`` flag 2): the histogram in the header and the distinct k-mers in ``.kmc_suf`` without the
per-k-mer counters (the kernels never materialise them).  ``transform X histogram`` and ``transform X set_counts v`` -- the
only two things the reference does with a step_3 table (exp_type_1.smk:184-191, 233-241) -- work on it; anything that needs
the counters is an error.

Interoperation with the real binaries: ``read_db`` also accepts KMC's own databases (KMC1 and KMC2
layouts, khoice_b200/kmc_format.py), and with the environment variable ``KHB_DB_FORMAT=kmc1`` ``write_db`` emits the KMC1
layout, so that single rules can be handed to / taken over from ``kmc`` and ``kmc_tools``.  That layout is restated from
KMC's API documentation and has not been checked against a KMC binary (none in this image).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import kmc_format

MAGIC = b"KHB200DB"
VERSION = 1
_HDR = struct.Struct("<8sIIQIIII24x")  # magic, version, k, n_keys, key_bytes, counter_max, hist_rows, stub
assert _HDR.size == 64


@dataclass
class KmerDB:
    k: int
    keys: np.ndarray               # uint64 [n] or [n, 2] (lo, hi), ascending
    counts: np.ndarray             # uint32 [n]
    hist: np.ndarray               # uint64 [hist_rows+1]; hist[c] = #keys with counter c
    counter_max: int = 255
    stub: bool = False

    @property
    def n_keys(self) -> int:
        return int(self.keys.shape[0]) if not self.stub else int(self._n)

    _n: int = 0


def _atomic_write(path: str, chunks) -> None:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as fd:
        for c in chunks:
            fd.write(c)
    os.replace(tmp, path)


def write_db(prefix: str, k: int, keys: Optional[np.ndarray], counts: Optional[np.ndarray], hist: np.ndarray,
             counter_max: int, n_keys: Optional[int] = None) -> None:
    """Write ``prefix.kmc_pre`` / ``prefix.kmc_suf``.  ``keys is None`` writes a fused-mode stub (header + histogram);
    ``counts is None`` with keys writes a set-only table (header + histogram + k-mers, no counters)."""
    stub = keys is None
    set_only = not stub and counts is None
    if not stub and not set_only and os.environ.get("KHB_DB_FORMAT", "").lower() == "kmc1":
        kmc_format.write_kmc1(prefix, k, keys, counts, counter_max)
        return
    n = int(n_keys if n_keys is not None else (0 if stub else keys.shape[0]))
    key_bytes = 8 if k <= 32 else 16
    hist = np.ascontiguousarray(hist, dtype=np.uint64)
    hdr = _HDR.pack(MAGIC, VERSION, k, n, key_bytes, counter_max, hist.size - 1, 1 if stub else 2 if set_only else 0)
    if stub:
        _atomic_write(prefix + ".kmc_suf", [b""])
    elif set_only:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        assert keys.shape[0] == n
        _atomic_write(prefix + ".kmc_suf", [keys.tobytes()])
    else:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        assert counts.shape[0] == keys.shape[0] == n
        _atomic_write(prefix + ".kmc_suf", [keys.tobytes(), counts.tobytes()])
    _atomic_write(prefix + ".kmc_pre", [hdr, hist.tobytes()])


def read_db(prefix: str, header_only: bool = False, keys_only: bool = False) -> KmerDB:
    """``header_only``: k, histogram and key count (what `transform X histogram` needs).  ``keys_only``: the caller ignores the
    counters (`transform X set_counts v`), so a set-only table is acceptable; its ``counts`` come back as None."""
    if kmc_format.is_kmc_database(prefix):
        hdr, keys, counts = kmc_format.read_kmc(prefix)
        cmax = min((1 << (8 * hdr["counter_size"])) - 1, 0xFFFFFFFF) if hdr["counter_size"] else 1
        rows = max(int(counts.max()) if counts.size else 0, 5000)
        hist = np.bincount(counts.astype(np.int64), minlength=rows + 1).astype(np.uint64)
        return KmerDB(hdr["k"], keys, counts, hist, cmax, False)
    with open(prefix + ".kmc_pre", "rb") as fd:
        raw = fd.read()
    if len(raw) < 64 or raw[:8] != MAGIC:
        raise ValueError(f"{prefix}.kmc_pre is not a khoice-b200 database (a real KMC database? see INTEGRATION.md)")
    _, version, k, n, key_bytes, cmax, rows, stub = _HDR.unpack(raw[:64])
    if version != VERSION:
        raise ValueError(f"{prefix}.kmc_pre: unsupported version {version}")
    hist = np.frombuffer(raw, dtype=np.uint64, count=rows + 1, offset=64).copy()
    shape = (n,) if key_bytes == 8 else (n, 2)
    if stub == 1 and not header_only:
        # a fused-mode placeholder (header + histogram, no k-mers): reading it as a k-mer set would silently give an EMPTY set
        # and wrong unions downstream (e.g. a rule-compatible `kmc_tools complex` re-run over step_2 stubs)
        raise ValueError(f"{prefix}.kmc_pre is a header-only stub written by the fused mode (it holds a histogram, no k-mers): rebuild this "
                         "input in rule-compatible mode (KHB_MODE=rules), or run the fused mode with stubs=False, before using it as a k-mer set")
    if stub == 2 and not header_only:
        if not keys_only:
            raise ValueError(f"{prefix}.kmc_pre is a set-only table written by the fused group job (k-mers and histogram, no per-k-mer counters): "
                             "only `transform X histogram` and `transform X set_counts v` can use it; rebuild it in rule-compatible mode "
                             "(KHB_MODE=rules) for anything that needs the counters")
        with open(prefix + ".kmc_suf", "rb") as fd:
            body = fd.read()
        if len(body) != n * key_bytes:
            raise ValueError(f"{prefix}.kmc_suf: size {len(body)} does not match header (n={n}, set-only)")
        return KmerDB(k, np.frombuffer(body, dtype=np.uint64).reshape(shape).copy(), None, hist, cmax, False)
    if stub or header_only:
        db = KmerDB(k, np.empty((0,) + shape[1:], np.uint64), np.empty(0, np.uint32), hist, cmax, True)
        db._n = n
        return db
    with open(prefix + ".kmc_suf", "rb") as fd:
        body = fd.read()
    if len(body) != n * key_bytes + n * 4:
        raise ValueError(f"{prefix}.kmc_suf: size {len(body)} does not match header (n={n})")
    keys = np.frombuffer(body, dtype=np.uint64, count=n * (key_bytes // 8)).reshape(shape).copy()
    counts = np.frombuffer(body, dtype=np.uint32, count=n, offset=n * key_bytes).copy()
    return KmerDB(k, keys, counts, hist, cmax, False)


def kmer_strings(keys: np.ndarray, k: int) -> np.ndarray:
    """k-mer words -> uint8 [n, k] ASCII letters (first base = most significant 2 bits, A C G T = 0 1 2 3)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    n = keys.shape[0]
    out = np.empty((n, k), dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for j in range(k):
        bit = 2 * (k - 1 - j)
        word = keys if keys.ndim == 1 else (keys[:, 1] if bit >= 64 else keys[:, 0])
        out[:, j] = lut[((word >> np.uint64(bit % 64)) & np.uint64(3)).astype(np.intp)]
    return out


def write_text_dump(path: str, keys: np.ndarray, counts: np.ndarray, k: int) -> None:
    """`kmc_tools transform X dump -s out.txt`: one "<kmer>\t<count>" line per k-mer, ascending (the format
    /root/reference/src/merge_lists.py:14-33 parses)."""
    letters = kmer_strings(keys, k)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as fd:
        fd.write(b"".join(letters[i].tobytes() + b"\t%d\n" % int(counts[i]) for i in range(letters.shape[0])))
    os.replace(tmp, path)
