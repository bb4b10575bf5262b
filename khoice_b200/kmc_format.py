"""KMC's own database files (``X.kmc_pre`` + ``X.kmc_suf``), SURVEY.md section 8f row N2.

The reference's rules hand KMC databases from rule to rule (/root/reference/workflow/rules/exp_type_1.smk:160-161,
170-171, 179-180, 238-239, 247-248; the same in exp_type_2/4/6).  This module writes the KMC1 layout (what ``kmc_tools``
itself emits and every KMC reader accepts) and reads both the KMC1 layout and the KMC2 layout that ``kmc`` emits, so
that a rule executed by the real binaries and a rule executed by this package can feed each other.

**[KMC-ext], not checked against a KMC binary.**  The layout is not described anywhere under /root/reference; it is
restated from KMC 3.2.1's published API documentation / ``kmc_api/kmc_file.cpp`` (the version pinned in
/root/reference/workflow/envs/khoice_exps.yaml:97).  No ``kmc`` / ``kmc_tools`` / ``kmc_dump`` exists in this image, so
the only checks are structural (tests/test_kmc_format.py: round trips, marker / header positions, a hand-assembled
KMC2 file).  The first thing to do when a KMC binary is available: ``kmc_tools transform <ours> dump -s`` and
``kmc_dump`` on files written here, and ``read_kmc`` on files written by ``kmc``.

Layout (all integers little-endian):

``X.kmc_suf``   ``"KMCS"``, records, ``"KMCS"``; record = suffix of the k-mer, ``(k - p) / 4`` bytes, most significant byte
                first (2 bits per base, A C G T = 0 1 2 3, first base most significant) + counter, ``counter_size`` bytes.
``X.kmc_pre``   ``"KMCP"``; prefix table, uint64 each; [KMC2 only: signature map, uint32[4^signature_len + 1]]; header;
                uint32 header_offset (bytes from the start of the header to this field); ``"KMCP"``.
                KMC1 (version 0): ONE table of 4^p entries, entry i = number of k-mers whose first p bases are < i
                (records are ascending by k-mer).  Header, 64 bytes: kmer_length, mode, counter_size, lut_prefix_length
                (= p), min_count, max_count (uint32 each), total_kmers (uint64), one byte "NOT both strands" (0 = canonical
                k-mers), zero padding; the last four bytes of the header are the database version (0).
                KMC2 (version 0x200): one table of 4^p entries PER BIN, concatenated, followed by one guard entry
                (= total_kmers); entries are global record numbers, records are ascending inside a bin only.  Header:
                kmer_length, mode, counter_size, lut_prefix_length, signature_len, min_count, max_count (uint32 each),
                total_kmers (uint64), the strand byte, padding, version 0x200.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Optional, Tuple

import numpy as np

MARK_PRE = b"KMCP"
MARK_SUF = b"KMCS"
KMC1_HEADER_BYTES = 64
_LUT_ORDER = (7, 6, 5, 8, 4, 3, 2, 9, 10, 11, 12, 1, 13, 14, 15)


def lut_prefix_len(k: int) -> int:
    """p with (k - p) divisible by 4 (whole suffix bytes), 7 preferred as in kmc_tools' writer."""
    for p in _LUT_ORDER:
        if p <= k and (k - p) % 4 == 0:
            return p
    raise ValueError(f"no LUT prefix length for k={k}")


def counter_bytes(counter_max: int) -> int:
    n = 1
    while counter_max >= (1 << (8 * n)) and n < 4:
        n += 1
    return n


def _as_hi_lo(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    if keys.ndim == 1:
        return np.zeros(keys.shape[0], np.uint64), keys
    return keys[:, 1].copy(), keys[:, 0].copy()


def _shift_right(hi: np.ndarray, lo: np.ndarray, s: int) -> np.ndarray:
    """low 64 bits of (hi:lo) >> s"""
    if s == 0:
        return lo
    if s >= 128:
        return np.zeros_like(lo)
    if s >= 64:
        return hi >> np.uint64(s - 64)
    return (hi << np.uint64(64 - s)) | (lo >> np.uint64(s))


def write_kmc1(prefix: str, k: int, keys: np.ndarray, counts: np.ndarray, counter_max: int = 255, min_count: int = 1,
               both_strands: bool = True, cutoff_max: int = 1_000_000_000) -> None:
    """Write ``prefix.kmc_pre`` / ``prefix.kmc_suf`` in the KMC1 layout.  keys: uint64[n] (k <= 32) or uint64[n, 2] (lo, hi),
    ascending; counts: per-k-mer counters, saturated at counter_max (KMC's -cs; it fixes the counter width).  The header's
    min_count / max_count are KMC's -ci / -cx cut-offs (-cx defaults to 1e9)."""
    if not 1 <= k <= 64:
        raise ValueError(f"k={k} outside 1..64")
    hi, lo = _as_hi_lo(keys)
    n = lo.shape[0]
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    if counts.shape[0] != n:
        raise ValueError("keys and counts differ in length")
    p = lut_prefix_len(k)
    s_bits = 2 * (k - p)
    s_bytes = (k - p) // 4
    cs = counter_bytes(counter_max)
    pref = _shift_right(hi, lo, s_bits).astype(np.int64)
    if n and (np.any(np.diff(pref) < 0)):
        raise ValueError("keys are not ascending")
    lut = np.zeros(1 << (2 * p), dtype=np.uint64)
    if n:
        per = np.bincount(pref, minlength=1 << (2 * p)).astype(np.uint64)
        lut[1:] = np.cumsum(per)[:-1]
    rec = np.zeros((n, s_bytes + cs), dtype=np.uint8)
    be = np.empty((n, 16), dtype=np.uint8)           # the whole key, most significant byte first
    be[:, :8] = hi.astype(">u8").view(np.uint8).reshape(n, 8)
    be[:, 8:] = lo.astype(">u8").view(np.uint8).reshape(n, 8)
    if s_bytes:
        rec[:, :s_bytes] = be[:, 16 - s_bytes:]
    capped = np.minimum(counts, np.uint64(counter_max))
    rec[:, s_bytes:] = capped.astype("<u8").view(np.uint8).reshape(n, 8)[:, :cs]
    header = struct.pack("<6IQB", k, 0, cs, p, min_count, cutoff_max & 0xFFFFFFFF, n, 0 if both_strands else 1)
    header += b"\0" * (KMC1_HEADER_BYTES - len(header))  # incl. the version word (0) in the last four bytes
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    for path, chunks in ((prefix + ".kmc_suf", (MARK_SUF, rec.tobytes(), MARK_SUF)),
                         (prefix + ".kmc_pre", (MARK_PRE, lut.tobytes(), header, struct.pack("<I", KMC1_HEADER_BYTES), MARK_PRE))):
        tmp = f"{path}.tmp.{os.getpid()}"
        with open(tmp, "wb") as fd:
            for c in chunks:
                fd.write(c)
        os.replace(tmp, path)


def is_kmc_database(prefix: str) -> bool:
    try:
        with open(prefix + ".kmc_pre", "rb") as fd:
            return fd.read(4) == MARK_PRE
    except OSError:
        return False


def read_header(prefix: str) -> Dict[str, int]:
    with open(prefix + ".kmc_pre", "rb") as fd:
        raw = fd.read()
    return _parse_pre(raw, prefix)[0]


def _parse_pre(raw: bytes, prefix: str):
    if len(raw) < 4 + 4 + 4 + 40 or raw[:4] != MARK_PRE or raw[-4:] != MARK_PRE:
        raise ValueError(f"{prefix}.kmc_pre: missing KMCP markers")
    version, header_offset = struct.unpack_from("<II", raw, len(raw) - 12)
    if version not in (0, 0x200):
        raise ValueError(f"{prefix}.kmc_pre: database version {version:#x} is not supported (0 and 0x200 are)")
    h0 = len(raw) - 8 - header_offset
    if h0 < 4:
        raise ValueError(f"{prefix}.kmc_pre: header offset {header_offset} outside the file")
    if version == 0:
        k, mode, cs, p, cmin, cmax_lo, total, strand = struct.unpack_from("<6IQB", raw, h0)
        cmax_hi = struct.unpack_from("<I", raw, h0 + 36)[0]
        sig = 0
        table_end = h0
    else:
        k, mode, cs, p, sig, cmin, cmax_lo, total, strand = struct.unpack_from("<7IQB", raw, h0)
        cmax_hi = 0
        table_end = h0 - 4 * ((1 << (2 * sig)) + 1)
    if mode != 0:
        raise ValueError(f"{prefix}.kmc_pre: mode {mode} (quality-aware counters) is not supported")
    if not (1 <= k <= 256 and p <= k and (k - p) % 4 == 0 and 0 <= cs <= 8):
        raise ValueError(f"{prefix}.kmc_pre: implausible header (k={k}, lut_prefix_length={p}, counter_size={cs})")
    if k > 64:
        raise ValueError(f"{prefix}.kmc_pre: k={k} beyond this package's 64")
    lut_bytes = table_end - 4
    if lut_bytes < 0 or lut_bytes % 8:
        raise ValueError(f"{prefix}.kmc_pre: prefix table of {lut_bytes} bytes")
    lut = np.frombuffer(raw, dtype="<u8", count=lut_bytes // 8, offset=4).astype(np.uint64)
    single = 1 << (2 * p)
    if version == 0:
        if lut.shape[0] != single:
            raise ValueError(f"{prefix}.kmc_pre: {lut.shape[0]} prefix entries, expected {single}")
    else:
        if lut.shape[0] % single == 1:
            lut = lut[:-1]                                      # the guard entry
        if lut.shape[0] == 0 or lut.shape[0] % single:
            raise ValueError(f"{prefix}.kmc_pre: {lut.shape[0]} prefix entries are not a multiple of {single}")
    hdr = {"version": version, "k": k, "mode": mode, "counter_size": cs, "lut_prefix_length": p, "signature_len": sig,
           "min_count": cmin, "max_count": (cmax_hi << 32) | cmax_lo, "total_kmers": total, "both_strands": strand == 0,
           "n_bins": lut.shape[0] // single}
    return hdr, lut


def read_kmc(prefix: str) -> Tuple[Dict[str, int], np.ndarray, np.ndarray]:
    """(header, keys ascending -- uint64[n] or uint64[n, 2] (lo, hi) for k > 32 --, counts uint32[n]) of a KMC1 or KMC2 database."""
    with open(prefix + ".kmc_pre", "rb") as fd:
        hdr, lut = _parse_pre(fd.read(), prefix)
    with open(prefix + ".kmc_suf", "rb") as fd:
        body = fd.read()
    if len(body) < 8 or body[:4] != MARK_SUF or body[-4:] != MARK_SUF:
        raise ValueError(f"{prefix}.kmc_suf: missing KMCS markers")
    k, p, cs, n = hdr["k"], hdr["lut_prefix_length"], hdr["counter_size"], hdr["total_kmers"]
    s_bytes = (k - p) // 4
    rec_bytes = s_bytes + cs
    if len(body) - 8 != n * rec_bytes:
        raise ValueError(f"{prefix}.kmc_suf: {len(body) - 8} record bytes, header says {n} x {rec_bytes}")
    rec = np.frombuffer(body, dtype=np.uint8, count=n * rec_bytes, offset=4).reshape(n, rec_bytes) if rec_bytes else np.zeros((n, 0), np.uint8)
    be = np.zeros((n, 16), dtype=np.uint8)
    if s_bytes:
        be[:, 16 - s_bytes:] = rec[:, :s_bytes]
    hi = be[:, :8].copy().view(">u8").reshape(n).astype(np.uint64)
    lo = be[:, 8:].copy().view(">u8").reshape(n).astype(np.uint64)
    cnt8 = np.zeros((n, 8), dtype=np.uint8)
    if cs:
        cnt8[:, :cs] = rec[:, s_bytes:]
        counts = cnt8.view("<u8").reshape(n)
    else:
        counts = np.ones(n, dtype=np.uint64)                 # a database of bare k-mers
    # prefix of record i: position of i in the (flattened) prefix table, modulo the table size
    single = 1 << (2 * p)
    if n:
        pos = np.searchsorted(lut, np.arange(n, dtype=np.uint64), side="right") - 1
        if pos.min() < 0:
            raise ValueError(f"{prefix}.kmc_pre: prefix table does not start at record 0")
        pref = (pos % single).astype(np.uint64)
        s_bits = 2 * (k - p)
        if s_bits >= 64:
            hi |= pref << np.uint64(s_bits - 64)
        else:
            lo |= pref << np.uint64(s_bits)
            if s_bits + 2 * p > 64:
                hi |= pref >> np.uint64(64 - s_bits)
    if k <= 32:
        keys = lo
        order = np.argsort(keys, kind="stable") if hdr["n_bins"] > 1 else None
    else:
        keys = np.stack([lo, hi], axis=1)
        order = np.lexsort((lo, hi)) if hdr["n_bins"] > 1 else None
    if order is not None:
        keys, counts = keys[order], counts[order]
    return hdr, np.ascontiguousarray(keys), np.minimum(counts, np.uint64(0xFFFFFFFF)).astype(np.uint32)
