"""KMC database files (khoice_b200/kmc_format.py).  [KMC-ext]: no KMC binary exists in this image, so
these are structural checks of the layout restated from KMC's API documentation -- byte-level on a hand-worked example,
round trips for every key width, a hand-assembled KMC2 (multi-bin) file, and the CLI-level database layer."""
import os
import struct

import numpy as np
import pytest

from khoice_b200 import kmc_format as KF
from khoice_b200 import kmcdb


def _key(s: str) -> int:
    v = 0
    for ch in s:
        v = (v << 2) | "ACGT".index(ch)
    return v


def _keys_array(vals, k):
    vals = sorted(vals)
    if k <= 32:
        return np.array(vals, dtype=np.uint64)
    return np.array([[v & 0xFFFFFFFFFFFFFFFF, v >> 64] for v in vals], dtype=np.uint64).reshape(len(vals), 2)


def test_hand_worked_kmc1_bytes(tmp_path):
    # k = 9: lut_prefix_length 5 (9 - 5 = 4 bases = 1 suffix byte); counters up to 5000 need 2 bytes
    kmers = ["AAAAAACGT", "AAAAATTTT", "AAAACAAAA", "TTTTTGGCC"]
    counts = [1, 300, 2, 5000]
    pre = str(tmp_path / "db")
    KF.write_kmc1(pre, 9, _keys_array([_key(s) for s in kmers], 9), np.array(counts), counter_max=5000)
    suf = open(pre + ".kmc_suf", "rb").read()
    assert suf == (b"KMCS" + bytes([0b00011011, 1, 0]) + bytes([0b11111111, 300 & 255, 300 >> 8]) + bytes([0, 2, 0])
                   + bytes([0b10100101, 5000 & 255, 5000 >> 8]) + b"KMCS")
    raw = open(pre + ".kmc_pre", "rb").read()
    assert raw[:4] == b"KMCP" and raw[-4:] == b"KMCP"
    assert len(raw) == 4 + 8 * 4 ** 5 + 64 + 4 + 4
    assert struct.unpack_from("<I", raw, len(raw) - 8)[0] == 64          # header_offset
    assert struct.unpack_from("<I", raw, len(raw) - 12)[0] == 0          # database version: KMC1
    lut = np.frombuffer(raw, "<u8", 4 ** 5, 4)
    assert lut[0] == 0 and lut[_key("AAAAA")] == 0 and lut[_key("AAAAC")] == 2 and lut[_key("AAAAG")] == 3
    assert lut[_key("TTTTT")] == 3 and np.all(np.diff(lut.astype(np.int64)) >= 0)
    h0 = len(raw) - 8 - 64
    assert struct.unpack_from("<6IQB", raw, h0) == (9, 0, 2, 5, 1, 1_000_000_000, 4, 0)
    hdr, keys, cnt = KF.read_kmc(pre)
    assert hdr["k"] == 9 and hdr["both_strands"] and hdr["n_bins"] == 1 and hdr["version"] == 0
    assert [int(x) for x in keys] == [_key(s) for s in kmers] and cnt.tolist() == counts
    assert kmcdb.kmer_strings(keys, 9).tobytes().decode() == "".join(kmers)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 8, 13, 21, 31, 32, 33, 40, 47, 63, 64])
def test_round_trip_every_width(tmp_path, k):
    rng = np.random.default_rng(k)
    space = 4 ** k
    n = min(space, 3000)
    vals = set()
    while len(vals) < n:
        vals.update(int(rng.integers(0, 1 << 62)) * int(rng.integers(1, 1 << 62)) % space for _ in range(n - len(vals)))
    vals.update([0, space - 1])
    keys = _keys_array(vals, k)
    counts = rng.integers(1, 256, size=keys.shape[0]).astype(np.uint32)
    pre = str(tmp_path / f"k{k}")
    KF.write_kmc1(pre, k, keys, counts, counter_max=255)
    hdr, k2, c2 = KF.read_kmc(pre)
    assert hdr["k"] == k and hdr["total_kmers"] == keys.shape[0] and hdr["counter_size"] == 1
    assert (k - hdr["lut_prefix_length"]) % 4 == 0
    assert np.array_equal(k2, keys) and np.array_equal(c2, counts)
    assert os.path.getsize(pre + ".kmc_suf") == 8 + keys.shape[0] * ((k - hdr["lut_prefix_length"]) // 4 + 1)


def test_counters_saturate_and_empty_database(tmp_path):
    pre = str(tmp_path / "sat")
    KF.write_kmc1(pre, 31, np.array([5, 9], np.uint64), np.array([70000, 3]), counter_max=5000)
    _, _, c = KF.read_kmc(pre)
    assert c.tolist() == [5000, 3]
    KF.write_kmc1(pre, 31, np.empty(0, np.uint64), np.empty(0, np.uint32))
    hdr, keys, c = KF.read_kmc(pre)
    assert hdr["total_kmers"] == 0 and keys.shape == (0,) and c.shape == (0,)
    with pytest.raises(ValueError):
        KF.write_kmc1(pre, 31, np.array([9, 5], np.uint64) << np.uint64(50), np.array([1, 1]))


def test_reads_a_hand_assembled_kmc2_database(tmp_path):
    """KMC2 layout (what `kmc` writes): one prefix table per bin, records ascending inside a bin only."""
    k, p, sig = 11, 3, 5
    rng = np.random.default_rng(3)
    vals = sorted(set(int(v) for v in rng.integers(0, 4 ** k, size=500)))
    bins = [[], [], []]
    for v in vals:
        bins[v % 3].append(v)                      # any assignment of k-mers to bins is legal for a reader
    recs, luts, base = [], [], 0
    cnt = {v: (v % 200) + 1 for v in vals}
    for b in bins:
        pref = np.array([v >> (2 * (k - p)) for v in b], dtype=np.int64)
        per = np.bincount(pref, minlength=4 ** p)
        luts.append(base + np.concatenate([[0], np.cumsum(per)[:-1]]))
        base += len(b)
        for v in b:
            recs.append((v & (4 ** (k - p) - 1)).to_bytes((k - p) // 4, "big") + bytes([cnt[v]]))
    lut = np.concatenate(luts + [[len(vals)]]).astype("<u8")
    sigmap = np.zeros(4 ** sig + 1, dtype="<u4")
    header = struct.pack("<7IQB", k, 0, 1, p, sig, 1, 1_000_000_000, len(vals), 0) + b"\0" * 27 + struct.pack("<I", 0x200)
    pre = str(tmp_path / "kmc2")
    with open(pre + ".kmc_pre", "wb") as fd:
        fd.write(b"KMCP" + lut.tobytes() + sigmap.tobytes() + header + struct.pack("<I", len(header)) + b"KMCP")
    with open(pre + ".kmc_suf", "wb") as fd:
        fd.write(b"KMCS" + b"".join(recs) + b"KMCS")
    hdr, keys, counts = KF.read_kmc(pre)
    assert hdr["version"] == 0x200 and hdr["n_bins"] == 3 and hdr["signature_len"] == sig
    assert [int(x) for x in keys] == vals
    assert counts.tolist() == [cnt[v] for v in vals]


def test_rejects_foreign_and_damaged_files(tmp_path):
    pre = str(tmp_path / "x")
    kmcdb.write_db(pre, 21, np.array([1, 2], np.uint64), np.array([1, 1], np.uint32), np.zeros(5001, np.uint64), 255)
    assert not KF.is_kmc_database(pre)
    with pytest.raises(ValueError):
        KF.read_kmc(pre)
    KF.write_kmc1(pre, 21, np.array([1, 2], np.uint64), np.array([1, 1], np.uint32))
    raw = open(pre + ".kmc_suf", "rb").read()
    open(pre + ".kmc_suf", "wb").write(raw[:-5] + b"KMCS")
    with pytest.raises(ValueError):
        KF.read_kmc(pre)


def test_database_layer_reads_and_writes_kmc_layout(tmp_path, monkeypatch):
    """kmcdb.read_db takes a KMC database wherever it takes its own; KHB_DB_FORMAT=kmc1 makes write_db emit one.  Text dumps
    (`kmc_tools transform dump -s`, the format src/merge_lists.py:14-33 parses) are identical either way."""
    rng = np.random.default_rng(9)
    keys = np.unique(rng.integers(0, 4 ** 31, size=2000).astype(np.uint64))
    counts = rng.integers(1, 50, size=keys.shape[0]).astype(np.uint32)
    hist = np.bincount(counts, minlength=5001).astype(np.uint64)
    own, kmc = str(tmp_path / "own"), str(tmp_path / "kmc")
    kmcdb.write_db(own, 31, keys, counts, hist, 5000)
    monkeypatch.setenv("KHB_DB_FORMAT", "kmc1")
    kmcdb.write_db(kmc, 31, keys, counts, hist, 5000)
    monkeypatch.delenv("KHB_DB_FORMAT")
    assert KF.is_kmc_database(kmc) and not KF.is_kmc_database(own)
    a, b = kmcdb.read_db(own), kmcdb.read_db(kmc)
    assert a.k == b.k == 31 and np.array_equal(a.keys, b.keys) and np.array_equal(a.counts, b.counts)
    assert np.array_equal(a.hist[:5001], b.hist[:5001])
    kmcdb.write_text_dump(own + ".txt", a.keys, a.counts, 31)
    kmcdb.write_text_dump(kmc + ".txt", b.keys, b.counts, 31)
    assert open(own + ".txt", "rb").read() == open(kmc + ".txt", "rb").read()
