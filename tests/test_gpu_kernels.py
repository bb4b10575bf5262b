"""Parity of every CUDA kernel against the CPU oracle / numpy, through the C ABI (ctypes).

Bit-exact: everything on this path is integer work."""
import numpy as np
import pytest

from helpers import EDGE_FASTAS, as_py, is_sentinel, random_fasta, sort_rows, symbol_stream, unpack_codes

pytestmark = pytest.mark.gpu

K_ALL = [1, 2, 3, 5, 15, 16, 17, 31, 32, 33, 40, 47, 63, 64]


def _files(rng):
    return list(EDGE_FASTAS) + [random_fasta(rng, 20000), random_fasta(rng, 33333, line=61, crlf=True),
                                random_fasta(rng, 70000, p_n=0.0005, n_records=5)]


def test_stage_and_pack_match_reference_stream(engine):
    rng = np.random.default_rng(1)
    files = _files(rng)
    st = engine.stage_fasta(files)
    assert st.nbytes % 16384 == 0
    staged = st.buf.download(np.uint8, st.nbytes).tobytes()
    for i, f in enumerate(files):
        b = int(st.begin[i])
        assert staged[b:b + len(f)] == f
        gap = staged[b + len(f):int(st.begin[i + 1])]
        assert len(gap) >= 3 and gap[:2] == b"\n>" and set(gap[2:]) == {10}
    packed = engine.pack_fasta(st)
    codes_ref, valid_ref = symbol_stream(staged)
    n = packed["n_symbols"]
    assert n == len(codes_ref)
    assert packed["n_breaks"] == symbol_stream.breaks
    cw = packed["codes"].download(np.uint64, packed["codes_words"])
    vw = packed["valid"].download(np.uint32, packed["valid_words"])
    c, v = unpack_codes(cw, vw, n)
    assert np.array_equal(v, np.array(valid_ref, np.uint8))
    assert np.array_equal(c, np.array(codes_ref, np.uint8))
    # nothing beyond the stream
    assert not cw[(n + 31) // 32:].any() and not vw[(n + 31) // 32:].any()
    # tile_base is the running symbol count at every 16 KiB boundary
    tb = packed["tile_base"]
    assert tb[0] == 0 and tb[-1] == n
    for t in range(0, len(tb) - 1, max(1, (len(tb) - 1) // 7)):
        assert tb[t] == len(symbol_stream(staged[:t * 16384])[0])


@pytest.mark.parametrize("k", K_ALL)
def test_extract_matches_oracle(engine, oracle, k):
    rng = np.random.default_rng(2)
    files = _files(rng)
    st = engine.stage_fasta(files)
    packed = engine.pack_fasta(st)
    n = packed["n_symbols"]
    keys = engine.extract_kmers(packed, k)
    w = 1 if k <= 32 else 2
    got = keys.download(np.uint64, n * w).reshape((n,) if w == 1 else (n, 2))
    got = got[~is_sentinel(got)]
    ref = [oracle.kmers(f, k)[0] for f in files]
    ref = np.concatenate(ref, axis=0)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


def _rand_keys(rng, n, k, dup=0.3):
    bits = 2 * k
    lo = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    hi = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    if bits <= 64:
        if bits < 64:
            lo &= np.uint64((1 << bits) - 1)
        keys = lo
    else:
        if bits < 128:
            hi &= np.uint64((1 << (bits - 64)) - 1)
        keys = np.stack([lo, hi], axis=1)
    ndup = int(n * dup)
    if ndup and n > 1:
        src = rng.integers(0, n, size=ndup)
        dst = rng.integers(0, n, size=ndup)
        keys[dst] = keys[src]
    return keys


@pytest.mark.parametrize("k,n", [(31, 1), (31, 100), (31, 6144), (31, 6145), (31, 200003), (7, 50000), (16, 70001),
                                 (32, 123457), (33, 4096), (33, 4097), (47, 150001), (64, 99999)])
def test_sort_single_segment(engine, k, n):
    rng = np.random.default_rng(k * 1000 + n)
    keys = _rand_keys(rng, n, k)
    w = 1 if k <= 32 else 2
    buf = engine.alloc((n + 4) * 8 * w)
    buf.upload(keys)
    res = engine.sort_keys(buf, n, k)
    got = res.download(np.uint64, n * w).reshape(keys.shape)
    assert np.array_equal(got, sort_rows(keys))


@pytest.mark.parametrize("k", [31, 63])
def test_sort_segments(engine, k):
    rng = np.random.default_rng(k)
    sizes = [0, 5, 6144, 0, 12289, 1, 70000, 3, 0]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    n = int(off[-1])
    keys = _rand_keys(rng, n, k)
    w = 1 if k <= 32 else 2
    buf = engine.alloc((n + 4) * 8 * w)
    buf.upload(keys)
    res = engine.sort_keys(buf, n, k, off)
    got = res.download(np.uint64, n * w).reshape(keys.shape)
    for s in range(len(sizes)):
        a, b = int(off[s]), int(off[s + 1])
        assert np.array_equal(got[a:b], sort_rows(keys[a:b])), f"segment {s}"


def _np_runs(sorted_keys):
    """distinct rows + run lengths of a numerically sorted array, sentinels excluded."""
    keep = ~is_sentinel(sorted_keys)
    s = sorted_keys[keep]
    if s.shape[0] == 0:
        return s, np.zeros(0, np.int64)
    if s.ndim == 1:
        head = np.concatenate([[True], s[1:] != s[:-1]])
    else:
        head = np.concatenate([[True], np.any(s[1:] != s[:-1], axis=1)])
    idx = np.flatnonzero(head)
    lens = np.diff(np.concatenate([idx, [s.shape[0]]]))
    return s[idx], lens


@pytest.mark.parametrize("k,n,nd", [(31, 1, 1), (31, 4096, 300), (31, 4097, 4097), (31, 100001, 5000), (31, 300000, 17),
                                    (31, 250000, 249000), (47, 4095, 100), (63, 120001, 9000), (5, 90000, 400)])
def test_unique_and_count_runs(engine, k, n, nd):
    rng = np.random.default_rng(n + nd)
    pool = _rand_keys(rng, nd, k, dup=0.0)
    keys = pool[rng.integers(0, nd, size=n)]
    # a sentinel tail, like a real genome segment
    nsent = int(rng.integers(0, 50))
    sent = np.full((nsent,) + keys.shape[1:], 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    keys = sort_rows(np.concatenate([keys, sent], axis=0))
    ntot = keys.shape[0]
    w = 1 if k <= 32 else 2
    buf = engine.alloc((ntot + 4) * 8 * w)
    buf.upload(keys)
    ref_keys, ref_len = _np_runs(keys)
    out, cnt = engine.unique(buf, ntot, k)
    assert cnt == ref_keys.shape[0]
    got = out.download(np.uint64, cnt * w).reshape(ref_keys.shape)
    assert np.array_equal(got, ref_keys)
    for cs in (5000, 7):
        hist, runs, ok, oc = engine.count_runs(buf, ntot, k, nbins=5000, cs=cs, want_keys=True, want_counts=True)
        assert runs == ref_keys.shape[0]
        sat = np.minimum(ref_len, cs)
        ref_hist = np.bincount(sat, minlength=5001).astype(np.uint64)
        assert np.array_equal(hist, ref_hist)
        assert np.array_equal(ok.download(np.uint64, runs * w).reshape(ref_keys.shape), ref_keys)
        assert np.array_equal(oc.download(np.uint32, runs), sat.astype(np.uint32))
    hist2, runs2, _, _ = engine.count_runs(buf, ntot, k)
    assert runs2 == runs and np.array_equal(hist2, np.bincount(np.minimum(ref_len, 5000), minlength=5001).astype(np.uint64))


def test_count_runs_long_runs_across_tiles(engine):
    # runs far longer than a 4096-key tile, and a run that starts exactly on a tile boundary
    lens = [4096, 1, 9000, 4095, 2, 12288, 6000, 1, 1, 1]
    keys = np.repeat(np.arange(10, 10 + len(lens), dtype=np.uint64) * np.uint64(1 << 40), lens)
    buf = engine.alloc((keys.size + 4) * 8)
    buf.upload(keys)
    hist, runs, ok, oc = engine.count_runs(buf, keys.size, 31, want_keys=True, want_counts=True)
    assert runs == len(lens)
    sat = np.minimum(np.array(lens), 5000)
    assert np.array_equal(hist, np.bincount(sat, minlength=5001).astype(np.uint64))
    assert np.array_equal(oc.download(np.uint32, runs), sat.astype(np.uint32))


@pytest.mark.parametrize("k,parts", [(31, 2), (31, 8), (63, 4), (31, 1)])
def test_partition_by_hash(engine, k, parts):
    rng = np.random.default_rng(5)
    n = 100003
    keys = _rand_keys(rng, n, k)
    w = 1 if k <= 32 else 2
    buf = engine.alloc((n + 4) * 8 * w)
    buf.upload(keys)
    out, off = engine.partition_by_hash(buf, n, k, parts)
    assert off[0] == 0 and off[-1] == n
    got = out.download(np.uint64, n * w).reshape(keys.shape)
    assert np.array_equal(sort_rows(got), sort_rows(keys))
    owner = {}
    for p in range(parts):
        for key in as_py(got[int(off[p]):int(off[p + 1])]):
            assert owner.setdefault(key, p) == p
    if parts > 1:
        sizes = np.diff(off.astype(np.int64))
        assert sizes.min() > 0.8 * n / parts


@pytest.mark.parametrize("k", [5, 12, 13, 21, 31, 32, 33, 47, 63, 64])
def test_mixer_is_a_bijection_that_unmix_inverts(engine, k):
    rng = np.random.default_rng(k)
    n = 50_000
    keys = _rand_keys(rng, n, k, dup=0.0)
    keys[:3] = 0
    w = 1 if k <= 32 else 2
    sent = np.full((4,) + keys.shape[1:], 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    allk = np.concatenate([keys, sent], axis=0)
    buf = engine.alloc((allk.shape[0] + 4) * 8 * w)
    buf.upload(allk)
    engine.remix_keys(buf, allk.shape[0], k, inverse=False)
    mixed = buf.download(np.uint64, allk.shape[0] * w).reshape(allk.shape)
    assert np.array_equal(mixed[-4:], sent)                       # sentinels stay sentinels
    assert not np.array_equal(mixed[:n], keys)
    if w == 1 and k < 32:
        assert int(mixed[:n].max()) < (1 << (2 * k))              # stays inside [0, 4^k)
    uniq_in = np.unique(keys, axis=0).shape[0]
    assert np.unique(mixed[:n], axis=0).shape[0] == uniq_in       # injective on the sample
    engine.remix_keys(buf, allk.shape[0], k, inverse=True)
    assert np.array_equal(buf.download(np.uint64, allk.shape[0] * w).reshape(allk.shape), allk)


@pytest.mark.parametrize("k,n,nd,slack_n", [(31, 200_000, 150_000, 200_000), (31, 300_000, 3000, 64), (21, 100_000, 99_000, 100_000),
                                            (47, 120_000, 40_000, 120_000), (63, 90_000, 500, 8), (13, 50_000, 20_000, 50_000)])
def test_prefix_sort_plus_resolve_equals_full_sort_results(engine, k, n, nd, slack_n):
    """Sorting hashed keys by a prefix only and resolving runs by comparison must give the same distinct set
    and the same multiplicity histogram as a full sort (slack_n small = a deliberately too short prefix,
    i.e. many distinct keys per prefix run)."""
    rng = np.random.default_rng(n + k)
    pool = _rand_keys(rng, nd, k, dup=0.0)
    keys = pool[rng.integers(0, nd, size=n)]
    w = 1 if k <= 32 else 2
    nsent = 37
    sent = np.full((nsent,) + keys.shape[1:], 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    allk = np.concatenate([keys[: n // 2], sent[:10], keys[n // 2:], sent[10:]], axis=0)
    ntot = allk.shape[0]
    ref_keys, ref_len = _np_runs(sort_rows(allk))
    buf = engine.alloc((ntot + 4) * 8 * w)
    buf.upload(allk)
    engine.remix_keys(buf, ntot, k, inverse=False)
    fb, npass = engine.prefix_plan(k, slack_n)
    assert 0 <= fb and 1 <= npass <= (2 * k + 7) // 8
    srt = engine.sort_key_bits(buf, ntot, k, fb, npass)
    out, cnt = engine.resolve_unique(srt, ntot, k, fb)
    assert cnt == ref_keys.shape[0]
    engine.remix_keys(out, cnt, k, inverse=True)
    assert np.array_equal(sort_rows(out.download(np.uint64, cnt * w).reshape(ref_keys.shape)), ref_keys)
    for cs in (5000, 3):
        hist, runs, ok = engine.resolve_count(srt, ntot, k, fb, cs=cs, want_keys=True)
        assert runs == ref_keys.shape[0]
        assert np.array_equal(hist, np.bincount(np.minimum(ref_len, cs), minlength=5001).astype(np.uint64))
        engine.remix_keys(ok, runs, k, inverse=True)
        assert np.array_equal(sort_rows(ok.download(np.uint64, runs * w).reshape(ref_keys.shape)), ref_keys)
