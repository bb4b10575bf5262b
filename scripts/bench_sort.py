"""Micro-benchmark of the radix-sort passes (tuning aid): KHB_SORT_VARIANT=<v> python scripts/bench_sort.py [n_keys] [k] [nseg]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from khoice_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 31
nseg = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rng = np.random.default_rng(0)
w = 1 if k <= 32 else 2
bits = 2 * k
keys = rng.integers(0, 2**63, size=(n * w), dtype=np.uint64)
if w == 1:
    keys &= np.uint64((1 << bits) - 1) if bits < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
else:
    keys = keys.reshape(n, 2)
    keys[:, 1] &= np.uint64((1 << (bits - 64)) - 1) if bits < 128 else np.uint64(0xFFFFFFFFFFFFFFFF)
eng = Engine(0)
src = eng.alloc((n + 4) * 8 * w)
buf = eng.alloc((n + 4) * 8 * w)
src.upload(keys)
seg = np.linspace(0, n, nseg + 1).astype(np.uint64)
res = None
for it in range(4):
    eng._chk(eng.lib.khb_memcpy_h2d(eng.ctx, buf.ptr, keys.ctypes.data, 0))  # no-op, keeps API warm
    import ctypes as C
    # device-to-device restore of the unsorted input
    eng._chk(eng.lib.khb_memcpy_d2h(eng.ctx, None, None, 0))
    from ctypes import c_void_p
    libcudart = None
    # use the library's own copy: download/upload would be slow, so re-upload once per iteration from host pinned? keep simple:
    buf.upload(keys)
    if it == 1:
        eng.profile_enable(True)
    t0 = time.time()
    res = eng.sort_keys(buf, n, k, seg)
    dt = time.time() - t0
    if res is not buf:
        buf, res = res, buf  # keep ownership sane: result buffer becomes the working buffer
        out = buf
    else:
        out = buf
prof = eng.profile_read()
eng.profile_enable(False)
for name in ("radix_hist", "onesweep"):
    v = prof[name]
    if v["launches"]:
        print(f"{name:10s} launches={v['launches']:4d} avg_ms={v['ms'] / v['launches']:8.3f} alg_GB/s={v['alg_bytes'] / v['ms'] / 1e6:8.1f}")
tot_ms = (prof["radix_hist"]["ms"] + prof["onesweep"]["ms"]) / 3
print(f"variant={os.environ.get('KHB_SORT_VARIANT', 'default')} n={n} k={k} nseg={nseg}: sort {tot_ms:.2f} ms = {n / tot_ms / 1e6:.2f} Gkeys/s")
# correctness spot check
got = out.download(np.uint64, min(n, 2_000_000) * w)
if nseg == 1:
    ref = np.sort(keys) if w == 1 else None
    if w == 1:
        assert np.array_equal(got, ref[: got.size]), "sort mismatch"
        print("check ok")
