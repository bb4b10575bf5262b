"""ctypes binding of libkhoice_b200.so (include/khoice_b200.h) -- the only way the Python host code
reaches the GPU.  There is no CPU fallback: if the library is missing, or no B200 is visible,
construction of :class:`Engine` raises :class:`KhbError`.

The reference reaches its engine by exec'ing the KMC command line tools from Snakemake rules
(/root/reference/workflow/rules/exp_type_1.smk:156-259); ``Engine.group_from_fasta`` /
``Engine.across_groups`` are the fused equivalents of those rule chains, the remaining methods expose
the individual kernels for the rule-compatible mode (khoice_b200/cli.py) and for the parity tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libkhoice_b200.so")
CSRC = os.path.join(_PKG, "csrc")

FASTA_TILE = 16384
COUNTER_MAX = 5000  # -cs5000, exp_type_1.smk:61,84

ERRORS = {0: "ok", -1: "invalid argument", -2: "CUDA error", -3: "out of memory", -4: "no CUDA device",
          -5: "buffer too small", -6: "call sequence error"}


class KhbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"khoice_b200 error {code} ({ERRORS.get(code, '?')}): {message}")
        self.code = code


class Stats(C.Structure):
    """Mirror of ``khb_stats``."""
    _fields_ = [("fasta_bytes", C.c_uint64), ("bases", C.c_uint64), ("windows", C.c_uint64),
                ("genome_distinct", C.c_uint64), ("distinct", C.c_uint64), ("ms_total", C.c_float),
                ("ms_h2d", C.c_float), ("ms_pack", C.c_float), ("ms_extract", C.c_float), ("ms_sort1", C.c_float),
                ("ms_unique", C.c_float), ("ms_sort2", C.c_float), ("ms_count", C.c_float),
                ("passes_genome", C.c_int), ("passes_group", C.c_int)]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


class TeamGroup(C.Structure):
    """Mirror of ``khb_team_group``: one group as the members of a team see it (the same values on every member but chunk_base)."""
    _fields_ = [("n_genomes_total", C.c_int32), ("n_chunks_total", C.c_int32), ("chunk_base", C.c_int32), ("parity", C.c_int32),
                ("n_sym_total", C.c_uint64), ("rho", C.c_double), ("region_cap", C.c_uint32), ("area_pct", C.c_uint32)]


# every symbol include/khoice_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SIGNATURES = [
    ("khb_abi_version", C.c_int, []),
    ("khb_init", C.c_int, [C.c_int, C.POINTER(_P)]),
    ("khb_destroy", C.c_int, [_P]),
    ("khb_last_error", C.c_char_p, [_P]),
    ("khb_device_info", C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    ("khb_launch_count", C.c_uint64, [_P]),
    ("khb_stream", _P, [_P]),
    ("khb_profile_enable", C.c_int, [_P, C.c_int]),
    ("khb_profile_read", C.c_int, [_P, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    ("khb_alloc", C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    ("khb_free", C.c_int, [_P, _P]),
    ("khb_alloc_host", C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    ("khb_free_host", C.c_int, [_P, _P]),
    ("khb_memcpy_h2d", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("khb_memcpy_d2h", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("khb_memcpy_d2d", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("khb_memset", C.c_int, [_P, _P, C.c_int, C.c_size_t]),
    ("khb_sync", C.c_int, [_P]),
    ("khb_staged_size", C.c_size_t, [C.c_int, _P]),
    ("khb_stage_fasta", C.c_int, [_P, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    ("khb_pack_fasta", C.c_int, [_P, _P, C.c_size_t, _P, _P, C.c_size_t, _P, _P]),
    ("khb_extract_kmers", C.c_int, [_P, _P, _P, C.c_size_t, C.c_int, _P]),
    ("khb_extract_kmers_hashed", C.c_int, [_P, _P, _P, C.c_size_t, C.c_int, _P]),
    ("khb_remix_keys", C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_int]),
    ("khb_prefix_plan", C.c_int, [C.c_int, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("khb_sort_key_bits", C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    ("khb_resolve_unique", C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_int, _P, C.POINTER(C.c_uint64)]),
    ("khb_resolve_count", C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_int, C.c_uint32, C.c_uint32, _P, _P, C.POINTER(C.c_uint64)]),
    ("khb_sort_keys", C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    ("khb_unique", C.c_int, [_P, _P, C.c_size_t, C.c_int, _P, C.POINTER(C.c_uint64)]),
    ("khb_count_runs", C.c_int, [_P, _P, C.c_size_t, C.c_int, C.c_uint32, C.c_uint32, _P, _P, _P, C.POINTER(C.c_uint64)]),
    ("khb_group_from_fasta", C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_uint32, _P, C.c_int, C.POINTER(Stats)]),
    ("khb_group_prefetch_fasta", C.c_int, [_P, C.c_int, _P, _P]),
    ("khb_group_from_staged", C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_uint32, _P, C.c_int, C.POINTER(Stats)]),
    ("khb_pack_group", C.c_int, [_P, C.c_int, _P, _P, C.POINTER(_P)]),
    ("khb_packed_info", C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("khb_packed_free", C.c_int, [_P, _P]),
    ("khb_group_from_packed", C.c_int, [_P, C.c_int, _P, C.c_uint32, _P, C.c_int, C.POINTER(Stats)]),
    ("khb_across_groups", C.c_int, [_P, C.c_uint32, _P, C.POINTER(Stats)]),
    ("khb_group_sets_info", C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    ("khb_group_sets_device", C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    ("khb_group_sets_append_device", C.c_int, [_P, C.c_int, _P, C.c_uint64, C.c_int, C.c_int]),
    ("khb_group_sets_hashed", C.c_int, [_P]),
    ("khb_group_sets_export", C.c_int, [_P, _P]),
    ("khb_group_sets_append_host", C.c_int, [_P, C.c_int, _P, C.c_uint64, C.c_int]),
    ("khb_group_sets_reset", C.c_int, [_P]),
    ("khb_pivot_group_from_packed", C.c_int, [_P, C.c_int, _P, C.c_uint32, _P, C.c_int, C.POINTER(Stats)]),
    ("khb_pivot_sets_info", C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("khb_pivot_across", C.c_int, [_P, C.c_uint32, _P, C.POINTER(Stats)]),
    ("khb_sorted_lookup", C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint64, C.c_int, _P]),
    ("khb_group_membership", C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P, _P, C.c_int]),
    ("khb_partition_by_hash", C.c_int, [_P, _P, C.c_uint64, C.c_int, C.c_int, _P, _P]),
    ("khb_read_votes", C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, C.c_uint64, _P, _P]),
    ("khb_bins_partition", C.c_int, [_P, _P, _P, C.c_uint64, _P, C.c_int, C.c_int, C.c_uint32, _P, _P]),
    ("khb_peer_alloc", C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_uint64, _P]),
    ("khb_peer_open", C.c_int, [_P, _P]),
    ("khb_peer_begin", C.c_int, [_P]),
    ("khb_peer_push", C.c_int, [_P]),
    ("khb_peer_counts", C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    ("khb_peer_import", C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int]),
    ("khb_peer_across", C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_uint32, _P, C.POINTER(Stats)]),
    ("khb_peer_unmap", C.c_int, [_P]),
    ("khb_peer_close", C.c_int, [_P]),
    ("khb_peer_region_keys", C.c_uint64, [_P]),
    ("khb_team_alloc", C.c_int, [_P, C.c_int, C.c_int, C.c_uint64, _P]),
    ("khb_team_open", C.c_int, [_P, _P]),
    ("khb_team_unmap", C.c_int, [_P]),
    ("khb_team_close", C.c_int, [_P]),
    ("khb_team_plan", C.c_int, [_P, C.c_int, C.POINTER(TeamGroup), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    ("khb_team_partition_fasta", C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.POINTER(TeamGroup), _P]),
    ("khb_team_partition_staged", C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.POINTER(TeamGroup), _P]),
    ("khb_team_partition_packed", C.c_int, [_P, C.c_int, _P, C.POINTER(TeamGroup), _P]),
    ("khb_team_count", C.c_int, [_P, C.c_int, C.POINTER(TeamGroup), C.c_uint32, _P, C.c_int, C.POINTER(Stats)]),
    ("khb_set_group_mode", C.c_int, [_P, C.c_int]),
    ("khb_hash_overflows", C.c_uint64, [_P]),
    ("khb_bins_counters", None, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("khb_across_counters", None, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]

_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libkhoice_b200.so for sm_100a with nvcc (khoice_b200/csrc/Makefile)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_PKG, "..", "include", "khoice_b200.h"))
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
        if verbose or r.returncode:
            print(r.stdout[-4000:], r.stderr[-4000:])
        if r.returncode:
            raise RuntimeError("nvcc build of libkhoice_b200.so failed")
    return LIB_PATH


def load_library():
    """dlopen the library and attach signatures.  Fails loudly when it is missing (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KhbError(-4, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, res, args in _SIGNATURES:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def key_words(k: int) -> int:
    return 1 if k <= 32 else 2


def key_shape(n: int, k: int):
    return (n,) if k <= 32 else (n, 2)


class DeviceBuffer:
    """A khb_alloc'd device buffer."""

    def __init__(self, eng: "Engine", nbytes: int):
        self.eng = eng
        self.nbytes = int(nbytes)
        p = _P()
        eng._chk(eng.lib.khb_alloc(eng.ctx, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def free(self):
        if self.ptr:
            self.eng.lib.khb_free(self.eng.ctx, self.ptr)
            self.ptr = None

    def upload(self, arr: np.ndarray, offset: int = 0):
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes
        self.eng._chk(self.eng.lib.khb_memcpy_h2d(self.eng.ctx, self.ptr + offset, arr.ctypes.data, arr.nbytes))
        self.eng.sync()

    def download(self, dtype, count: int, offset: int = 0) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        assert offset + out.nbytes <= self.nbytes
        if out.nbytes:
            self.eng._chk(self.eng.lib.khb_memcpy_d2h(self.eng.ctx, out.ctypes.data, self.ptr + offset, out.nbytes))
            self.eng.sync()
        return out

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PackedGroup:
    """Handle of a group packed by khb_pack_group (2-bit symbol stream resident in HBM)."""

    def __init__(self, eng: "Engine", handle):
        self.eng, self.handle = eng, handle

    def info(self) -> dict:
        n, b, d = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.eng._chk(self.eng.lib.khb_packed_info(self.handle, C.byref(n), C.byref(b), C.byref(d)))
        return {"n_symbols": n.value, "bases": b.value, "device_bytes": d.value}

    def free(self):
        if self.handle:
            self.eng.lib.khb_packed_free(self.eng.ctx, self.handle)
            self.handle = None

    def __del__(self):
        try:
            if getattr(self.eng, "ctx", None):
                self.free()
        except Exception:
            pass


@dataclass
class StagedFasta:
    buf: DeviceBuffer
    begin: np.ndarray  # uint64 [n_files+1], byte offsets (multiples of FASTA_TILE)

    @property
    def nbytes(self) -> int:
        return int(self.begin[-1])


class Engine:
    """One context on one B200.  All methods are synchronous unless stated otherwise."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.ctx = _P()
        rc = self.lib.khb_init(device, C.byref(self.ctx))
        if rc != 0:
            msg = self.lib.khb_last_error(None)
            self.ctx = None
            raise KhbError(rc, msg.decode() if msg else "")
        self.device = device

    # -- plumbing -------------------------------------------------------------------------------------
    def _chk(self, rc: int):
        if rc != 0:
            msg = self.lib.khb_last_error(self.ctx)
            raise KhbError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.khb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._chk(self.lib.khb_sync(self.ctx))

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, max(int(nbytes), 16))

    @property
    def launch_count(self) -> int:
        return int(self.lib.khb_launch_count(self.ctx))

    KERNELS = {"pack": 0, "extract": 1, "radix_hist": 2, "onesweep": 3, "unique": 4, "rle_hist": 5, "partition": 6,
               "hash_insert": 7, "hash_count": 8, "bin_partition": 9, "bin_count": 10, "bin_across": 11}

    # ---- multi-GPU exchange over peer memory (csrc/peer.cu) ----
    def peer_alloc(self, world: int, rank: int, key_bytes: int, region_keys: int) -> bytes:
        """Allocate this rank's receive buffer; returns its 64-byte CUDA IPC handle for the other ranks."""
        h = (C.c_ubyte * 64)()
        self._chk(self.lib.khb_peer_alloc(self.ctx, world, rank, key_bytes, int(region_keys), h))
        return bytes(h)

    def peer_open(self, handles: bytes):
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        self._chk(self.lib.khb_peer_open(self.ctx, buf))

    def peer_begin(self):
        self._chk(self.lib.khb_peer_begin(self.ctx))

    def peer_push(self):
        self._chk(self.lib.khb_peer_push(self.ctx))

    def peer_counts(self, world: int):
        counts = np.zeros(world, dtype=np.uint64)
        ovf = C.c_int()
        self._chk(self.lib.khb_peer_counts(self.ctx, counts.ctypes.data, C.byref(ovf)))
        return counts, bool(ovf.value)

    def peer_import(self, recv_counts, k: int, n_groups: int, hashed: bool):
        rc = np.ascontiguousarray(recv_counts, dtype=np.uint64)
        self._chk(self.lib.khb_peer_import(self.ctx, rc.ctypes.data, k, n_groups, int(hashed)))

    def peer_across(self, recv_counts, k: int, n_groups: int, hashed: bool, nbins: int = COUNTER_MAX):
        """peer_import + across_groups without the copy in between (the first radix pass gathers the pushed regions)."""
        rc = np.ascontiguousarray(recv_counts, dtype=np.uint64)
        hist = np.zeros(nbins + 1, dtype=np.uint64)
        st = Stats()
        self._chk(self.lib.khb_peer_across(self.ctx, rc.ctypes.data, k, n_groups, int(hashed), nbins, hist.ctypes.data, C.byref(st)))
        return hist, st.as_dict()

    def peer_unmap(self):
        self._chk(self.lib.khb_peer_unmap(self.ctx))

    def peer_close(self):
        self._chk(self.lib.khb_peer_close(self.ctx))

    @property
    def peer_region_keys(self) -> int:
        return int(self.lib.khb_peer_region_keys(self.ctx))

    # ---- one group on several GPUs (csrc/team.cu, csrc/bins.cu; driver: dist.TeamSharder) ----
    def team_alloc(self, team_size: int, member: int, half_bytes: int) -> bytes:
        """Allocate this member's two record receive buffers; returns the 64-byte CUDA IPC handle for the other members."""
        h = (C.c_ubyte * 64)()
        self._chk(self.lib.khb_team_alloc(self.ctx, team_size, member, int(half_bytes), h))
        return bytes(h)

    def team_open(self, handles: bytes):
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        self._chk(self.lib.khb_team_open(self.ctx, buf))

    def team_unmap(self):
        self._chk(self.lib.khb_team_unmap(self.ctx))

    def team_close(self):
        self._chk(self.lib.khb_team_close(self.ctx))

    def team_plan(self, k: int, tg: TeamGroup) -> dict:
        nb, cap, hb = C.c_uint32(), C.c_uint32(), C.c_uint64()
        self._chk(self.lib.khb_team_plan(self.ctx, k, C.byref(tg), C.byref(nb), C.byref(cap), C.byref(hb)))
        return {"n_bins": int(nb.value), "region_cap": int(cap.value), "half_bytes": int(hb.value)}

    def team_partition(self, source, k: int, tg: TeamGroup) -> dict:
        """K1 + pass P of this member's slice; `source` is a list of host FASTA texts, a StagedFasta or a PackedGroup."""
        info = np.zeros(4, dtype=np.uint64)
        if isinstance(source, PackedGroup):
            self._chk(self.lib.khb_team_partition_packed(self.ctx, k, source.handle, C.byref(tg), info.ctypes.data))
        elif isinstance(source, StagedFasta):
            self._chk(self.lib.khb_team_partition_staged(self.ctx, k, len(source.begin) - 1, source.buf.ptr, source.begin.ctypes.data, C.byref(tg),
                                                         info.ctypes.data))
        else:
            arrs, ptrs, sizes = self._file_tables(source)
            self._chk(self.lib.khb_team_partition_fasta(self.ctx, k, len(arrs), ptrs, sizes, C.byref(tg), info.ctypes.data))
        return {"overflow": bool(int(info[0]) & 1), "area_overflow": bool(int(info[0]) & 2), "fullest_region": int(info[1]), "windows": int(info[2]),
                "bases": int(info[3])}

    def team_count(self, k: int, tg: TeamGroup, nbins: int = COUNTER_MAX, keep_set: bool = True):
        """Passes C and B over the bins this member owns: (partial step_4 histogram, stats)."""
        hist = np.zeros(nbins + 1, dtype=np.uint64)
        st = Stats()
        self._chk(self.lib.khb_team_count(self.ctx, k, C.byref(tg), nbins, hist.ctypes.data, int(keep_set), C.byref(st)))
        return hist, st.as_dict()

    GROUP_MODES = {"auto": 0, "single-sort": 1, "two-sort": 2, "hash": 3, "bins": 4}

    def set_group_mode(self, mode: str):
        """How the group stage finds shared k-mers (include/khoice_b200.h: KHB_GROUP_*); results do not depend on it."""
        self._chk(self.lib.khb_set_group_mode(self.ctx, self.GROUP_MODES[mode]))

    @property
    def hash_overflows(self) -> int:
        return int(self.lib.khb_hash_overflows(self.ctx))

    def healthy(self) -> bool:
        """False once the context carries a sticky CUDA error (every later call would fail): a long-lived worker exits then."""
        try:
            self._chk(self.lib.khb_sync(self.ctx))
            return True
        except KhbError:
            return False

    @property
    def bins_counters(self) -> dict:
        """Minimizer-bin path: groups handed to the sort path, bins redone in hash classes, groups partitioned a second time."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.khb_bins_counters(self.ctx, C.byref(a), C.byref(b), C.byref(c))
        d, e = C.c_uint64(), C.c_uint64()
        self.lib.khb_across_counters(self.ctx, C.byref(d), C.byref(e))
        return {"fallbacks": int(a.value), "big_bins": int(b.value), "repartitions": int(c.value), "across_by_bins": int(d.value), "across_by_sort": int(e.value)}

    def profile_enable(self, on: bool = True):
        """Bracket every kernel launch with CUDA events (clears earlier records)."""
        self._chk(self.lib.khb_profile_enable(self.ctx, int(on)))

    def profile_read(self) -> dict:
        """{kernel: {launches, ms, alg_bytes}} accumulated since profile_enable()."""
        out = {}
        for name, kid in self.KERNELS.items():
            n, ms, b = C.c_uint64(), C.c_double(), C.c_uint64()
            self._chk(self.lib.khb_profile_read(self.ctx, kid, C.byref(n), C.byref(ms), C.byref(b)))
            out[name] = {"launches": int(n.value), "ms": float(ms.value), "alg_bytes": int(b.value)}
        return out

    @property
    def stream_ptr(self) -> int:
        return int(self.lib.khb_stream(self.ctx) or 0)

    def device_info(self) -> dict:
        sms, fr, tot = C.c_int(), C.c_size_t(), C.c_size_t()
        self._chk(self.lib.khb_device_info(self.ctx, C.byref(sms), C.byref(fr), C.byref(tot)))
        return {"num_sms": sms.value, "free_bytes": fr.value, "total_bytes": tot.value}

    @staticmethod
    def _file_tables(files: Sequence):
        arrs = [np.frombuffer(f, dtype=np.uint8) if isinstance(f, (bytes, bytearray, memoryview)) else np.ascontiguousarray(f, dtype=np.uint8)
                for f in files]
        n = len(arrs)
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data if a.size else None for a in arrs])
        sizes = (C.c_size_t * max(n, 1))(*[a.size for a in arrs])
        return arrs, ptrs, sizes

    # -- kernels -----------------------------------------------------------------------------------------
    def stage_fasta(self, files: Sequence) -> StagedFasta:
        """Host FASTA texts -> device staging buffer (tile-aligned, separator-filled)."""
        arrs, ptrs, sizes = self._file_tables(files)
        n = len(arrs)
        total = int(self.lib.khb_staged_size(n, sizes))
        buf = self.alloc(total)
        begin = np.zeros(n + 1, dtype=np.uint64)
        self._chk(self.lib.khb_stage_fasta(self.ctx, n, ptrs, sizes, buf.ptr, buf.nbytes, begin.ctypes.data))
        self.sync()
        return StagedFasta(buf, begin)

    def pack_fasta(self, staged: StagedFasta):
        """K1.  Returns dict(codes, valid (DeviceBuffers), n_symbols, n_breaks, tile_base (np.uint64))."""
        nbytes = staged.nbytes
        ntiles = nbytes // FASTA_TILE
        cw, vw = nbytes // 32 + 4, nbytes // 32 + 4
        codes, valid = self.alloc(cw * 8), self.alloc(vw * 4)
        tile_base, counts = self.alloc((ntiles + 1) * 8), self.alloc(16)
        self._chk(self.lib.khb_pack_fasta(self.ctx, staged.buf.ptr, nbytes, codes.ptr, valid.ptr, nbytes, tile_base.ptr, counts.ptr))
        cnt = counts.download(np.uint64, 2)
        tb = tile_base.download(np.uint64, ntiles + 1)
        return {"codes": codes, "valid": valid, "n_symbols": int(cnt[0]), "n_breaks": int(cnt[1]), "tile_base": tb,
                "codes_words": cw, "valid_words": vw}

    def extract_kmers(self, packed: dict, k: int, hashed: bool = False) -> DeviceBuffer:
        """K2.  Returns a DeviceBuffer of n_symbols k-mer words (sentinel = all ones); hashed=True applies the
        bijective mixer of the fused path."""
        n = packed["n_symbols"]
        keys = self.alloc((n + 4) * 8 * key_words(k))
        fn = self.lib.khb_extract_kmers_hashed if hashed else self.lib.khb_extract_kmers
        self._chk(fn(self.ctx, packed["codes"].ptr, packed["valid"].ptr, n, k, keys.ptr))
        self.sync()
        return keys

    def remix_keys(self, keys: DeviceBuffer, n: int, k: int, inverse: bool):
        """Apply the mixer (inverse=False) or its inverse (inverse=True) to n keys in place."""
        self._chk(self.lib.khb_remix_keys(self.ctx, keys.ptr, n, k, int(inverse)))
        self.sync()

    def prefix_plan(self, k: int, n_max: int):
        fb, np_ = C.c_int(), C.c_int()
        self._chk(self.lib.khb_prefix_plan(k, n_max, C.byref(fb), C.byref(np_)))
        return fb.value, np_.value

    def sort_key_bits(self, keys: DeviceBuffer, n: int, k: int, first_bit: int, npass: int,
                      seg_off: Optional[Sequence[int]] = None) -> DeviceBuffer:
        """K3 on the digits [first_bit, first_bit + 8*npass) only."""
        seg = np.ascontiguousarray(seg_off if seg_off is not None else [0, n], dtype=np.uint64)
        tmp = self.alloc(keys.nbytes)
        flag = C.c_int(0)
        self._chk(self.lib.khb_sort_key_bits(self.ctx, keys.ptr, tmp.ptr, seg.ctypes.data, len(seg) - 1, 8 * key_words(k),
                                             first_bit, npass, C.byref(flag)))
        self.sync()
        if flag.value:
            return tmp
        tmp.free()
        return keys

    def resolve_unique(self, sorted_keys: DeviceBuffer, n: int, k: int, prefix_shift: int) -> Tuple[DeviceBuffer, int]:
        out = self.alloc(sorted_keys.nbytes)
        cnt = C.c_uint64(0)
        self._chk(self.lib.khb_resolve_unique(self.ctx, sorted_keys.ptr, n, k, prefix_shift, out.ptr, C.byref(cnt)))
        return out, int(cnt.value)

    def resolve_count(self, sorted_keys: DeviceBuffer, n: int, k: int, prefix_shift: int, nbins: int = COUNTER_MAX,
                      cs: int = COUNTER_MAX, want_keys: bool = False):
        hist = np.zeros(nbins + 1, dtype=np.uint64)
        runs = C.c_uint64(0)
        ok = self.alloc(sorted_keys.nbytes) if want_keys else None
        self._chk(self.lib.khb_resolve_count(self.ctx, sorted_keys.ptr, n, k, prefix_shift, cs, nbins, hist.ctypes.data,
                                             ok.ptr if ok else None, C.byref(runs)))
        return hist, int(runs.value), ok

    def sort_keys(self, keys: DeviceBuffer, n: int, k: int, seg_off: Optional[Sequence[int]] = None) -> DeviceBuffer:
        """K3.  Sorts segments of `keys` (default: one segment [0, n)); returns the buffer holding the result."""
        seg = np.ascontiguousarray(seg_off if seg_off is not None else [0, n], dtype=np.uint64)
        tmp = self.alloc(keys.nbytes)
        flag = C.c_int(0)
        self._chk(self.lib.khb_sort_keys(self.ctx, keys.ptr, tmp.ptr, seg.ctypes.data, len(seg) - 1, k, C.byref(flag)))
        self.sync()
        if flag.value:
            return tmp
        tmp.free()
        return keys

    def unique(self, sorted_keys: DeviceBuffer, n: int, k: int) -> Tuple[DeviceBuffer, int]:
        """K4.  Returns (buffer of distinct non-sentinel keys, count)."""
        out = self.alloc(sorted_keys.nbytes)
        cnt = C.c_uint64(0)
        self._chk(self.lib.khb_unique(self.ctx, sorted_keys.ptr, n, k, out.ptr, C.byref(cnt)))
        return out, int(cnt.value)

    def count_runs(self, sorted_keys: DeviceBuffer, n: int, k: int, nbins: int = COUNTER_MAX, cs: int = COUNTER_MAX,
                   want_keys: bool = False, want_counts: bool = False):
        """K5/K6.  Returns (hist uint64[nbins+1], n_distinct, keys DeviceBuffer|None, counts DeviceBuffer|None)."""
        hist = np.zeros(nbins + 1, dtype=np.uint64)
        runs = C.c_uint64(0)
        ok = self.alloc(sorted_keys.nbytes) if want_keys else None
        oc = self.alloc(max(n, 1) * 4) if want_counts else None
        self._chk(self.lib.khb_count_runs(self.ctx, sorted_keys.ptr, n, k, cs, nbins, hist.ctypes.data,
                                          ok.ptr if ok else None, oc.ptr if oc else None, C.byref(runs)))
        return hist, int(runs.value), ok, oc

    def partition_by_hash(self, keys: DeviceBuffer, n: int, k: int, n_parts: int) -> Tuple[DeviceBuffer, np.ndarray]:
        """K7.  Returns (keys grouped by hash bucket, uint64 bucket offsets [n_parts+1])."""
        out = self.alloc(keys.nbytes)
        off = np.zeros(n_parts + 1, dtype=np.uint64)
        self._chk(self.lib.khb_partition_by_hash(self.ctx, keys.ptr, n, k, n_parts, out.ptr, off.ctypes.data))
        self.sync()
        return out, off

    # -- fused stages ------------------------------------------------------------------------------------
    def group_from_fasta(self, files: Sequence, k: int, nbins: int = COUNTER_MAX, keep_set: bool = True):
        """Rules build_kmc_database_on_genome .. within_group_union_histogram (+ build_group_kmer_set) for one
        (k, group), from FASTA texts in host memory.  Returns (step_4 histogram uint64[nbins+1], stats dict)."""
        arrs, ptrs, sizes = self._file_tables(files)
        hist = np.zeros(nbins + 1, dtype=np.uint64)
        st = Stats()
        self._chk(self.lib.khb_group_from_fasta(self.ctx, k, len(arrs), ptrs, sizes, nbins, hist.ctypes.data, int(keep_set), C.byref(st)))
        return hist, st.as_dict()

    def prefetch_fasta(self, files: Sequence):
        """Start the host->device copy of the NEXT group's FASTA texts on a second stream (double buffering).
        The caller must keep `files` alive and unchanged until the matching group_from_fasta call."""
        arrs, ptrs, sizes = self._file_tables(files)
        self._pf_keepalive = getattr(self, "_pf_keepalive", [])[-2:] + [(arrs, ptrs, sizes)]
        self._chk(self.lib.khb_group_prefetch_fasta(self.ctx, len(arrs), ptrs, sizes))

    def group_from_staged(self, staged: StagedFasta, k: int, nbins: int = COUNTER_MAX, keep_set: bool = True,
                          first: int = 0, count: Optional[int] = None):
        """Same stage on text already resident in HBM (files first .. first+count of `staged`)."""
        n = (len(staged.begin) - 1 - first) if count is None else count
        base = int(staged.begin[first])
        begin = np.ascontiguousarray(staged.begin[first:first + n + 1] - np.uint64(base), dtype=np.uint64)
        hist = np.zeros(nbins + 1, dtype=np.uint64)
        st = Stats()
        self._chk(self.lib.khb_group_from_staged(self.ctx, k, n, staged.buf.ptr + base, begin.ctypes.data, nbins,
                                                 hist.ctypes.data, int(keep_set), C.byref(st)))
        return hist, st.as_dict()

    def pack_group(self, files: Sequence) -> "PackedGroup":
        """K1 once for a group; the 2-bit stream stays in HBM for a sweep over k (see group_from_packed)."""
        arrs, ptrs, sizes = self._file_tables(files)
        h = _P()
        self._chk(self.lib.khb_pack_group(self.ctx, len(arrs), ptrs, sizes, C.byref(h)))
        return PackedGroup(self, h)

    def group_from_packed(self, packed: "PackedGroup", k: int, nbins: int = COUNTER_MAX, keep_set: bool = True):
        """K2..K5 for one k on a packed group; same results as group_from_fasta."""
        hist = np.zeros(nbins + 1, dtype=np.uint64)
        st = Stats()
        self._chk(self.lib.khb_group_from_packed(self.ctx, k, packed.handle, nbins, hist.ctypes.data, int(keep_set), C.byref(st)))
        return hist, st.as_dict()

    def across_groups(self, nbins: int = COUNTER_MAX):
        """Rules across_group_union + across_group_union_histogram for one k over the retained group sets."""
        hist = np.zeros(nbins + 1, dtype=np.uint64)
        st = Stats()
        self._chk(self.lib.khb_across_groups(self.ctx, nbins, hist.ctypes.data, C.byref(st)))
        return hist, st.as_dict()

    # -- experiment type 2 (pivot analysis, /root/reference/workflow/rules/exp_type_2.smk) ----------------------
    def pivot_group_from_packed(self, packed: "PackedGroup", k: int, nbins: int = COUNTER_MAX, keep_sets: bool = True):
        """One (k, dataset) of experiment type 2.  `packed` = pack_group(rest-of-set genomes + [pivot]), the pivot LAST.
        Returns (hist uint64[nbins+1], stats): hist[1] = size of `pivot kmers_subtract rest` (exp_type_2.smk:367-380),
        hist[c >= 2] = histogram of `pivot intersect rest -ocsum` (exp_type_2.smk:354-365), c = 1 + #rest genomes."""
        hist = np.zeros(nbins + 1, dtype=np.uint64)
        st = Stats()
        self._chk(self.lib.khb_pivot_group_from_packed(self.ctx, k, packed.handle, nbins, hist.ctypes.data, int(keep_sets), C.byref(st)))
        return hist, st.as_dict()

    def pivot_sets_info(self) -> dict:
        g, npk, nuk = C.c_int(), C.c_uint64(), C.c_uint64()
        self._chk(self.lib.khb_pivot_sets_info(self.ctx, C.byref(g), C.byref(npk), C.byref(nuk)))
        return {"n_pivots": g.value, "n_pivot_keys": npk.value, "n_union_keys": nuk.value}

    def pivot_across(self, nbins: int = COUNTER_MAX):
        """Across-group stage of experiment type 2 for all retained pivots (exp_type_2.smk:440-508).
        Returns (hists uint64[G, nbins+1], stats): hists[j, c] = #k-mers of pivot j+1 seen in c-1 OTHER groups' unions."""
        g = self.pivot_sets_info()["n_pivots"]
        hists = np.zeros((max(g, 1), nbins + 1), dtype=np.uint64)
        st = Stats()
        self._chk(self.lib.khb_pivot_across(self.ctx, nbins, hists.ctypes.data, C.byref(st)))
        return hists[:g], st.as_dict()

    def sorted_lookup(self, a: np.ndarray, b: np.ndarray, k: int) -> np.ndarray:
        """Position of every key of `a` in the sorted, duplicate-free `b` (int64, -1 = absent): the join behind the
        rule-compatible `kmc_tools simple A B intersect|kmers_subtract` (exp_type_2.smk:354-380)."""
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        na, nb = a.shape[0], b.shape[0]
        if na == 0:
            return np.empty(0, dtype=np.int64)
        da, db, di = self.alloc(max(a.nbytes, 16)), self.alloc(max(b.nbytes, 16)), self.alloc(na * 8)
        try:
            da.upload(a)
            if nb:
                db.upload(b)
            self._chk(self.lib.khb_sorted_lookup(self.ctx, da.ptr, na, db.ptr, nb, k, di.ptr))
            return di.download(np.uint64, na).view(np.int64)
        finally:
            da.free(); db.free(); di.free()

    # -- experiment type 4 (confusion matrix, exp_type_4.smk + src/merge_lists.py) -----------------------------------
    def kmer_counts(self, fasta: bytes, k: int, cs: int = 255):
        """`kmc -fm -k{k} -ci1` of one FASTA text (exp_type_4.smk:146-153): (keys DeviceBuffer [canonical, ascending],
        counts uint32 ndarray, n).  The keys stay on the device for group_membership()."""
        staged = self.stage_fasta([fasta])
        packed = self.pack_fasta(staged)
        n = packed["n_symbols"]
        keys = self.extract_kmers(packed, k)
        srt = self.sort_keys(keys, n, k)
        _, runs, ok, oc = self.count_runs(srt, n, k, nbins=16, cs=cs, want_keys=True, want_counts=True)
        counts = oc.download(np.uint32, runs)
        for b in (oc, srt, keys, staged.buf, packed["codes"], packed["valid"]):
            if b is not ok:
                b.free()
        return ok, counts, runs

    def group_membership(self, group_off: Sequence[int], query_bufs: Sequence[DeviceBuffer], query_sizes: Sequence[int], k: int) -> np.ndarray:
        """For every k-mer of the query sets (device buffers of canonical, ascending keys) the bitmask of retained group
        sets containing it: uint64 [sum(query_sizes), words].  group_off[g] = n_keys of the store before group g."""
        g = len(group_off) - 1
        words = max(1, (g + 63) // 64)
        w = key_words(k)
        q_off = np.zeros(len(query_bufs) + 1, dtype=np.uint64)
        q_off[1:] = np.cumsum(np.asarray(query_sizes, dtype=np.uint64))
        nq = int(q_off[-1])
        cat = self.alloc(max(nq, 1) * 8 * w + 64)
        mask = self.alloc(max(nq, 1) * 8 * words)
        try:
            for buf, off, sz in zip(query_bufs, q_off[:-1], query_sizes):
                if sz:
                    self._chk(self.lib.khb_memcpy_d2d(self.ctx, cat.ptr + int(off) * 8 * w, buf.ptr, int(sz) * 8 * w))
            goff = np.ascontiguousarray(group_off, dtype=np.uint64)
            self._chk(self.lib.khb_group_membership(self.ctx, g, goff.ctypes.data, cat.ptr, len(query_bufs), q_off.ctypes.data, mask.ptr, words))
            return mask.download(np.uint64, nq * words).reshape(nq, words)
        finally:
            cat.free()
            mask.free()

    def read_votes(self, reads: Sequence[bytes], k: int, pivot_keys: DeviceBuffer, n_pivot: int, masks: np.ndarray, n_groups: int):
        """Experiment type 6, read level: per-read votes (src/merge_lists.py:157-174) on the GPU.  reads: upper-case ACGT
        strings (merge_lists.split_reads); pivot_keys / masks: the pivot's ascending distinct k-mers (kmer_counts) and their
        membership masks (group_membership).  Returns (votes float64 [n_reads, n_groups], unmatched uint32 [n_reads])."""
        n_reads = len(reads)
        if n_reads == 0:
            return np.zeros((0, n_groups), np.float64), np.zeros(0, np.uint32)
        lens = np.fromiter((len(r) for r in reads), dtype=np.int64, count=n_reads)
        text = b"".join(b">\n" + r + b"\n" for r in reads)      # every read its own record: windows never span two reads
        first = (np.arange(1, n_reads + 1, dtype=np.int64) + np.concatenate([[0], np.cumsum(lens)[:-1]])).astype(np.uint64)
        nwin = np.maximum(lens - k + 1, 0).astype(np.uint32)
        masks = np.ascontiguousarray(masks, dtype=np.uint64).reshape(n_pivot, -1)
        words = masks.shape[1]
        staged = self.stage_fasta([text])
        packed = self.pack_fasta(staged)
        n = packed["n_symbols"]
        if not 0 <= n - (n_reads + int(lens.sum())) <= 1:        # the staging filler behind the text ends with one more break symbol
            raise KhbError(-1, f"read_votes: {n} symbols packed, {n_reads + int(lens.sum())} expected")
        keys = self.extract_kmers(packed, k)
        bufs = [keys, staged.buf, packed["codes"], packed["valid"]]
        try:
            index = self.alloc(max(n, 1) * 8); bufs.append(index)
            self._chk(self.lib.khb_sorted_lookup(self.ctx, keys.ptr, n, pivot_keys.ptr, n_pivot, k, index.ptr))
            d_mask = self.alloc(max(masks.nbytes, 16)); bufs.append(d_mask)
            if masks.nbytes:
                d_mask.upload(masks)
            d_first = self.alloc(n_reads * 8); bufs.append(d_first)
            d_first.upload(first)
            d_nwin = self.alloc(n_reads * 4); bufs.append(d_nwin)
            d_nwin.upload(nwin)
            d_votes = self.alloc(n_reads * n_groups * 8); bufs.append(d_votes)
            d_un = self.alloc(n_reads * 4); bufs.append(d_un)
            self._chk(self.lib.khb_read_votes(self.ctx, index.ptr, d_mask.ptr, words, n_groups, d_first.ptr, d_nwin.ptr, n_reads, d_votes.ptr, d_un.ptr))
            return d_votes.download(np.float64, n_reads * n_groups).reshape(n_reads, n_groups), d_un.download(np.uint32, n_reads)
        finally:
            for b in bufs:
                b.free()

    def bins_partition(self, files: Sequence, k: int, n_bins: int):
        """The partition pass of the minimizer-bin group stage on its own (include/khoice_b200.h: khb_bins_partition).
        Returns (records uint32 [n_bins, chunks], windows uint32 [n_bins, chunks], first symbol of every genome in the group's stream),
        chunks = ceil(len(files) / 64)."""
        staged = self.stage_fasta(files)
        packed = self.pack_fasta(staged)
        n = len(staged.begin) - 1
        seg = np.zeros(n + 1, dtype=np.uint64)
        for g in range(n):
            seg[g] = packed["tile_base"][int(staged.begin[g]) // FASTA_TILE]
        seg[0] = 0
        seg[n] = packed["n_symbols"]
        d_seg = self.alloc((n + 1) * 8)
        try:
            d_seg.upload(seg)
            chunks = (n + 63) // 64
            rec = np.zeros(n_bins * chunks, dtype=np.uint32)
            win = np.zeros(n_bins * chunks, dtype=np.uint32)
            self._chk(self.lib.khb_bins_partition(self.ctx, packed["codes"].ptr, packed["valid"].ptr, packed["n_symbols"], d_seg.ptr, n, k, n_bins,
                                                  rec.ctypes.data, win.ctypes.data))
            return rec.reshape(n_bins, chunks), win.reshape(n_bins, chunks), seg[:n].astype(np.int64)
        finally:
            for b in (d_seg, staged.buf, packed["codes"], packed["valid"]):
                b.free()

    def group_sets_info(self) -> dict:
        k, g, n = C.c_int(), C.c_int(), C.c_uint64()
        self._chk(self.lib.khb_group_sets_info(self.ctx, C.byref(k), C.byref(g), C.byref(n)))
        return {"k": k.value, "n_groups": g.value, "n_keys": n.value}

    def group_sets_download(self) -> np.ndarray:
        """Concatenation of the retained group sets as canonical k-mer values (host), group after group; the
        order inside a group is the fused path's prefix order, not numeric order."""
        info = self.group_sets_info()
        n = info["n_keys"]
        w = key_words(info["k"]) if info["k"] else 1
        out = np.empty(n * w, dtype=np.uint64)
        if out.nbytes:
            self._chk(self.lib.khb_group_sets_export(self.ctx, out.ctypes.data))
        return out.reshape(key_shape(n, info["k"] or 1))

    @property
    def group_sets_hashed(self) -> bool:
        return bool(self.lib.khb_group_sets_hashed(self.ctx))

    def group_sets_device(self) -> Tuple[int, int]:
        p, n = _P(), C.c_uint64()
        self._chk(self.lib.khb_group_sets_device(self.ctx, C.byref(p), C.byref(n)))
        return (p.value or 0), int(n.value)

    def group_sets_append_host(self, keys: np.ndarray, k: int, n_groups: int = 1):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        self._chk(self.lib.khb_group_sets_append_host(self.ctx, k, keys.ctypes.data, keys.shape[0], n_groups))

    def group_sets_append_device(self, ptr: int, n_keys: int, k: int, n_groups: int = 1, hashed: bool = False):
        self._chk(self.lib.khb_group_sets_append_device(self.ctx, k, ptr, n_keys, n_groups, int(hashed)))

    def group_sets_reset(self):
        self._chk(self.lib.khb_group_sets_reset(self.ctx))
