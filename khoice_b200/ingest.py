"""Ingest side of the fused drivers: the reference hands `kmc` one `.fna.gz` at a time
and KMC inflates it itself (/root/reference/workflow/rules/exp_type_1.smk:156-163); here the host inflates the files of
a group in parallel (zlib releases the GIL) and stays ONE GROUP AHEAD of the GPU, so that inflating group g+1 overlaps
the kernels of group g.  Multi-member gzip files are handled by `gzip` (rule R9)."""
from __future__ import annotations

import gzip
import os
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Dict, Hashable, List, Optional, Sequence


def read_fasta(path: str) -> bytes:
    """FASTA text of a .fna.gz (multi-member gzip handled) or plain file."""
    if path.endswith(".gz"):
        with gzip.open(path, "rb") as fd:
            return fd.read()
    with open(path, "rb") as fd:
        return fd.read()


def default_workers() -> int:
    return max(1, min(32, int(os.environ.get("KHB_INFLATE_THREADS", os.cpu_count() or 1))))


def read_many(paths: Sequence[str], workers: Optional[int] = None) -> List[bytes]:
    """All files, inflated concurrently, in the order given."""
    if len(paths) <= 1:
        return [read_fasta(p) for p in paths]
    with ThreadPoolExecutor(max_workers=min(workers or default_workers(), len(paths))) as pool:
        return list(pool.map(read_fasta, paths))


class GroupReader:
    """Inflates groups of files in a fixed order with one group of look-ahead.

    reader = GroupReader({1: [paths...], 2: [...]}, order=[1, 2]);  texts = reader.get(1)   # group 2 starts inflating now
    """

    def __init__(self, paths: Dict[Hashable, Sequence[str]], order: Sequence[Hashable], workers: Optional[int] = None):
        self.paths = {k: list(v) for k, v in paths.items()}
        self.order = list(order)
        self.pool = ThreadPoolExecutor(max_workers=workers or default_workers())
        self.pending: Dict[Hashable, List[Future]] = {}
        if self.order:
            self._submit(self.order[0])

    def _submit(self, key) -> None:
        if key not in self.pending and key in self.paths:
            self.pending[key] = [self.pool.submit(read_fasta, p) for p in self.paths[key]]

    def get(self, key) -> List[bytes]:
        self._submit(key)
        if key in self.order:
            i = self.order.index(key)
            if i + 1 < len(self.order):
                self._submit(self.order[i + 1])   # look-ahead: overlaps the caller's GPU work on `key`
        return [f.result() for f in self.pending.pop(key)]

    def close(self) -> None:
        for futs in self.pending.values():
            for f in futs:
                f.cancel()
        self.pending.clear()
        self.pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
