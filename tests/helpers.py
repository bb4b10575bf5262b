"""Shared test inputs and host-side reference helpers (numpy / pure Python)."""
import numpy as np

EDGE_FASTAS = [
    b">r1 plain\nACGTACGTACGTTTGACCA\nGGATTACAGATTACA\n",
    b">r2 lower and N\nacgtnnACGTTGCAacgtNNNNNNNNACGGTCA\nNACGT\n",
    b">r3 crlf\r\nACGTTGCA\r\nGGCCAATT\r\n>r3b\r\nTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT\r\n",
    b">r4 iupac\nACGTRYKMACGTSWBDHVACGTACGTACGTACGTACGTACGTACGTACGTAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n",
    b">short\nACG\n>empty\n>another\nAC\nGT\nAC\n",
    b"ACGTACGTTTGACA\n>no header at start above\nGGGCCCATATAT",          # no trailing newline
    b">header only no newline",
    b"",
    b">x\n" + b"ACGT" * 40 + b"\n" + b"TTGCA" * 33 + b"\n>y " + b"h" * 300 + b"\n" + b"GATTACA" * 50,
    b"\n\n>z\n\nAC\n\nGT\n>>double\nACGTAC>GTAC\nGGGG\n",
]


def random_fasta(rng, n_bases, line=80, p_n=0.002, p_lower=0.05, n_records=3, crlf=False):
    seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=n_bases)
    mask = rng.random(n_bases) < p_n
    seq[mask] = rng.choice(np.frombuffer(b"NRYKMSWnx-*", np.uint8), size=int(mask.sum()))
    low = rng.random(n_bases) < p_lower
    seq[low] |= 0x20
    cuts = sorted(rng.integers(0, n_bases + 1, size=n_records - 1).tolist())
    cuts = [0] + cuts + [n_bases]
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(n_records):
        out.append(b">rec%d some description" % i + nl)
        s = seq[cuts[i]:cuts[i + 1]].tobytes()
        for j in range(0, len(s), line):
            out.append(s[j:j + line] + nl)
    return b"".join(out)


def symbol_stream(text: bytes):
    """Reference of K1: (codes list, valid list) for a FASTA byte string, incl. one break per header."""
    codes, valid = [], []
    in_hdr = False
    symbol_stream.breaks = 0
    for b in text:
        if in_hdr:
            if b == 10:
                in_hdr = False
            continue
        if b == 62:  # '>'
            codes.append(0); valid.append(0); in_hdr = True
            symbol_stream.breaks += 1
            continue
        if b in (10, 13):
            continue
        pos = b"ACGTacgt".find(bytes([b]))
        c = pos % 4 if pos >= 0 else None
        if c is None:
            codes.append(0); valid.append(0)
        else:
            codes.append(c); valid.append(1)
    return codes, valid


def unpack_codes(codes_words: np.ndarray, valid_words: np.ndarray, n: int):
    """Decode the K1 layout (MSB-first u64 codes, MSB-first u32 validity) into per-symbol arrays."""
    idx = np.arange(n, dtype=np.uint64)
    cw = codes_words[(idx >> np.uint64(5)).astype(np.int64)]
    sh = (np.uint64(62) - np.uint64(2) * (idx & np.uint64(31)))
    c = ((cw >> sh) & np.uint64(3)).astype(np.uint8)
    vw = valid_words[(idx >> np.uint64(5)).astype(np.int64)].astype(np.uint64)
    v = ((vw >> (np.uint64(31) - (idx & np.uint64(31)))) & np.uint64(1)).astype(np.uint8)
    return c, v


def as_py(keys: np.ndarray):
    """k-mer words -> list of Python ints (handles the [n,2] lo/hi layout)."""
    if keys.ndim == 1:
        return [int(x) for x in keys]
    return [(int(h) << 64) | int(l) for l, h in keys]


def sort_rows(keys: np.ndarray) -> np.ndarray:
    """Numeric sort of k-mer words on the host."""
    if keys.ndim == 1:
        return np.sort(keys)
    order = np.lexsort((keys[:, 0], keys[:, 1]))
    return keys[order]


def is_sentinel(keys: np.ndarray) -> np.ndarray:
    full = np.uint64(0xFFFFFFFFFFFFFFFF)
    return keys == full if keys.ndim == 1 else (keys[:, 0] == full) & (keys[:, 1] == full)


def synth_genomes(n_genomes: int, group: int = 1, genome_len: int = 5_000_000, procs: int = 16):
    """[fasta bytes] of one synthetic group at BASELINE size, generated in forked worker processes (they only run numpy; the
    session's CUDA context is never touched in a child -- the same pattern as bench.py's cpu_baseline leg)."""
    import multiprocessing as mp
    import os
    from concurrent.futures import ProcessPoolExecutor
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=max(group, 1), genomes_per_group=n_genomes, genome_len=genome_len)
    procs = max(1, min(procs, os.cpu_count() or 1, n_genomes))
    if procs == 1:
        return [synth.make_genome(cfg, group, i) for i in range(1, n_genomes + 1)]
    with ProcessPoolExecutor(procs, mp_context=mp.get_context("fork")) as pool:
        return list(pool.map(synth.make_genome, [cfg] * n_genomes, [group] * n_genomes, range(1, n_genomes + 1), chunksize=max(1, n_genomes // (4 * procs))))
