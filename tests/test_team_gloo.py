"""ONE group on several ranks (dist.TeamSharder) over gloo, on a stand-in engine (TEST ONLY -- the product engine is
csrc/team.cu + csrc/bins.cu and needs a B200; tests/test_gpu_team.py runs that one with two processes on one GPU).

What is checked on CPU is the host protocol: genome slices padded to chunks of 64 ids, one writer per (bin, chunk) region, the
all-gather that is barrier + overflow flags + hints, the retry with larger regions, the two alternating receive buffers, and that
the members' partial histograms add up to the oracle's step_4 histogram whatever the team size."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_BINS = 12


class FakeTeamEngine:
    """The calls of khb_team_* with "peer memory" = .npy files in a directory every member sees.  A region = (bin, chunk of 64 genome
    ids); a record = (k-mer, genome id); pass P bins a slice's k-mers by a hash of the k-mer, counting = union with counter sum."""

    def __init__(self, shared_dir, first_cap):
        from oracle import oracle as O
        self.O, self.dir, self.first_cap = O, shared_dir, first_cap
        self.device = 0
        self.allocs = 0
        self.partition_calls = []
        self.adapter = None

    def team_alloc(self, team_size, member, half_bytes):
        self.T, self.member, self.half = team_size, member, int(half_bytes)
        self.allocs += 1
        return bytes([member]) + bytes(63)

    def team_open(self, handles):
        assert [handles[64 * t] for t in range(self.T)] == list(range(self.T))

    def team_unmap(self):
        pass

    def team_close(self):
        pass

    def team_plan(self, k, tg):
        cap = int(tg.region_cap) or self.first_cap
        bpo = -(-N_BINS // self.T)
        return {"n_bins": N_BINS, "region_cap": cap, "half_bytes": bpo * tg.n_chunks_total * cap * 1024}

    @staticmethod
    def _bin_of(keys):
        flat = keys.reshape(keys.shape[0], -1)
        h = flat[:, 0] * np.uint64(0x9E3779B97F4A7C15)
        if flat.shape[1] == 2:
            h = h ^ (flat[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F))
        return ((h >> np.uint64(33)) % np.uint64(N_BINS)).astype(np.int64)

    def team_partition(self, source, k, tg):
        source = getattr(source, "texts", source)
        plan = self.team_plan(k, tg)
        assert plan["half_bytes"] <= self.half, "the driver must have grown the buffers first"
        cap, bpo = plan["region_cap"], -(-N_BINS // self.T)
        self.partition_calls.append((int(tg.parity), cap))
        sets = [self.O.genome_set(f, k) for f in source]
        windows = sum(self.O.kmers(f, k)[1] for f in source)
        fullest, overflow = 0, False
        per_owner = {t: [] for t in range(self.T)}
        for i, keys in enumerate(sets):
            gid = 64 * int(tg.chunk_base) + i
            b = self._bin_of(keys)
            for bin_ in range(N_BINS):
                part = keys[b == bin_]
                per_owner[bin_ // bpo].append((bin_, gid, part))
        # region sizes: records of one (bin, chunk)
        sizes = {}
        for t, recs in per_owner.items():
            for bin_, gid, part in recs:
                sizes[(bin_, gid // 64)] = sizes.get((bin_, gid // 64), 0) + part.shape[0]
        if sizes:
            fullest = max(sizes.values())
            overflow = fullest > cap
        for t, recs in per_owner.items():
            np.save(os.path.join(self.dir, f"p{int(tg.parity)}_from{self.member}_to{t}.npy"), np.array([(b, g, p) for b, g, p in recs], dtype=object),
                    allow_pickle=True)
        return {"overflow": overflow, "fullest_region": fullest, "windows": windows, "bases": windows}

    def team_count(self, k, tg, nbins=64, keep_set=True):
        by_genome = {}
        for t in range(self.T):
            recs = np.load(os.path.join(self.dir, f"p{int(tg.parity)}_from{t}_to{self.member}.npy"), allow_pickle=True)
            for b, g, part in recs:
                by_genome.setdefault(int(g), []).append(part)
        w = 1 if k <= 32 else 2
        sets = [self.O.sort_unique(np.concatenate(v, axis=0), k) for v in by_genome.values() if sum(p.shape[0] for p in v)]
        if not sets:
            return np.zeros(nbins + 1, dtype=np.uint64), {"distinct": 0}
        keys, counts = self.O.union_sum(sets, k)
        if keep_set and self.adapter is not None:
            self.adapter.sets.append(keys)             # the distinct k-mers of this member's bins enter its group-set store
            self.adapter.k = k
        return self.O.histogram(counts, nbins), {"distinct": int(keys.shape[0])}


def _groups(n_groups, n_genomes):
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=n_groups, genomes_per_group=n_genomes, genome_len=6_000, seed=99)
    return {g: [synth.make_genome(cfg, g, i) for i in range(1, n_genomes + 1)] for g in range(1, n_groups + 1)}


def _worker(rank, world, port, k, n_genomes, first_cap, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from khoice_b200 import dist as kd
    kd.init_from_env("gloo")
    groups = _groups(3, n_genomes)
    eng = FakeTeamEngine(out_dir, first_cap)
    ts = kd.TeamSharder(eng, world, rank)
    slices = kd.genome_slices(n_genomes, world)
    sizes = [hi - lo for lo, hi in slices]
    lo, hi = slices[rank]
    for g in sorted(groups):
        n_sym = sum(len(f) for f in groups[g])
        hist, st = ts.run_group(groups[g][lo:hi], k, n_genomes, sizes, n_sym, nbins=64)
        h = torch.from_numpy(hist.astype(np.int64))
        dist.all_reduce(h)
        ref = np.load(os.path.join(out_dir, f"ref_{g}.npy"))
        assert np.array_equal(h.numpy().astype(np.uint64), ref), (g, rank)
    # the two receive buffers alternate from group to group; an overflowing first attempt is repeated into the same buffer
    parities = [p for p, _ in eng.partition_calls]
    assert parities[0] == 0 and parities[-1] == 0 and 1 in parities
    if first_cap < 100:
        assert ts.retries >= 1 and eng.allocs >= 2          # the regions (and the buffers) grew
        assert len(eng.partition_calls) > 3
    else:
        assert ts.retries == 0 and len(eng.partition_calls) == 3
    # from the second group on the regions are sized from the fullest region the team saw
    assert eng.partition_calls[-1][1] != first_cap
    assert ts.hints[(k, n_genomes, sum(-(-s // 64) for s in sizes))]["rho"] > 0
    ts.close()
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("k,world,n_genomes,first_cap", [(21, 2, 5, 1 << 20), (35, 3, 7, 16), (31, 2, 4, 16)])
def test_team_histograms_add_up_to_the_oracle(tmp_path, oracle, k, world, n_genomes, first_cap):
    groups = _groups(3, n_genomes)
    for g, texts in groups.items():
        w_ref, _, _ = oracle.exp1(texts, [0] * len(texts), 1, k, nbins=64)
        np.save(tmp_path / f"ref_{g}.npy", w_ref[0])
    mp.spawn(_worker, args=(world, _free_port(), k, n_genomes, first_cap, str(tmp_path)), nprocs=world, join=True)


def test_team_shapes():
    from khoice_b200.dist import chunk_layout, genome_slices, team_shape
    assert team_shape(20, 8) == 2 and team_shape(80, 8) == 1 and team_shape(1, 8) == 8
    assert team_shape(100, 8) == 1          # 13 / 12 whole groups per GPU (0.96) beat sharding every group
    assert team_shape(3, 4) == 4 and team_shape(10, 1) == 1 and team_shape(6, 4) == 2 and team_shape(3, 2) == 2 and team_shape(7, 8) == 1
    for n, t in ((100, 2), (200, 8), (5, 3), (64, 2), (130, 2)):
        sl = genome_slices(n, t)
        assert sl[0][0] == 0 and sl[-1][1] == n and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        assert max(h - l for l, h in sl) - min(h - l for l, h in sl) <= 1
    assert chunk_layout([50, 50]) == ([0, 1], 2)
    assert chunk_layout([65, 65]) == ([0, 2], 4)
    assert chunk_layout([64, 1, 130]) == ([0, 1, 2], 5)


# ---- the work-root driver with teams (pipeline.run_fused_distributed(team=T)) on the stand-in ------------------------------------
class _Packed:
    def __init__(self, texts, O):
        self.texts, self.O = list(texts), O

    def info(self):
        return {"n_symbols": sum(len(t) for t in self.texts)}

    def free(self):
        pass


def _pipeline_worker(rank, world, port, root, shared, ks, team):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from khoice_b200 import dist as kd, pipeline
    from test_dist_gloo import OracleAdapter
    kd.init_from_env("gloo")

    class Adapter(OracleAdapter):
        def __init__(self, k):
            super().__init__(k)
            self.eng = FakeTeamEngine(os.path.join(shared, f"team{rank // max(team, 1)}"), 1 << 20)
            self.eng.adapter = self

        def pack_group(self, files):
            return _Packed(files, self.O)

        def group_from_packed(self, packed, k, nbins):
            self.k = k
            return self.group(packed.texts, k, nbins)

    os.makedirs(os.path.join(shared, f"team{rank // max(team, 1)}"), exist_ok=True)
    ad = Adapter(int(ks[0]))
    rep = pipeline.run_fused_distributed(root, 3, ks, adapter=ad, exchange="nccl", team=team)
    assert rep["team_size"] == team and rep["world"] == world
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,team", [(2, 2), (4, 2), (3, 3)])
def test_work_root_driver_with_teams_writes_the_single_process_files(tmp_path, oracle, world, team):
    """3 groups of 4 genomes: one team of 2, two teams of 2 (groups dealt 2 / 1), one team of 3 -- the step_4 / step_8 histogram files
    are those of the oracle whatever the sharding."""
    from khoice_b200 import pipeline, synth, tables
    cfg = synth.SynthConfig(n_groups=3, genomes_per_group=4, genome_len=7_000, seed=5)
    root = str(tmp_path / "w")
    synth.write_dataset(cfg, root)
    ks = ["21", "35"]
    shared = str(tmp_path / "shared")
    os.makedirs(shared)
    mp.spawn(_pipeline_worker, args=(world, _free_port(), root, shared, ks, team), nprocs=world, join=True)
    flat, gid = [], []
    for g in range(1, 4):
        for i in range(1, 5):
            flat.append(synth.make_genome(cfg, g, i))
            gid.append(g - 1)
    for k in ks:
        w_ref, a_ref, _ = oracle.exp1(flat, gid, 3, int(k), nbins=tables.HIST_ROWS)
        for num in range(1, 4):
            got = tables.read_histogram_file(os.path.join(root, pipeline.p_step4(k, num)))
            assert got == [int(x) for x in w_ref[num - 1][1:]], (k, num)
        assert tables.read_histogram_file(os.path.join(root, pipeline.p_step8(k))) == [int(x) for x in a_ref[1:]], k
    assert os.path.exists(os.path.join(root, pipeline.P_STEP5)) and os.path.exists(os.path.join(root, pipeline.P_STEP9))
