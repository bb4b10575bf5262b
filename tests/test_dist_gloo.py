"""World-size-2 (and 3) CPU test of the multi-GPU driver (khoice_b200/dist.py) over gloo.

The collective driver is engine-agnostic; here it runs on a numpy/oracle stand-in adapter (TEST ONLY --
the product adapter is dist.CudaAdapter and needs a B200).  What is checked is the host-side logic: group
dealing, size exchange, variable all-to-all, histogram all-reduce, and that the result does not depend on
the number of ranks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleAdapter:
    """Stand-in for dist.CudaAdapter: same four methods, CPU oracle arithmetic, CPU tensors."""

    def __init__(self, k):
        from oracle import oracle as O
        self.O = O
        self.sets = []
        self.k = k

    def reset(self):
        self.sets = []

    def group(self, files, k, nbins):
        O = self.O
        keys, counts = O.union_sum([O.genome_set(f, k) for f in files], k)
        self.sets.append(keys)
        nsym = sum(O.kmers(f, k)[1] for f in files)
        st = {"bases": nsym, "windows": nsym, "genome_distinct": 0, "distinct": keys.shape[0], "ms_total": 0.0}
        return O.histogram(counts, nbins), st

    def export_partitions(self, k, world):
        w = 1 if k <= 32 else 2
        allk = np.concatenate(self.sets, axis=0) if self.sets else np.empty((0,) if w == 1 else (0, 2), np.uint64)
        flat = allk.reshape(-1, w)
        h = (flat[:, 0] * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)
        if w == 2:
            h = h ^ (flat[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F) >> np.uint64(40))
        dest = (h % np.uint64(world)).astype(np.int64)
        order = np.argsort(dest, kind="stable")
        words = [int((dest == r).sum()) * w for r in range(world)]
        return torch.from_numpy(flat[order].reshape(-1).astype(np.int64)), words

    def import_keys(self, recv, k, n_groups):
        w = 1 if k <= 32 else 2
        a = recv.numpy().astype(np.uint64)
        self.sets = [a if w == 1 else a.reshape(-1, 2)]

    def across(self, nbins):
        keys, counts = self.O.union_sum(self.sets, self.k)
        return self.O.histogram(counts, nbins), {"distinct": keys.shape[0], "ms_total": 0.0}

    def new_tensor(self, n):
        return torch.empty(n, dtype=torch.int64)


def _make_groups(n_groups, k):
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=n_groups, genomes_per_group=3, genome_len=12_000, seed=321)
    return {g: [synth.make_genome(cfg, g, i) for i in range(1, 4)] for g in range(1, n_groups + 1)}


def _worker(rank, world, port, n_groups, k, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from khoice_b200 import dist as kd
    r, w, _ = kd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    allg = _make_groups(n_groups, k)
    mine = {g: allg[g] for g in kd.groups_of_rank(n_groups, rank, world)}
    within, across, stats = kd.run_exp1_k(OracleAdapter(k), mine, n_groups, k)
    np.save(os.path.join(out_dir, f"within_{world}_{rank}.npy"), within)
    np.save(os.path.join(out_dir, f"across_{world}_{rank}.npy"), across)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("k", [21, 35])
def test_rank_count_does_not_change_the_histograms(tmp_path, oracle, k):
    n_groups = 5
    allg = _make_groups(n_groups, k)
    flat = [f for g in sorted(allg) for f in allg[g]]
    gid = [g - 1 for g in sorted(allg) for _ in allg[g]]
    w_ref, a_ref, _ = oracle.exp1(flat, gid, n_groups, k)
    for world in (1, 2, 3):
        mp.spawn(_worker, args=(world, _free_port(), n_groups, k, str(tmp_path)), nprocs=world, join=True)
        for rank in range(world):
            assert np.array_equal(np.load(tmp_path / f"within_{world}_{rank}.npy"), w_ref), (world, rank)
            assert np.array_equal(np.load(tmp_path / f"across_{world}_{rank}.npy"), a_ref), (world, rank)


def test_group_dealing():
    from khoice_b200.dist import groups_of_rank
    for n, world in ((10, 1), (10, 4), (3, 8), (100, 8)):
        owned = [groups_of_rank(n, r, world) for r in range(world)]
        assert sorted(g for o in owned for g in o) == list(range(1, n + 1))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


# ---- the peer-memory exchange protocol (dist.AcrossExchanger) on a stand-in engine ------------------------------------
class FakePeerEngine:
    """TEST ONLY.  The calls of csrc/peer.cu with "peer memory" = numpy memmap files in a directory all ranks see.  What is
    checked on CPU is the host protocol: sizing round over the all-to-all route, regions, count table, overflow -> fallback ->
    regrow, and that the histograms do not depend on the route or the number of ranks."""

    def __init__(self, adapter, shared_dir):
        self.ad, self.dir = adapter, shared_dir
        self.gen = 0
        self.maps = None
        self.group_sets_hashed = False

    def _path(self, rank, gen):
        return os.path.join(self.dir, f"recv_{rank}_{gen}.bin")

    def peer_alloc(self, world, rank, key_bytes, region_keys):
        self.gen += 1
        self.world, self.rank, self.w, self.region = world, rank, key_bytes // 8, int(region_keys)
        np.lib.format.open_memmap(self._path(rank, self.gen), mode="w+", dtype=np.uint64, shape=(world, self.region, self.w))
        return self.gen.to_bytes(8, "little") + bytes(56)

    def peer_open(self, handles):
        gens = [int.from_bytes(handles[64 * r:64 * r + 8], "little") for r in range(self.world)]
        self.maps = [np.load(self._path(r, gens[r]), mmap_mode="r+") for r in range(self.world)]

    def peer_begin(self):
        self.cursor = np.zeros(self.world, dtype=np.uint64)
        self.ovf = False
        self.pushed = 0

    def peer_push(self):
        new = self.ad.sets[self.pushed:]
        self.pushed = len(self.ad.sets)
        for keys in new:
            flat = keys.reshape(-1, self.w)
            dest = self.ad.owner(flat, self.world)
            for r in range(self.world):
                part = flat[dest == r]
                c = int(self.cursor[r])
                room = max(min(self.region - c, part.shape[0]), 0)
                self.maps[r][self.rank, c:c + room] = part[:room]
                self.ovf |= room < part.shape[0]
                self.cursor[r] += np.uint64(part.shape[0])

    def peer_counts(self, world):
        for m in self.maps:
            m.flush()
        return self.cursor.copy(), self.ovf

    def peer_import(self, recv_counts, k, n_groups, hashed):
        mine = np.load(self._path(self.rank, self.gen), mmap_mode="r")
        got = np.concatenate([np.array(mine[s, :int(recv_counts[s])]) for s in range(self.world)], axis=0)
        self.ad.sets = [got.reshape(-1) if self.w == 1 else got]

    def peer_unmap(self):
        self.maps = None

    def peer_close(self):
        pass

    @property
    def peer_region_keys(self):
        return self.region


class PeerOracleAdapter(OracleAdapter):
    def __init__(self, k, shared_dir):
        super().__init__(k)
        self.eng = FakePeerEngine(self, shared_dir)

    @staticmethod
    def owner(flat, world):
        h = (flat[:, 0] * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)
        if flat.shape[1] == 2:
            h = h ^ (flat[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F) >> np.uint64(40))
        return (h % np.uint64(world)).astype(np.int64)


def _peer_worker(rank, world, port, n_groups, k, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from khoice_b200 import dist as kd
    kd.init_from_env("gloo")
    allg = _make_groups(n_groups, k)
    mine = kd.groups_of_rank(n_groups, rank, world)
    ref = np.load(os.path.join(out_dir, "across_ref.npy"))
    for region, expect in ((None, ["nccl", "peer", "peer"]), (8, ["nccl", "peer", "peer"]), (1 << 20, ["peer", "peer", "peer"])):
        ad = PeerOracleAdapter(k, out_dir)
        ex = kd.AcrossExchanger(ad, k, n_groups, nbins=64, mode="peer", region_keys=region)
        for rnd in range(3):
            ad.reset()
            ex.begin()
            for g in mine:
                ad.group(allg[g], k, 64)
                ex.after_group()
            hist, info = ex.finish()
            assert info["exchange"] == expect[rnd], (region, rnd, info["exchange"])
            assert np.array_equal(hist, ref), (region, rnd, rank)
        if region == 8:
            assert ad.eng.peer_region_keys > 8          # the overflowing round made the regions grow
        ex.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k,world", [(21, 2), (35, 3)])
def test_peer_exchange_protocol_on_stand_in_engine(tmp_path, oracle, k, world):
    n_groups = 5
    allg = _make_groups(n_groups, k)
    flat = [f for g in sorted(allg) for f in allg[g]]
    gid = [g - 1 for g in sorted(allg) for _ in allg[g]]
    _, a_ref, _ = oracle.exp1(flat, gid, n_groups, k, nbins=64)
    np.save(tmp_path / "across_ref.npy", a_ref)
    mp.spawn(_peer_worker, args=(world, _free_port(), n_groups, k, str(tmp_path)), nprocs=world, join=True)


# ---- the distributed fused driver of the work-root pipeline (pipeline.run_fused_distributed) ---------------------------
class _PackedTexts:
    def __init__(self, texts):
        self.texts = list(texts)

    def free(self):
        pass


class PipelineOracleAdapter(PeerOracleAdapter):
    """+ the two calls the work-root driver makes per group (pack once, count per k)."""

    def pack_group(self, files):
        return _PackedTexts(files)

    def group_from_packed(self, packed, k, nbins):
        self.k = k
        return self.group(packed.texts, k, nbins)


def _pipeline_worker(rank, world, port, root, shared, ks, exchange):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from khoice_b200 import dist as kd, pipeline
    kd.init_from_env("gloo")
    rep = pipeline.run_fused_distributed(root, 5, ks, adapter=PipelineOracleAdapter(int(ks[0]), shared), exchange=exchange)
    assert rep["world"] == world
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,exchange", [(1, "peer"), (2, "peer"), (3, "nccl")])
def test_distributed_work_root_driver_writes_the_single_process_files(tmp_path, oracle, world, exchange):
    """step_4 / step_8 histogram files and the step_5 / step_9 CSVs do not depend on the number of ranks or the route."""
    from khoice_b200 import pipeline, synth, tables
    cfg = synth.SynthConfig(n_groups=5, genomes_per_group=3, genome_len=9_000, seed=321)
    root = str(tmp_path / "w")
    synth.write_dataset(cfg, root)
    ks = ["11", "21", "25"]
    shared = str(tmp_path / "shared")
    os.makedirs(shared)
    mp.spawn(_pipeline_worker, args=(world, _free_port(), root, shared, ks, exchange), nprocs=world, join=True)
    flat, gid = [], []
    for g in range(1, 6):
        for i in range(1, 4):
            flat.append(synth.make_genome(cfg, g, i))
            gid.append(g - 1)
    for k in ks:
        w_ref, a_ref, _ = oracle.exp1(flat, gid, 5, int(k), nbins=tables.HIST_ROWS)
        for num in range(1, 6):
            got = tables.read_histogram_file(os.path.join(root, pipeline.p_step4(k, num)))
            assert got == [int(x) for x in w_ref[num - 1][1:]], (k, num)
        assert tables.read_histogram_file(os.path.join(root, pipeline.p_step8(k))) == [int(x) for x in a_ref[1:]], k
    # the CSVs are those of the table builders over exactly these histogram files
    ref_root = str(tmp_path / "ref")
    synth.write_dataset(cfg, ref_root)
    for k in ks:
        for num in range(1, 6):
            os.makedirs(os.path.dirname(os.path.join(ref_root, pipeline.p_step4(k, num))), exist_ok=True)
            os.replace(os.path.join(root, pipeline.p_step4(k, num)), os.path.join(ref_root, pipeline.p_step4(k, num)))
        os.makedirs(os.path.dirname(os.path.join(ref_root, pipeline.p_step8(k))), exist_ok=True)
        os.replace(os.path.join(root, pipeline.p_step8(k)), os.path.join(ref_root, pipeline.p_step8(k)))
    pipeline.build_tables(ref_root, ks, 5)
    for f in (pipeline.P_STEP5, pipeline.P_STEP9) + pipeline.P_FINAL:
        assert open(os.path.join(root, f), "rb").read() == open(os.path.join(ref_root, f), "rb").read(), f
    assert os.path.exists(os.path.join(root, pipeline.p_step7(ks[0]) + ".kmc_pre"))
