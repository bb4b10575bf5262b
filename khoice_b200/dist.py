"""Multi-GPU execution of experiment type 1: one process per GPU, `torch.distributed` for the plumbing.

How the path shards:

* steps 1-6 (per genome / per group; /root/reference/workflow/rules/exp_type_1.smk:156-241) are independent per
  group -> whole groups are dealt to ranks round-robin, NO data-path collective;
* steps 7-8 (across groups; exp_type_1.smk:243-259) need every copy of a k-mer on one GPU -> the k-mer space is
  hash-range partitioned (K7, ``khb_partition_by_hash``): ONE variable-size all-to-all of the retained group
  sets over NVLink (``all_to_all_single`` on NCCL, sizes exchanged first), then a local sort + run-length
  histogram, then an all-reduce(sum) of the <= 5001-bin histogram.  The per-group histograms are
  all-reduced into a [G, 5001] table so that every rank can write identical step_4 / step_8 files.

The reference has no distributed code at all (its parallelism is Snakemake running independent rules as
processes); the invariant to keep is that the histograms -- hence the CSVs -- do not depend on the GPU
count.

The collective driver below is engine-agnostic: it talks to an *adapter* with four methods (group / export
partitions / import keys / across).  ``CudaAdapter`` is the product adapter (device buffers, NCCL); the CPU
test suite drives the same driver over gloo with a numpy stand-in defined in tests/ (never shipped).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .engine import COUNTER_MAX, Engine, key_words


def groups_of_rank(n_groups: int, rank: int, world: int) -> List[int]:
    """1-based group numbers owned by `rank` (round-robin keeps big and small groups mixed)."""
    return [g for g in range(1, n_groups + 1) if (g - 1) % world == rank]


class CudaAdapter:
    """Product adapter: everything stays in HBM; torch only owns the exchange buffers."""

    def __init__(self, engine: Engine, device: torch.device):
        self.eng = engine
        self.device = device

    def group(self, files: Sequence, k: int, nbins: int):
        return self.eng.group_from_fasta(files, k, nbins=nbins, keep_set=True)

    def pack_group(self, files: Sequence):
        """K1 once per group; the 2-bit stream stays in HBM for the k sweep (khb_pack_group)."""
        return self.eng.pack_group(files)

    def group_from_packed(self, packed, k: int, nbins: int):
        return self.eng.group_from_packed(packed, k, nbins=nbins, keep_set=True)

    def export_partitions(self, k: int, world: int) -> Tuple[torch.Tensor, List[int]]:
        """Retained group sets, grouped by destination rank.  Returns (int64 tensor of words, words per rank)."""
        w = key_words(k)
        self._hashed = self.eng.group_sets_hashed
        ptr, n = self.eng.group_sets_device()
        send = torch.empty(max(n, 1) * w, dtype=torch.int64, device=self.device)
        off = np.zeros(world + 1, dtype=np.uint64)
        if n:
            self.eng._chk(self.eng.lib.khb_partition_by_hash(self.eng.ctx, ptr, n, k, world, send.data_ptr(), off.ctypes.data))
            self.eng.sync()
        return send[: n * w], [int(off[i + 1] - off[i]) * w for i in range(world)]

    def import_keys(self, recv: torch.Tensor, k: int, n_groups: int) -> None:
        """Replace the retained sets by the received keys of this rank's hash range."""
        self.eng.group_sets_reset()
        torch.cuda.synchronize(self.device)
        self.eng.group_sets_append_device(recv.data_ptr() if recv.numel() else 0, recv.numel() // key_words(k), k, n_groups,
                                          hashed=getattr(self, "_hashed", False))
        self.eng.sync()

    def across(self, nbins: int):
        return self.eng.across_groups(nbins=nbins)

    def reset(self):
        self.eng.group_sets_reset()

    def new_tensor(self, n: int) -> torch.Tensor:
        return torch.empty(n, dtype=torch.int64, device=self.device)


def exchange_and_count(adapter, k: int, n_groups_total: int, nbins: int = COUNTER_MAX, group=None):
    """Across-group stage on `world` ranks: hash partition -> all-to-all -> local count -> all-reduce.
    Returns (step_8 histogram int64[nbins+1] identical on every rank, dict of exchange sizes)."""
    world = dist.get_world_size(group)
    send, send_words = adapter.export_partitions(k, world)
    dev = send.device
    # NCCL moves device tensors; any other backend (gloo: CPU tests, two processes sharing one GPU) goes through host copies
    wire = dev if dist.get_backend(group) == "nccl" else torch.device("cpu")
    # sizes first (one tiny all-to-all), then the payload
    s = torch.tensor(send_words, dtype=torch.int64, device=wire)
    r = torch.empty(world, dtype=torch.int64, device=wire)
    dist.all_to_all_single(r, s, group=group)
    recv_words = [int(x) for x in r.tolist()]
    if wire == dev:
        recv = adapter.new_tensor(sum(recv_words))
        dist.all_to_all_single(recv, send, output_split_sizes=recv_words, input_split_sizes=send_words, group=group)
    else:
        recv_w = torch.empty(sum(recv_words), dtype=torch.int64)
        dist.all_to_all_single(recv_w, send.to(wire), output_split_sizes=recv_words, input_split_sizes=send_words, group=group)
        recv = recv_w.to(dev)
    adapter.import_keys(recv, k, n_groups_total)
    hist, st = adapter.across(nbins)
    h = torch.from_numpy(np.ascontiguousarray(hist).astype(np.int64)).to(wire)
    dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
    info = {"send_words": sum(send_words), "recv_words": sum(recv_words), "local_distinct": int(st.get("distinct", 0)),
            "across_ms": float(st.get("ms_total", 0.0)), "send_words_per_rank": list(send_words), "exchange": "nccl"}
    return h.cpu().numpy().astype(np.uint64), info


class AcrossExchanger:
    """The across-group stage of one rank, round after round (one round = one k over all groups).

    ``mode="nccl"``: partition + all-to-all + copy (exchange_and_count).  ``mode="peer"`` (CUDA adapters only): every group's
    new keys are stored straight into their owner's receive buffer by one kernel behind the group's K5 (csrc/peer.cu, CUDA
    IPC over NVLink), the only collective on the path is the table of counts.  The first round always runs over NCCL: it
    measures how many keys every (sender, owner) pair exchanges, which sizes the regions (x 1.3, agreed with an
    all-reduce(max)); a round in which a region still overflows is redone over NCCL and the regions grow.

    Usage per round:  begin();  after every group_from_*(keep_set=True): after_group();  finish() -> (histogram, info)."""

    def __init__(self, adapter, k: int, n_groups_total: int, nbins: int = COUNTER_MAX, mode: str = "peer", group=None,
                 region_keys: Optional[int] = None):
        self.ad, self.k, self.n_groups, self.nbins, self.group = adapter, k, n_groups_total, nbins, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.mode = mode if self.world > 1 and hasattr(adapter, "eng") else "nccl"
        self.ready = False
        self.region_keys = 0
        self.rounds_peer = self.rounds_nccl = 0
        self.ctrl = adapter.new_tensor(0).device if dist.is_initialized() and dist.get_backend(group) == "nccl" else torch.device("cpu")
        if self.mode == "peer" and region_keys:
            self._setup(int(region_keys))

    def _setup(self, region_keys: int):
        """Collective.  (Re)allocate the receive buffers and map everybody's.  If any rank cannot (no IPC between the
        processes, out of memory), ALL ranks drop to the NCCL route for the rest of the run."""
        import sys
        from .engine import KhbError
        eng = self.ad.eng
        if self.ready:
            self.close()
        self.region_keys = int(region_keys)
        ok, why, handle = 1, "", bytes(64)
        try:
            handle = eng.peer_alloc(self.world, self.rank, 8 * key_words(self.k), region_keys)
        except KhbError as e:
            ok, why = 0, str(e)
        mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(self.ctrl)
        allh = torch.empty(64 * self.world, dtype=torch.uint8, device=self.ctrl)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        flag = torch.tensor([ok], dtype=torch.int64, device=self.ctrl)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()):
            try:
                eng.peer_open(bytes(allh.cpu().numpy().tobytes()))
            except KhbError as e:
                ok, why = 0, str(e)
            flag = torch.tensor([ok], dtype=torch.int64, device=self.ctrl)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)   # also: nobody pushes before everybody has mapped everybody
        if not int(flag.item()):
            if why:
                print(f"khoice_b200.dist: rank {self.rank}: peer-memory exchange unavailable ({why}); using NCCL", file=sys.stderr)
            eng.peer_unmap()
            dist.barrier(group=self.group)
            eng.peer_close()
            self.mode = "nccl"
            return
        self.ready = True

    def set_k(self, k: int):
        """Reuse the exchanger (and its mapped regions) for another k of the same key width (a k sweep)."""
        if key_words(k) != key_words(self.k):
            raise ValueError(f"exchanger set up for {8 * key_words(self.k)}-byte keys, k={k} needs {8 * key_words(k)}")
        self.k = k

    def begin(self):
        if self.ready:
            self.ad.eng.peer_begin()

    def after_group(self):
        if self.ready:
            self.ad.eng.peer_push()

    def _grow_from(self, per_dest_keys: Sequence[int]):
        need = torch.tensor([max(per_dest_keys) if len(per_dest_keys) else 0], dtype=torch.int64, device=self.ctrl)
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)
        # 1.3 x what this round needed -- and, when regions that existed already were too small (a k sweep: the distinct k-mers grow
        # ~4 x per step at small k), at least twice the old size, so that a sweep re-maps the buffers O(log) times, not once per k
        self._setup(max(int(need.item()) * 13 // 10, 2 * self.region_keys, 1024))

    def finish(self):
        if self.world == 1:
            return self.ad.across(self.nbins)
        if not self.ready:
            hist, info = exchange_and_count(self.ad, self.k, self.n_groups, self.nbins, self.group)
            self.rounds_nccl += 1
            if self.mode == "peer":
                w = key_words(self.k)
                self._grow_from([x // w for x in info["send_words_per_rank"]])
            return hist, info
        eng = self.ad.eng
        hashed = 0 if self.k in (32, 64) else 1      # a property of k (csrc/api.cu: count_stage), the same on every rank -- also on one without groups
        counts, ovf = eng.peer_counts(self.world)
        row = torch.from_numpy(np.concatenate([counts.astype(np.int64), [1 if ovf else 0]])).to(self.ctrl)
        table = torch.empty((self.world, self.world + 1), dtype=torch.int64, device=self.ctrl)
        dist.all_gather_into_tensor(table.view(-1), row, group=self.group)   # also the barrier: every rank's pushes are done
        table = table.cpu().numpy()
        if table[:, self.world].any():
            hist, info = exchange_and_count(self.ad, self.k, self.n_groups, self.nbins, self.group)  # the local store is intact
            self.rounds_nccl += 1
            self._grow_from([int(x) for x in table[self.rank, : self.world]])
            return hist, info
        if hasattr(eng, "peer_across") and os.environ.get("KHB_PEER_GATHER", "1") != "0":
            # the pushed regions are sorted where they lie: no import copy (KHB_PEER_GATHER=0 keeps import + across)
            hist, st = eng.peer_across(table[:, self.rank].astype(np.uint64), self.k, self.n_groups, hashed, nbins=self.nbins)
        else:
            eng.peer_import(table[:, self.rank].astype(np.uint64), self.k, self.n_groups, hashed)
            hist, st = self.ad.across(self.nbins)
        dev = self.ad.new_tensor(0).device
        h = torch.from_numpy(np.ascontiguousarray(hist).astype(np.int64)).to(dev if self.ctrl.type != "cpu" else self.ctrl)
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=self.group)           # also orders the next round's pushes behind this import
        self.rounds_peer += 1
        info = {"send_words": int(counts.sum()) * key_words(self.k), "recv_words": int(table[:, self.rank].sum()) * key_words(self.k),
                "local_distinct": int(st.get("distinct", 0)), "across_ms": float(st.get("ms_total", 0.0)), "exchange": "peer"}
        return h.cpu().numpy().astype(np.uint64), info

    def close(self):
        """Collective: every rank drops its mappings before any rank frees its buffer."""
        if self.ready:
            self.ad.eng.peer_unmap()
            dist.barrier(group=self.group)
            self.ad.eng.peer_close()
            self.ready = False


# ---- ONE group on several GPUs ---------------------------------------------------------------------------------------------
def team_shape(n_groups: int, world: int, max_team: int = 8, min_whole_efficiency: float = 0.85) -> int:
    """Members per team.  Whole groups (T = 1) wherever dealing them keeps the GPUs busy: n_groups / (ceil(n_groups / world) * world)
    >= min_whole_efficiency -- sharding a group costs about a tenth (the region transfer, the barrier, measured on 2 B200s), so a
    13 / 12 deal of 100 groups over 8 GPUs (0.96) stays whole.  Otherwise the smallest divisor T of `world` for which the groups divide
    evenly over the world / T teams: 20 groups on 8 GPUs (3/3/3/3/2/2/2/2 = 0.83) -> T = 2, four teams of five groups; 3 groups on
    2 GPUs -> T = 2; one group -> every GPU on it."""
    if world <= 1 or n_groups < 1:
        return 1
    whole = n_groups / (-(-n_groups // world) * world)
    if whole >= min_whole_efficiency:
        return 1
    for t in range(2, min(world, max_team) + 1):
        if world % t == 0 and n_groups % (world // t) == 0:
            return t
    return 1


def genome_slices(n_genomes: int, team_size: int) -> List[Tuple[int, int]]:
    """[lo, hi) of every member's contiguous slice of a group's genomes (sizes differ by at most one)."""
    q, r = divmod(n_genomes, team_size)
    out, lo = [], 0
    for t in range(team_size):
        hi = lo + q + (1 if t < r else 0)
        out.append((lo, hi))
        lo = hi
    return out


def chunk_layout(slice_sizes: Sequence[int]) -> Tuple[List[int], int]:
    """Every slice is padded to whole chunks of 64 genome ids, so that a (bin, chunk) region of the record buffers has exactly one
    writer: (first chunk of every member, chunks of the whole group)."""
    base, c = [], 0
    for n in slice_sizes:
        base.append(c)
        c += -(-int(n) // 64)
    return base, c


class TeamSharder:
    """Steps 1-4 of ONE group (exp_type_1.smk:156-191) on the GPUs of a team (include/khoice_b200.h: khb_team_*).

    Every member holds a slice of the group's genomes.  Per group and member: K1 + the partition pass of the minimizer-bin stage,
    whose super-k-mer records go straight into the record buffer of the bin's owner (peer memory, CUDA IPC over NVLink); ONE small
    all-gather in the team -- the barrier behind which every member's records are at their owners, and the carrier of the overflow
    flags, the fullest region (sizes the next group's regions) and the distinct-k-mer ratio of the previous group (sizes its
    tables), so that all members keep planning from the same numbers --; then every member counts the bins it owns.  The partial
    histograms add up to the group's step_4 histogram (the caller's all-reduce over all ranks does that), the distinct keys of
    the member's bins enter its group-set store / the across-group exchange like those of a whole group.

    `group`: the process group of the team (None: the default group, when the world is one team)."""

    def __init__(self, eng: Engine, team_size: int, member: int, group=None):
        self.eng, self.T, self.member, self.group = eng, int(team_size), int(member), group
        self.ready = False
        self.half_bytes = 0
        self.parity = 0
        self.hints: Dict[Tuple[int, int, int], dict] = {}     # (k, genomes, chunks) -> {"rho", "cap"} agreed in the team
        self.prev = (0, 0, (0, 0, 0))                         # distinct k-mers / windows this member counted in the previous group, its shape
        self.retries = self.setups = self.groups = 0
        nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"
        self.ctrl = torch.device("cuda", eng.device) if nccl else torch.device("cpu")

    def _setup(self, half_bytes: int):
        """Collective in the team: (re)allocate the receive buffers and map everybody's."""
        if self.ready:
            self.close()
        handle = self.eng.team_alloc(self.T, self.member, int(half_bytes))
        mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(self.ctrl)
        allh = torch.empty(64 * self.T, dtype=torch.uint8, device=self.ctrl)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        self.eng.team_open(bytes(allh.cpu().numpy().tobytes()))
        dist.barrier(group=self.group)          # nobody stores before everybody has mapped everybody
        self.half_bytes = int(half_bytes)
        self.ready = True
        self.setups += 1

    def close(self):
        """Collective in the team: every member drops its mappings before any member frees its buffers."""
        if self.ready:
            self.eng.team_unmap()
            dist.barrier(group=self.group)
            self.eng.team_close()
            self.ready = False

    def _fit(self, k: int, tg):
        plan = self.eng.team_plan(k, tg) if self.ready else None
        if plan is None:
            # the plan needs a team object for its size: allocate a token first (every member takes the same path)
            self._setup(1 << 20)
            plan = self.eng.team_plan(k, tg)
        if plan["half_bytes"] > self.half_bytes:
            self._setup(plan["half_bytes"] + plan["half_bytes"] // 4)
        return plan

    def run_group(self, source, k: int, n_genomes_total: int, slice_sizes: Sequence[int], n_sym_total: int, nbins: int = COUNTER_MAX,
                  keep_set: bool = True):
        """`source`: this member's slice (host FASTA texts, StagedFasta or PackedGroup).  `slice_sizes`: genomes of every member's slice;
        `n_sym_total`: symbols of the whole group -- the same estimate on every member.  Returns (partial histogram, stats)."""
        from .engine import TeamGroup
        if len(slice_sizes) != self.T or sum(slice_sizes) != n_genomes_total or min(slice_sizes) < 1:
            raise ValueError(f"a team of {self.T} needs {self.T} non-empty slices of the group's {n_genomes_total} genomes, got {list(slice_sizes)}")
        base, n_chunks = chunk_layout(slice_sizes)
        shape = (int(k), int(n_genomes_total), int(n_chunks))
        hint = self.hints.get(shape, {})
        tg = TeamGroup(n_genomes_total, n_chunks, base[self.member], self.parity, int(n_sym_total), float(hint.get("rho", 0.0)),
                       int(hint.get("cap", 0)), int(hint.get("area_pct", 0)))
        self._fit(k, tg)
        for attempt in range(8):
            info = self.eng.team_partition(source, k, tg)
            flags = (1 if info["overflow"] else 0) | (2 if info.get("area_overflow") else 0)
            row = torch.tensor([flags, info["fullest_region"], self.prev[0], self.prev[1]], dtype=torch.int64, device=self.ctrl)
            table = torch.empty((self.T, 4), dtype=torch.int64, device=self.ctrl)
            dist.all_gather_into_tensor(table.view(-1), row, group=self.group)     # also the barrier: every member's records are at their owners
            table = table.cpu().numpy()
            fullest = int(table[:, 1].max())
            flags = int(np.bitwise_or.reduce(table[:, 0]))
            if not flags:
                break
            # something was too small somewhere: every member partitions again -- into larger local regions (the senders count what was
            # asked for, so the fullest region is known exactly) and / or into larger areas of the owners' buffers
            self.retries += 1
            if flags & 1:
                tg.region_cap = max(fullest + fullest // 4, 2 * int(self.eng.team_plan(k, tg)["region_cap"])) + 64
            if flags & 2:
                tg.area_pct = 2 * (int(tg.area_pct) or 250)
                hint = dict(hint, area_pct=int(tg.area_pct))
            self._fit(k, tg)
        else:
            raise RuntimeError(f"team: the regions of a k={k} group of {n_genomes_total} genomes still overflow at {tg.region_cap} records")
        hist, st = self.eng.team_count(k, tg, nbins=nbins, keep_set=keep_set)
        # what the team agreed on before this count: the previous group's ratio (if it had this shape) and this group's fullest region
        if self.prev[2] == shape and int(table[:, 3].sum()) > 0:
            hint = dict(hint, rho=float(table[:, 2].sum()) / float(table[:, 3].sum()))
        hint = dict(hint, cap=int(fullest * 1.4) + 64)
        self.hints[shape] = hint
        self.prev = (int(st["distinct"]), int(info["windows"]), shape)
        self.parity ^= 1
        self.groups += 1
        return hist, st


def run_exp1_k(adapter, groups: Dict[int, Sequence], n_groups_total: int, k: int, nbins: int = COUNTER_MAX, group=None):
    """One k on this rank's share.  `groups` maps the 1-based numbers of the groups this rank owns to their
    FASTA texts.  Returns (within [n_groups_total, nbins+1] uint64, across [nbins+1] uint64, stats)."""
    adapter.reset()
    within = np.zeros((n_groups_total, nbins + 1), dtype=np.int64)
    stats = {"bases": 0, "windows": 0, "genome_distinct": 0, "group_distinct": 0, "group_ms": 0.0}
    for num in sorted(groups):
        hist, st = adapter.group(groups[num], k, nbins)
        within[num - 1] = hist.astype(np.int64)
        stats["bases"] += int(st["bases"])
        stats["windows"] += int(st["windows"])
        stats["genome_distinct"] += int(st["genome_distinct"])
        stats["group_distinct"] += int(st["distinct"])
        stats["group_ms"] += float(st["ms_total"])
    across, info = exchange_and_count(adapter, k, n_groups_total, nbins, group)
    stats.update(info)
    dev = adapter.new_tensor(0).device
    w = torch.from_numpy(within).to(dev)
    dist.all_reduce(w, op=dist.ReduceOp.SUM, group=group)
    return w.cpu().numpy().astype(np.uint64), across, stats


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the default process group."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local
