"""Host-side logic of the experiment type 2 / type 4 drivers without a GPU: khoice_b200.pipeline2 / pipeline4 run on a
stand-in engine whose arithmetic is the CPU oracle (TEST ONLY -- the product engine is khoice_b200.engine.Engine and needs
a B200).  Checked: file layout, histogram splitting, parse-time files, CSV / matrix bytes against the golden fixtures
made with the reference's own code, resume-free reruns."""
import json
import os
import sys

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)


class _Packed:
    def __init__(self, texts):
        self.texts = list(texts)

    def free(self):
        pass


class _Buf:
    def __init__(self, keys):
        self.keys = keys

    def free(self):
        pass


class OracleEngine:
    """The methods pipeline2 / pipeline4 call on Engine, computed with the CPU oracle."""

    def __init__(self):
        from oracle import oracle as O
        self.O = O
        self.group_sets_reset()

    def group_sets_reset(self):
        self.sets, self.pivots = [], []

    def pack_group(self, texts):
        return _Packed(texts)

    def _union(self, texts, k):
        return self.O.union_sum([self.O.genome_set(t, k) for t in texts], k)

    def group_from_packed(self, pk, k, nbins=5000, keep_set=True):
        keys, counts = self._union(pk.texts, k)
        if keep_set:
            self.sets.append(keys)
        return self.O.histogram(counts, nbins), {"distinct": int(keys.shape[0])}

    def group_sets_info(self):
        return {"n_keys": int(sum(s.shape[0] for s in self.sets))}

    def pivot_group_from_packed(self, pk, k, nbins=5000, keep_sets=True):
        O = self.O
        rest, pivot = pk.texts[:-1], pk.texts[-1]
        ukeys, ucnt = self._union(rest, k)
        pset = O.genome_set(pivot, k)
        ones = np.ones(pset.shape[0], np.uint32)
        _, ci = O.simple_intersect_ocsum(pset, ones, ukeys, ucnt, k)
        _, cs = O.simple_kmers_subtract(pset, ones, ukeys, k)
        if keep_sets:
            self.sets.append(ukeys)
            self.pivots.append(pset)
        return O.histogram(ci, nbins) + O.histogram(cs, nbins), {"distinct": int(ukeys.shape[0])}

    def pivot_across(self, nbins=5000):
        O = self.O
        G = len(self.pivots)
        k = self.k_hint
        out = np.zeros((G, nbins + 1), dtype=np.uint64)
        for d in range(G):
            okeys, ocnt = O.union_sum([self.sets[i] for i in range(G) if i != d], k)
            ones = np.ones(self.pivots[d].shape[0], np.uint32)
            _, ci = O.simple_intersect_ocsum(self.pivots[d], ones, okeys, ocnt, k)
            _, cs = O.simple_kmers_subtract(self.pivots[d], ones, okeys, k)
            out[d] = O.histogram(ci, nbins) + O.histogram(cs, nbins)
        return out, {}

    def kmer_counts(self, text, k, cs=255):
        keys, counts = self.O.kmer_counts(text, k, cs)
        return _Buf(keys), counts, int(counts.shape[0])

    def group_membership(self, group_off, bufs, sizes, k):
        O = self.O
        assert list(group_off) == [0] + list(np.cumsum([s.shape[0] for s in self.sets]))
        rows = []
        for b in bufs:
            ids = O._row_ids([b.keys] + self.sets, k)
            m = np.zeros((b.keys.shape[0], 1), dtype=np.uint64)
            for d in range(len(self.sets)):
                m[np.isin(ids[0], ids[1 + d]), 0] |= np.uint64(1 << d)
            rows.append(m)
        return np.concatenate(rows, axis=0)

    def close(self):
        pass


def test_pipeline2_fused_on_oracle_engine(oracle, tmp_path):
    from khoice_b200 import pipeline2, synth, tables
    cfg = synth.SynthConfig(n_groups=3, genomes_per_group=4, genome_len=6_000, seed=77)
    root = str(tmp_path / "w")
    synth.write_dataset_type2(cfg, root)
    ks = ["9", "21", "34"]
    eng = OracleEngine()
    # pivot_across has no k argument in the engine API (the store knows it); give the stand-in the hint per k
    orig = pipeline2.run_fused

    class PerK(OracleEngine):
        def pivot_group_from_packed(self, pk, k, nbins=5000, keep_sets=True):
            self.k_hint = k
            return super().pivot_group_from_packed(pk, k, nbins, keep_sets)

    rep = orig(root, cfg.n_groups, ks, engine=PerK())
    assert rep["exp_type"] == 2 and len(rep["stages"]) == len(ks) * (cfg.n_groups + 1)
    groups = [[synth.make_genome(cfg, g, i) for i in range(1, 4)] for g in (1, 2, 3)]
    pivots = [synth.make_genome(cfg, g, 4) for g in (1, 2, 3)]
    for k in ks:
        w_ref, a_ref = oracle.exp2(groups, pivots, int(k), nbins=tables.HIST_ROWS)
        for num in (1, 2, 3):
            for scope, ref, fn in (("within", w_ref, pipeline2.p_within), ("across", a_ref, pipeline2.p_across)):
                for j, op in enumerate(pipeline2.OPS):
                    got = tables.read_histogram_file(os.path.join(root, fn(k, num, op) + ".hist.txt"))
                    assert got == [int(x) for x in ref[num - 1][j][1:]], (scope, k, num, op)
    # every output the reference's rules declare exists; the CSVs have one row per (dataset, k)
    for rule, outputs, _ in pipeline2._rule_jobs(root, ks, cfg.n_groups):
        for o in outputs:
            assert os.path.exists(os.path.join(root, o)), (rule, o)
    for f in (pipeline2.P_WITHIN_CSV, pipeline2.P_ACROSS_CSV):
        lines = open(os.path.join(root, f)).read().splitlines()
        assert len(lines) == 1 + cfg.n_groups * len(ks) and lines[1].startswith("group_1,9,")
    ops = open(os.path.join(root, pipeline2.p_ops_across("21", 2))).read()
    assert "dataset_2/dataset_2.transformed" not in ops.split("OUTPUT:")[0] and "dataset_1/" in ops and "dataset_3/" in ops


def test_pipeline4_fused_on_oracle_engine_matches_reference_merge_lists(tmp_path):
    from khoice_b200 import pipeline4, synth
    import make_golden_exp4 as G4
    cases = json.load(open(os.path.join(GOLDEN, "exp4_cases.json")))["cases"]
    for c, case in enumerate(cases):
        cfg, _, _ = G4.inputs_of(case)
        root = str(tmp_path / f"case{c}")
        synth.write_dataset_type4(cfg, root, out_pivot=case["out_pivot"])
        ks = [str(k) for k in case["k_values"]]
        pipeline4.run_fused(root, case["n_groups"], ks, engine=OracleEngine())
        for k in ks:
            for ours, gold in ((f"accuracies_type_4/confusion_matrix/k_{k}_confusion_matrix.txt", "confusion_matrix.txt"),
                               (f"accuracies_type_4/confusion_matrix/k_{k}_confusion_matrix_with_unidentified.txt", "confusion_matrix_with_unidentified.txt"),
                               (f"accuracies_type_4/values/k_{k}_accuracy_values.csv", "accuracy_values.csv")):
                assert open(os.path.join(root, ours), "rb").read() == open(os.path.join(GOLDEN, f"exp4_case{c}_k{k}_{gold}"), "rb").read(), (c, k, ours)
        # `cat values/*.csv`: shell glob order
        names = sorted(f for f in os.listdir(os.path.join(root, "accuracies_type_4/values")))
        cat = "".join(open(os.path.join(root, "accuracies_type_4/values", f)).read() for f in names)
        assert open(os.path.join(root, pipeline4.P_FINAL)).read() == cat
        fl = open(os.path.join(root, f"filelists_type_4/k_{ks[0]}/intersections_filelist.txt")).read().splitlines()
        assert len(fl) == case["n_groups"] ** 2 and fl[0].startswith(os.path.abspath(root))


class OracleEngine6(OracleEngine):
    """+ Engine.read_votes for pipeline6, as the reference's own loop over the stand-in's tables."""

    def read_votes(self, reads, k, pivot_keys, n_pivot, masks, n_groups):
        from khoice_b200 import merge_lists
        keys = pivot_keys.keys
        as_int = [int(x) for x in keys] if keys.ndim == 1 else [(int(h) << 64) | int(l) for l, h in keys]
        member = np.zeros((n_pivot, n_groups), dtype=bool)
        m = np.asarray(masks, dtype=np.uint64).reshape(n_pivot, -1)
        for d in range(n_groups):
            member[:, d] = (m[:, d // 64] >> np.uint64(d % 64)) & np.uint64(1)
        votes = merge_lists.votes_from_dump_index(reads, k, {x: i for i, x in enumerate(as_int)}, member, n_groups)
        return votes, np.zeros(len(reads), np.uint32)


def test_pipeline6_fused_on_oracle_engine_matches_reference_merge_lists(tmp_path):
    """Experiment type 6 driver (layout, file lists, per-(read type, k) seeding, final concatenation) without a GPU."""
    from khoice_b200 import pipeline6, synth
    import make_golden_exp6 as G6
    cases = json.load(open(os.path.join(GOLDEN, "exp6_cases.json")))["cases"]
    c, case = 1, cases[1]
    cfg, _, reads = G6.inputs_of(case)
    root = str(tmp_path / "w6")
    synth.write_dataset_type6(cfg, root, n_reads=case["n_reads"])
    for rt in G6.READ_TYPES:
        for p in range(case["n_groups"]):
            with open(os.path.join(root, pipeline6.p_reads(rt, p + 1)), "wb") as fd:
                fd.write(reads[rt][p])
    ks = [str(k) for k in case["k_values"]]
    rep = pipeline6.run_fused(root, case["n_groups"], ks, engine=OracleEngine6(), seed_fn=lambda rt, k: G6.seed_of(c, rt, k))
    assert rep["exp_type"] == 6 and rep["level"] == "read"
    for rt in G6.READ_TYPES:
        for k in ks:
            got = open(os.path.join(root, f"exp6_accuracies/{rt}/confusion_matrix/k_{k}_confusion_matrix.txt"), "rb").read()
            assert got == open(os.path.join(GOLDEN, f"exp6_case{c}_{rt}_k{k}_confusion_matrix.txt"), "rb").read(), (rt, k)
            got = open(os.path.join(root, f"exp6_accuracies/{rt}/values/k_{k}_accuracy_values.csv"), "rb").read()
            assert got == open(os.path.join(GOLDEN, f"exp6_case{c}_{rt}_k{k}_accuracy_values.csv"), "rb").read(), (rt, k)
        final = open(os.path.join(root, pipeline6.p_final(1, rt))).read()
        names = sorted(os.listdir(os.path.join(root, f"exp6_accuracies/{rt}/values")))
        assert final == pipeline6.HEADER + "".join(open(os.path.join(root, f"exp6_accuracies/{rt}/values", f)).read() for f in names)
        fl = open(os.path.join(root, f"exp6_filelists/k_{ks[0]}/{rt}/intersections_filelist.txt")).read().splitlines()
        assert len(fl) == case["n_groups"] ** 2 and fl[0].startswith(os.path.abspath(root)) and f"/{rt}/intersection/pivot_1/" in fl[0]
    assert os.path.exists(os.path.join(root, pipeline6.p_union_hist(ks[0], 1)))
    assert "exp6_genome_sets/rest_of_set" in open(os.path.join(root, pipeline6.p_ops(ks[0], 2))).read()
