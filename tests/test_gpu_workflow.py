"""End-to-end drop-in check on a GPU: the exp-1 rule chain over a work root, in all three modes, against
the oracle's histograms and byte-identical CSVs."""
import filecmp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K_VALUES = ["7", "12", "21", "31", "34"]


@pytest.fixture(scope="module")
def work_roots(tmp_path_factory):
    from khoice_b200 import synth
    cfg = synth.SynthConfig(n_groups=2, genomes_per_group=3, genome_len=40_000, seed=5)
    roots = {}
    for mode in ("fused", "rules", "rules-subprocess", "smk-fused"):
        root = str(tmp_path_factory.mktemp(mode.replace("-", "_")))
        synth.write_dataset(cfg, root)
        roots[mode] = root
    return cfg, roots


def _oracle_hists(oracle, cfg, k):
    from khoice_b200 import synth
    flat, gid = [], []
    for g in range(1, cfg.n_groups + 1):
        for i in range(1, cfg.genomes_per_group + 1):
            flat.append(synth.make_genome(cfg, g, i))
            gid.append(g - 1)
    return oracle.exp1(flat, gid, cfg.n_groups, k)


def test_three_modes_agree_with_oracle(engine, oracle, work_roots):
    from khoice_b200 import pipeline, tables
    cfg, roots = work_roots
    pipeline.run_fused(roots["fused"], cfg.n_groups, K_VALUES, engine=engine)
    pipeline.run_rules(roots["rules"], cfg.n_groups, K_VALUES, engine=engine)
    pipeline.run_rules(roots["rules-subprocess"], cfg.n_groups, K_VALUES[:2] , subprocess_mode=True)
    pipeline.run_rules(roots["smk-fused"], cfg.n_groups, K_VALUES, engine=engine, fused_rules=True)
    for k in K_VALUES:
        w_ref, a_ref, _ = _oracle_hists(oracle, cfg, int(k))
        for mode in ("fused", "rules", "smk-fused") + (("rules-subprocess",) if k in K_VALUES[:2] else ()):
            root = roots[mode]
            for num in range(1, cfg.n_groups + 1):
                got = tables.read_histogram_file(os.path.join(root, pipeline.p_step4(k, num)))
                assert len(got) == 5000
                assert got == [int(x) for x in w_ref[num - 1][1:]], (mode, k, num)
            got = tables.read_histogram_file(os.path.join(root, pipeline.p_step8(k)))
            assert got == [int(x) for x in a_ref[1:]], (mode, k)
    for f in (pipeline.P_STEP5, pipeline.P_STEP9) + pipeline.P_FINAL:
        assert filecmp.cmp(os.path.join(roots["fused"], f), os.path.join(roots["rules"], f), shallow=False), f
        assert filecmp.cmp(os.path.join(roots["fused"], f), os.path.join(roots["smk-fused"], f), shallow=False), f
    # every declared rule output exists in every mode (file DAG / resume semantics)
    for mode in ("fused", "rules", "smk-fused"):
        root = roots[mode]
        for k in K_VALUES:
            for num in range(1, cfg.n_groups + 1):
                for g in pipeline.genomes_of(root, num):
                    for p in (pipeline.p_step1(k, num, g), pipeline.p_step2(k, num, g)):
                        assert os.path.exists(os.path.join(root, p + ".kmc_pre")) and os.path.exists(os.path.join(root, p + ".kmc_suf"))
                for p in (pipeline.p_step3(k, num), pipeline.p_step6(k, num)):
                    assert os.path.exists(os.path.join(root, p + ".kmc_pre"))
            assert os.path.exists(os.path.join(root, pipeline.p_step7(k) + ".kmc_pre"))
    # resume: a second run does nothing
    rep = pipeline.run_rules(roots["rules"], cfg.n_groups, K_VALUES, engine=engine)
    assert rep["jobs_run"] == 0
    rep = pipeline.run_rules(roots["smk-fused"], cfg.n_groups, K_VALUES, engine=engine, fused_rules=True)
    assert rep["jobs_run"] == 0
    # fused placeholders are never mistaken for k-mer sets: a rule-compatible re-run over them fails loudly (exit 1), it does
    # not write an empty union; the set-only step_3 table refuses whatever needs its counters
    from khoice_b200 import cli
    k = K_VALUES[0]
    cli.set_engine(engine)
    try:
        cwd = os.getcwd()
        os.chdir(roots["smk-fused"])
        os.remove(pipeline.p_step3(k, 1) + ".kmc_pre")
        assert cli.main(["kmc_tools", "complex", pipeline.p_ops_within(k, 1)]) == 1
        assert not os.path.exists(pipeline.p_step3(k, 1) + ".kmc_pre")
        assert cli.main(["kmc_tools", "transform", pipeline.p_step3(k, 2), "dump", "-s", "dump.txt"]) == 1 and not os.path.exists("dump.txt")
        assert cli.main(["kmc_tools", "transform", pipeline.p_step3(k, 2), "histogram", "h.txt"]) == 0
    finally:
        os.chdir(cwd)
        cli.set_engine(None)


def test_rule_databases_hold_the_oracle_sets(engine, oracle, work_roots):
    from khoice_b200 import kmcdb, pipeline, synth
    cfg, roots = work_roots
    root = roots["rules"]
    if not os.path.exists(os.path.join(root, pipeline.P_STEP5)):
        pipeline.run_rules(root, cfg.n_groups, K_VALUES, engine=engine)
    for k in ("21", "34"):
        for num in (1, 2):
            genomes = [synth.make_genome(cfg, num, i) for i in range(1, cfg.genomes_per_group + 1)]
            sets = [oracle.genome_set(g, int(k)) for g in genomes]
            for name, ref in zip(pipeline.genomes_of(root, num), sets):
                db1 = kmcdb.read_db(os.path.join(root, pipeline.p_step1(k, num, name)))
                db2 = kmcdb.read_db(os.path.join(root, pipeline.p_step2(k, num, name)))
                assert np.array_equal(db1.keys, ref) and np.array_equal(db2.keys, ref)
                assert (db2.counts == 1).all() and db1.counts.min() >= 1
            keys, counts = oracle.union_sum(sets, int(k))
            db3 = kmcdb.read_db(os.path.join(root, pipeline.p_step3(k, num)))
            assert np.array_equal(db3.keys, keys) and np.array_equal(db3.counts, counts)


def test_cli_error_behaviour(tmp_path):
    from khoice_b200 import cli
    assert cli.main(["kmc", "-k31", "-ci1", "x.fna", "out", "tmp"]) == 1          # no -fm
    assert cli.main(["kmc_tools", "simple", "a", "b", "intersect", "c"]) == 1     # not an exp-1 operation
    assert cli.main(["kmc_tools", "transform", str(tmp_path / "missing"), "histogram", str(tmp_path / "h.txt")]) == 1
    assert not (tmp_path / "h.txt").exists()


def test_rule_chain_on_kmc_layout_databases(engine, oracle, work_roots, tmp_path, monkeypatch):
    """with KHB_DB_FORMAT=kmc1 every rule exchanges databases in KMC's own (KMC1) layout -- the files a
    real kmc_tools would be handed -- and the histograms / CSVs do not change."""
    from khoice_b200 import kmc_format, kmcdb, pipeline, synth
    cfg, roots = work_roots
    root = str(tmp_path / "kmc_layout")
    synth.write_dataset(cfg, root)
    ks = ["12", "31", "34"]
    monkeypatch.setenv("KHB_DB_FORMAT", "kmc1")
    pipeline.run_rules(root, cfg.n_groups, ks, engine=engine)
    monkeypatch.delenv("KHB_DB_FORMAT")
    ref_root = str(tmp_path / "own_layout")
    synth.write_dataset(cfg, ref_root)
    pipeline.run_rules(ref_root, cfg.n_groups, ks, engine=engine)
    for f in (pipeline.P_STEP5, pipeline.P_STEP9):
        assert filecmp.cmp(os.path.join(root, f), os.path.join(ref_root, f), shallow=False), f
    for k in ks:
        for num in range(1, cfg.n_groups + 1):
            assert filecmp.cmp(os.path.join(root, pipeline.p_step4(k, num)), os.path.join(ref_root, pipeline.p_step4(k, num)), shallow=False)
            p3 = os.path.join(root, pipeline.p_step3(k, num))
            assert kmc_format.is_kmc_database(p3)
            hdr = kmc_format.read_header(p3)
            assert hdr["k"] == int(k) and hdr["counter_size"] == 2 and hdr["version"] == 0     # -cs5000: two counter bytes
            a, b = kmcdb.read_db(p3), kmcdb.read_db(os.path.join(ref_root, pipeline.p_step3(k, num)))
            assert np.array_equal(a.keys, b.keys) and np.array_equal(a.counts, b.counts)
        assert filecmp.cmp(os.path.join(root, pipeline.p_step8(k)), os.path.join(ref_root, pipeline.p_step8(k)), shallow=False)


def test_unmodified_rule_chain_through_the_worker(work_roots, tmp_path):
    """Every rule instance is its own `sh -c "kmc ..."` process, as under Snakemake, but the shims forward to ONE long-lived
    worker (khoice_b200/worker.py) instead of creating a CUDA context each: same files, same CSV bytes."""
    import subprocess, sys, time
    from khoice_b200 import pipeline, synth
    cfg, roots = work_roots
    root = str(tmp_path / "via_worker")
    synth.write_dataset(cfg, root)
    sock = str(tmp_path / "khb.sock")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    srv = subprocess.Popen([sys.executable, "-m", "khoice_b200.worker", "--socket", sock], cwd=repo, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        assert "ready" in srv.stdout.readline()
        os.environ["KHB_WORKER_SOCKET"] = sock
        t0 = time.time()
        rep = pipeline.run_rules(root, cfg.n_groups, K_VALUES[:3], subprocess_mode=True)
        dt = time.time() - t0
    finally:
        os.environ.pop("KHB_WORKER_SOCKET", None)
        subprocess.run([sys.executable, "-m", "khoice_b200.worker", "--socket", sock, "--stop"], cwd=repo, timeout=60)
        out, err = srv.communicate(timeout=60)
    assert srv.returncode == 0 and f"served {rep['jobs_run']} requests" in out, out + err
    assert rep["jobs_run"] == 3 * (2 * cfg.n_groups * cfg.genomes_per_group + 3 * cfg.n_groups + 2)   # per k: kmc + set_counts per genome; union, histogram, set per group; across union + histogram
    ref = str(tmp_path / "in_process")
    synth.write_dataset(cfg, ref)
    from khoice_b200.engine import Engine
    eng = Engine(0)
    try:
        pipeline.run_rules(ref, cfg.n_groups, K_VALUES[:3], engine=eng)
    finally:
        eng.close()
    for f in (pipeline.P_STEP5, pipeline.P_STEP9):
        assert filecmp.cmp(os.path.join(root, f), os.path.join(ref, f), shallow=False), f
    print(f"{rep['jobs_run']} rule processes through the worker in {dt:.1f} s ({dt / rep['jobs_run'] * 1e3:.0f} ms each)")
