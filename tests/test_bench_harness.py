"""bench.py's host-side harness, on the CPU: the workload switch (BASELINE configs 2-5), the KMC diff hook of SURVEY.md 8(c)
(a KMC 3 on PATH / under baseline/_ref runs the reference's unmodified shell strings and its histograms are diffed with the
oracle's), and the reference arm's thread count under torchrun's OMP_NUM_THREADS=1.  No KMC binary exists in this image, so
the hook is driven with a stand-in pair of executables that answer the same command lines through the oracle."""
import json
import os
import stat
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FAKE = textwrap.dedent('''\
    #!{python}
    # stand-in for kmc / kmc_tools (tests only): same command lines as /root/reference/workflow/rules/exp_type_1.smk:163-259
    import gzip, os, re, sys
    sys.path.insert(0, {root!r})
    import numpy as np
    from oracle import oracle as O
    a = sys.argv[1:]
    def load(p):
        d = np.load(p + ".kmc_pre.npz"); return d["keys"], d["counts"], int(d["k"])
    def save(p, keys, counts, k):
        np.savez(p + ".kmc_pre", keys=keys, counts=counts, k=k); os.replace(p + ".kmc_pre.npz", p + ".kmc_pre.npz"); open(p + ".kmc_suf", "wb").close()
    if os.path.basename(sys.argv[0]) == "kmc":
        k = int([x for x in a if x.startswith("-k")][0][2:])
        pos = [x for x in a if not x.startswith("-")]
        keys = O.genome_set(gzip.open(pos[0], "rb").read(), k)
        save(pos[1], keys, np.ones(keys.shape[0], np.uint32), k)
    elif a[0] == "transform" and a[2] == "set_counts":
        keys, counts, k = load(a[1]); save(a[4], keys, np.full(keys.shape[0], int(a[3]), np.uint32), k)
    elif a[0] == "transform" and a[2] == "histogram":
        keys, counts, k = load(a[1])
        h = O.histogram(counts, 300)
        open(a[3], "w").write("".join("%d\\t%d\\n" % (c, h[c]) for c in range(1, 256)))   # a different row count than ours
    elif a[0] == "complex":
        text = open(a[1]).read()
        ins = dict(re.findall(r"^(set\\d+) = (\\S+)", text, re.M))
        out, expr = re.search(r"OUTPUT:\\n(\\S+) = \\((.*)\\)", text).groups()
        cs = int(re.search(r"-cs(\\d+)", text).group(1))
        sets = [load(ins[s.strip()]) for s in expr.split("+")]
        keys, counts = O.union_sum([s[0] for s in sets], sets[0][2], cs)
        save(out, keys, counts, sets[0][2])
    else:
        sys.exit("unsupported: " + " ".join(a))
    ''')


@pytest.fixture()
def fake_kmc(tmp_path):
    d = tmp_path / "bin"
    d.mkdir()
    for exe in ("kmc", "kmc_tools"):
        p = d / exe
        p.write_text(FAKE.format(python=sys.executable, root=ROOT))
        p.chmod(p.stat().st_mode | stat.S_IEXEC)
    return str(d)


def test_workload_switch(monkeypatch):
    import bench
    for v in ("KHB_BENCH_CONFIG", "KHB_BENCH_GROUPS", "KHB_BENCH_GROUPS_TOTAL", "KHB_BENCH_GENOMES", "KHB_BENCH_LEN", "KHB_BENCH_K"):
        monkeypatch.delenv(v, raising=False)
    w = bench.workload(8)
    assert (w["groups_total"], w["genomes"], w["k"], w["scaling"], w["default_shape"]) == (80, 50, 31, "weak", True)
    monkeypatch.setenv("KHB_BENCH_CONFIG", "5")
    w = bench.workload(8)
    assert (w["groups_total"], w["genomes"], w["k"], w["scaling"], w["groups_per_gpu"]) == (100, 200, 31, "strong", 13)
    monkeypatch.setenv("KHB_BENCH_CONFIG", "4")
    monkeypatch.setenv("KHB_BENCH_K", "63")
    w = bench.workload(2)
    assert (w["groups_total"], w["genomes"], w["k"], w["scaling"], w["default_shape"]) == (20, 100, 63, "strong", True)
    monkeypatch.setenv("KHB_BENCH_CONFIG", "3")
    monkeypatch.delenv("KHB_BENCH_K")
    assert bench.workload(1)["ks"] == [7, 9, 11, 13, 15, 17, 19, 21, 23, 25, 27, 29, 31]
    # ~306 + 152 rho bytes per base at k = 31
    assert abs(bench.survey_bytes_per_base(31, 0.15) - (306 + 152 * 0.15)) < 2.0


def test_kmc_hook_runs_the_unmodified_rule_chain_and_diffs(monkeypatch, fake_kmc, oracle):
    import bench
    monkeypatch.setenv("PATH", fake_kmc + os.pathsep + os.environ["PATH"])
    for v, x in (("KHB_BENCH_CONFIG", "2"), ("KHB_BENCH_GROUPS", "2"), ("KHB_BENCH_GENOMES", "3"), ("KHB_BENCH_LEN", "30000"), ("KHB_BENCH_K", "21")):
        monkeypatch.setenv(v, x)
    found = bench.find_kmc()
    assert found and found["kmc"].startswith(fake_kmc)
    base, smp = bench.cpu_sample(bench.workload(1), 5.0)
    assert base["kind"] == "reference" and smp["kmc_diff"] == {"within_equal_oracle": True, "across_equal_oracle": True}
    assert smp["n_groups"] == 2 and smp["within"].shape[0] == 2 and int(smp["across"].sum()) > 0


def test_no_kmc_means_port(monkeypatch):
    import bench
    monkeypatch.setenv("PATH", "/usr/bin:/bin")
    assert bench.find_kmc() is None


def test_reference_arm_ignores_torchruns_single_thread(oracle):
    env = dict(os.environ, OMP_NUM_THREADS="1", KHB_BENCH_GROUPS="2", KHB_BENCH_GENOMES="2", KHB_BENCH_LEN="20000", RANK="0", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         env=env, capture_output=True, text=True, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["cores"] == oracle.host_cores() and line["e2e"]["h2d_bytes_per_step"] == 0
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env, capture_output=True, text=True, check=True).stdout
    assert out.strip() == ""
