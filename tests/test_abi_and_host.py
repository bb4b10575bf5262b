"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the product never touches
the oracle, the host-side logic (database files, operation files, CLI parsing) behaves like the reference."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    from khoice_b200 import engine
    engine.build()
    hdr = open(os.path.join(ROOT, "include", "khoice_b200.h")).read()
    declared = re.findall(r"KHB_API[^;(]*?\b(khb_[a-z0-9_]+)\s*\(", hdr)
    assert len(declared) >= 30 and len(set(declared)) == len(declared)
    lib = engine.load_library()
    for name in declared:
        assert getattr(lib, name) is not None
    assert sorted(declared) == sorted(engine.EXPORTED_SYMBOLS)
    assert lib.khb_abi_version() == 1
    nm = subprocess.run(["nm", "-D", "--defined-only", engine.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (khb_[a-z0-9_]+)", nm))
    assert exported == set(declared)


def test_no_gpu_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from khoice_b200.engine import Engine, KhbError
    with pytest.raises(KhbError) as e:
        Engine(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "khoice_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".smk")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "libkmer_oracle" not in text and "oracle/" not in text.replace("tests/", ""), f
    code = "import sys; import khoice_b200.pipeline, khoice_b200.cli, khoice_b200.dist; assert not any(m.split('.')[0]=='oracle' for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_kmcdb_roundtrip(tmp_path):
    from khoice_b200 import kmcdb
    rng = np.random.default_rng(0)
    for k, shape in ((31, (1000,)), (47, (777, 2))):
        keys = rng.integers(0, 2**62, size=shape, dtype=np.uint64)
        counts = rng.integers(1, 9, size=shape[0]).astype(np.uint32)
        hist = np.bincount(counts, minlength=5001).astype(np.uint64)
        p = str(tmp_path / f"db{k}")
        kmcdb.write_db(p, k, keys, counts, hist, 255)
        db = kmcdb.read_db(p)
        assert db.k == k and np.array_equal(db.keys, keys) and np.array_equal(db.counts, counts) and np.array_equal(db.hist, hist)
        assert kmcdb.read_db(p, header_only=True).n_keys == shape[0]
    kmcdb.write_db(str(tmp_path / "stub"), 31, None, None, np.zeros(5001, np.uint64), 5000, n_keys=42)
    s = kmcdb.read_db(str(tmp_path / "stub"), header_only=True)           # `transform X histogram` may read a stub ...
    assert s.stub and s.n_keys == 42 and os.path.exists(tmp_path / "stub.kmc_suf")
    with pytest.raises(ValueError, match="header-only stub"):               # ... nothing may read it as a k-mer set
        kmcdb.read_db(str(tmp_path / "stub"))
    from khoice_b200 import cli
    for argv in (["kmc_tools", "transform", str(tmp_path / "stub"), "set_counts", "1", str(tmp_path / "o")],
                 ["kmc_tools", "transform", str(tmp_path / "stub"), "dump", "-s", str(tmp_path / "o.txt")]):
        assert cli.main(argv) == 1 and not os.path.exists(tmp_path / "o.kmc_pre") and not os.path.exists(tmp_path / "o.txt")
    assert cli.main(["kmc_tools", "transform", str(tmp_path / "stub"), "histogram", str(tmp_path / "h.txt")]) == 0
    (tmp_path / "bad.kmc_pre").write_bytes(b"KMCP" + b"\0" * 100)
    with pytest.raises(ValueError):
        kmcdb.read_db(str(tmp_path / "bad"))


def test_complex_ops_files_match_reference(tmp_path):
    from khoice_b200 import cli, pipeline
    gold = json.load(open(os.path.join(GOLDEN, "complex_ops.json")))
    for n, names in gold["members"].items():
        d = tmp_path / "data" / f"dataset_{n}"
        d.mkdir(parents=True)
        for g in names:
            (d / f"{g}.fna.gz").write_bytes(b"")
        (d / "notes.txt").write_text("x")
    pipeline.write_complex_ops(str(tmp_path), gold["k_values"], len(gold["members"]))
    assert (tmp_path / "tmp").is_dir()
    for rel, ref_text in gold["files"].items():
        mine = (tmp_path / rel).read_text()
        # set numbering follows os.listdir order in the reference (arbitrary) -> compare parsed content
        ref_path = tmp_path / "ref_ops.txt"
        ref_path.write_text(ref_text)
        r_in, r_out, r_cs = cli.parse_complex(str(ref_path))
        m_in, m_out, m_cs = cli.parse_complex(str(tmp_path / rel))
        assert sorted(r_in) == sorted(m_in) and r_out == m_out and r_cs == m_cs == 5000
        assert mine.splitlines()[0] == "INPUT:" and mine.endswith("OUTPUT_PARAMS:\n-cs5000\n")
        assert len(mine.splitlines()) == len(ref_text.splitlines())


def test_cli_argument_handling_without_gpu(tmp_path):
    from khoice_b200 import cli
    with pytest.raises(cli.UsageError):
        cli.kmc_main(["-k31", "-ci1", "a.fna.gz", "out", "tmp/"])            # -fm missing
    with pytest.raises(cli.UsageError):
        cli.kmc_main(["-fm", "-k99", "-ci1", "a.fna.gz", "out", "tmp/"])     # k out of range
    with pytest.raises(cli.UsageError):
        cli.kmc_tools_main(["simple", "a", "b", "intersect", "c"])
    ops = tmp_path / "ops.txt"
    ops.write_text("INPUT:\nset1 = a\nset2 = b\nOUTPUT:\nout = (set1 - set2)\nOUTPUT_PARAMS:\n-cs5000\n")
    with pytest.raises(cli.UsageError):
        cli.parse_complex(str(ops))
    ops.write_text("INPUT:\nset1 = a\nset2 = b\nOUTPUT:\nout = (set1 + set2 )\nOUTPUT_PARAMS:\n-cs5000\n")
    assert cli.parse_complex(str(ops)) == (["a", "b"], "out", 5000)
    # set_counts / histogram are host-side file rewrites and work without a GPU
    from khoice_b200 import kmcdb
    keys = np.arange(5, dtype=np.uint64)
    kmcdb.write_db(str(tmp_path / "x"), 9, keys, np.array([3, 1, 2, 2, 9], np.uint32), np.zeros(5001, np.uint64), 255)
    assert cli.main(["kmc_tools", "transform", str(tmp_path / "x"), "set_counts", "1", str(tmp_path / "y")]) == 0
    assert cli.main(["kmc_tools", "transform", str(tmp_path / "y"), "histogram", str(tmp_path / "h.txt")]) == 0
    lines = (tmp_path / "h.txt").read_text().splitlines()
    assert lines[0] == "1\t5" and lines[1] == "2\t0" and len(lines) == 5000
    assert cli.main(["kmc_tools", "transform", str(tmp_path / "nope"), "histogram", str(tmp_path / "h2.txt")]) == 1


def test_default_k_values_are_the_reference_list():
    from khoice_b200 import pipeline
    assert pipeline.DEFAULT_K_VALUES == [str(x) for x in range(7, 31)] + ["34", "37", "40", "43", "46", "49"]
