"""Boundary fidelity of the Snakemake drop-in (khoice_b200/workflow/exp_type_1.smk): in BOTH modes every rule name of
/root/reference/workflow/rules/exp_type_1.smk:156-308 exists and every output pattern of the reference is produced by exactly
one rule -- the rule of the same name -- so any target a user of the reference names still resolves.  The reference's rules
come from tests/golden/exp1_rules.json (written by tests/golden/make_golden_rules.py from the reference file itself)."""
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = json.load(open(os.path.join(ROOT, "tests", "golden", "exp1_rules.json")))["rules"]
SMK = open(os.path.join(ROOT, "khoice_b200", "workflow", "exp_type_1.smk")).read()


def _rules_of(text):
    consts = {}
    for m in re.finditer(r'^(S\d) = (".*")$', SMK, re.M):
        consts[m.group(1)] = eval(m.group(2))
    out = {}
    for m in re.finditer(r"^\s*rule (\w+):\n(.*?)(?=^\s*rule |\Z)", text, re.M | re.S):
        name, body = m.group(1), m.group(2)
        o = re.search(r"^\s*output:\s*(.+)$", body, re.M).group(1)
        out[name] = {"output": list(eval("[" + o + "]", dict(consts))), "shell": (re.search(r'^\s*shell:\s*"(.*)"$', body, re.M) or [None, None])[1]}
    return out


def _modes():
    a = SMK.index('if KHB_MODE == "rules":')
    b = SMK.index("\nelse:\n", a)
    c = SMK.index("\nrule within_group_union_analysis:")
    common = _rules_of(SMK[c:])
    return {"rules": {**_rules_of(SMK[a:b]), **common}, "fused": {**_rules_of(SMK[b:c]), **common}}


@pytest.mark.parametrize("mode", ["rules", "fused"])
def test_every_reference_rule_and_output_pattern_has_exactly_one_producer(mode):
    mine = _modes()[mode]
    assert sorted(mine) == sorted(REF), (sorted(set(REF) - set(mine)), sorted(set(mine) - set(REF)))
    producers = {}
    for name, r in mine.items():
        for o in r["output"]:
            producers.setdefault(o, []).append(name)
    for name, r in REF.items():
        for o in r["output"]:
            assert producers.get(o) == [name], (mode, name, o, producers.get(o))
    assert sum(len(r["output"]) for r in mine.values()) == sum(len(r["output"]) for r in REF.values())   # and nothing else


def test_rules_mode_runs_the_reference_shell_strings():
    strip = lambda s: re.sub(r"^PATH=\{KHB_BIN\}:\$PATH ", "", s)
    mine = _modes()["rules"]
    for name, r in REF.items():
        if r["shell"]:
            ours = strip(mine[name]["shell"]).replace("{output}", r["output"][0].replace("{k}", "{wildcards.k}").replace("{num}", "{wildcards.num}"))
            assert ours == r["shell"], name
    # fused mode: the three transform rules keep the reference's strings too
    fused = _modes()["fused"]
    for name in ("within_group_union_histogram", "build_group_kmer_set", "across_group_union_histogram"):
        ours = strip(fused[name]["shell"]).replace("{output}", REF[name]["output"][0].replace("{k}", "{wildcards.k}").replace("{num}", "{wildcards.num}"))
        assert ours == REF[name]["shell"], name


@pytest.mark.parametrize("fused", [False, True])
def test_mini_runner_jobs_instantiate_the_same_dag(tmp_path, fused):
    """pipeline._rule_jobs / _fused_rule_jobs (the runner used where snakemake is not installed): one job per instance of
    every shell rule of the reference, the reference's outputs, each produced once; topological order."""
    from khoice_b200 import pipeline
    root = str(tmp_path)
    names = {1: ["a", "b"], 2: ["c"]}
    for n, gs in names.items():
        os.makedirs(os.path.join(root, f"data/dataset_{n}"))
        for g in gs:
            open(os.path.join(root, f"data/dataset_{n}/{g}.fna.gz"), "wb").close()
    jobs = (pipeline._fused_rule_jobs if fused else pipeline._rule_jobs)(root, ["9", "21"], 2)
    expect = {}
    for name, r in REF.items():
        if not r["shell"]:
            continue
        for k in ("9", "21"):
            insts = [dict(k=k)] if "{num}" not in r["output"][0] else \
                [dict(k=k, num=n, genome=g) for n, gs in names.items() for g in gs] if "{genome}" in r["output"][0] else [dict(k=k, num=n) for n in names]
            for w in insts:
                for o in r["output"]:
                    assert o.format(**w) not in expect
                    expect[o.format(**w)] = (name, r["shell"].replace("{input}", REF[name]["input"][0] if REF[name]["input"] else "").replace("wildcards.", "").format(**w))
    seen = {}
    for rule, outs, shell in jobs:
        for o in outs:
            assert o not in seen, o
            seen[o] = (rule, shell)
    assert sorted(seen) == sorted(expect)
    for o, (rule, shell) in seen.items():
        assert rule == expect[o][0], o
        if not fused:
            assert shell == expect[o][1], o
    # inputs exist before they are used: every step_N prefix a shell string reads was produced by an earlier job
    done = set()
    for rule, outs, shell in jobs:
        for tok in shell.split():
            if re.match(r"step_\d/", tok) and tok + ".kmc_pre" not in outs and not tok.endswith(".txt"):
                assert tok + ".kmc_pre" in done, (rule, tok)
        done.update(outs)
