"""Writes tests/golden/exp1_rules.json: for every rule of /root/reference/workflow/rules/exp_type_1.smk:156-308 its name and
its literal input / output path patterns, parsed from the reference file itself (run in a container that has /root/reference;
the fixture travels to the GPU box, the reference does not).  tests/test_workflow_dag.py holds the drop-in's .smk to it."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/workflow/rules/exp_type_1.smk"
text = open(REF).read()
rules = {}
for m in re.finditer(r"^rule (\w+):\n(.*?)(?=^rule |\Z)", text, re.M | re.S):
    name, body = m.group(1), m.group(2)
    sec = {}
    for s in re.finditer(r"^    (input|output|shell|run):[ \t]*\n?(.*?)(?=^    (?:input|output|shell|run|params):|\Z)", body, re.M | re.S):
        sec[s.group(1)] = s.group(2)
    strings = lambda t: re.findall(r'"([^"]+)"', t or "")
    rules[name] = {"input": [x for x in strings(sec.get("input")) if "/" in x], "output": strings(sec.get("output")),
                   "shell": strings(sec.get("shell"))[0] if "shell" in sec else None}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exp1_rules.json")
json.dump({"source": "workflow/rules/exp_type_1.smk", "rules": rules}, open(out, "w"), indent=1)
print(out, len(rules), "rules")
