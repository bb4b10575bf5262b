"""Per-source-line instruction / stall-sample totals of one kernel of an ncu report, by joining the report's SASS
page with `nvdisasm -g` line info of the object file.
usage: python scripts/ncu_lines.py report.ncu-rep kernel-index object.o mangled-substring [topn]"""
import csv, io, re, subprocess, sys, tempfile, os, collections
KSEL = (["--kernel-name", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else [])  # select one kernel of a multi-kernel report
rep, kidx, obj, sub = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"] + KSEL, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; data = []; seen = -1
for r in rows:
    if r and r[0] == "Address":
        seen += 1
        if seen > kidx: break
        hdr = r; data = []; continue
    if hdr and len(r) == len(hdr): data.append(r)
iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines = []; cur = None; infn = False
for l in dis:
    if l.startswith(".text."):
        infn = sub in l; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), "inl" if "inlined" in m.group(3) else ""); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l): lines.append(cur)
print("sass lines", len(lines), "ncu lines", len(data))
agg = collections.defaultdict(lambda: [0, 0])
for ln, r in zip(lines, data):
    key = ln[:2] if ln else ("?", 0)
    agg[key][0] += int(r[iI] or 0); agg[key][1] += int(r[iS] or 0)
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
srcs = {}
for (f, n), (i, s) in sorted(agg.items(), key=lambda t: -t[1][0])[:topn]:
    p = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
    if p not in srcs and os.path.exists(p): srcs[p] = open(p).read().splitlines()
    text = srcs.get(p, [""] * (n + 1))[n - 1].strip()[:90] if n and p in srcs and n <= len(srcs[p]) else ""
    print(f"{f}:{n:4d} inst {100*i/ti:5.1f}% ({i/1e6:8.2f} M)  samples {100*s/ts:5.1f}%  {text}")
# NCU_REGIONS="a-b,c-d": totals over source line ranges of the kernel's main file
if os.environ.get("NCU_REGIONS"):
    main = os.environ.get("NCU_FILE", "bins.cu")
    for rg in os.environ["NCU_REGIONS"].split(","):
        a, b = map(int, rg.split("-"))
        i = sum(v[0] for (f, n), v in agg.items() if f == main and a <= n <= b); sm_ = sum(v[1] for (f, n), v in agg.items() if f == main and a <= n <= b)
        print(f"region {main}:{a}-{b}: inst {100*i/ti:5.1f}% ({i/1e6:8.2f} M)  samples {100*sm_/ts:5.1f}%")
