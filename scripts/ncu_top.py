"""Summarise an ncu report: key raw metrics + top stall SASS lines.  usage: python scripts/ncu_top.py report.ncu-rep [n] [kernel-index]"""
import os, csv, subprocess, sys, io
KSEL = (["--kernel-name", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else [])  # select one kernel of a multi-kernel report
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30; kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + KSEL, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size',
        'smsp__thread_inst_executed.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor']
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w} [{units[i]}]: {[r[i] for r in rows[2:]]}")
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        vals = [float(r[i]) for r in rows[2:]]
        if max(vals) > 0.3: print(f"  stall {h.split('issue_stalled_')[1].split('_per_')[0]:20s} {vals}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"] + KSEL, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; data = []; seen = -1
for r in rows:
    if r and r[0] == "Address":
        seen += 1
        if seen > kidx: break
        hdr = r; data = []; continue
    if hdr and len(r) == len(hdr): data.append(r)
iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); iInst = hdr.index("Instructions Executed")
tot = sum(int(r[iS] or 0) for r in data); ninst = sum(int(r[iInst] or 0) for r in data)
print("total samples", tot, "SASS lines", len(data), "warp-instr", ninst)
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][iS] or 0))[:topn]:
    print(f"{idx:4d} {int(r[iS]):7d} {100 * int(r[iS]) / tot:5.1f}%  inst={r[iInst]:>9}  {r[iSrc][:100]}")
