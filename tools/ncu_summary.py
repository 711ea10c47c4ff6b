"""Summarise an .ncu-rep: key metrics per kernel + top stall sites.   python tools/ncu_summary.py rep [--top 25]"""
import csv
import subprocess
import sys
import collections

rep = sys.argv[1]
top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "launch__registers_per_thread", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("=" * 100)
    for i, c in enumerate(h):
        if c in want or c.startswith("smsp__average_warps_issue_stalled") and c.endswith("per_issue_active.ratio") and float(r[i] or 0) > 0.3:
            print(f"  {c:95s} {rows[1][i]:10s} {r[i][:80]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
kern = None
data = []
def flush():
    if not data:
        return
    hh = hdr
    isrc, isamp, iex = hh.index("Source"), hh.index("# Samples"), hh.index("Instructions Executed")
    stall = [i for i, c in enumerate(hh) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[isamp]) for r in data)
    print("=" * 100)
    print(kern, "samples", tot, "warp-instr", sum(int(r[iex]) for r in data), "SASS lines", len(data))
    agg = collections.Counter()
    for r in data:
        for j in stall:
            agg[hh[j][6:]] += int(r[j])
    print("  stall mix:", ", ".join(f"{k} {100 * v / max(1, tot):.1f}%" for k, v in agg.most_common(8)))
    for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top_n]):
        r = data[i]
        st = sorted(((hh[j][6:], int(r[j])) for j in stall if int(r[j]) > 0), key=lambda x: -x[1])[:2]
        print(f"  {i:5d} {100 * int(r[isamp]) / max(1, tot):5.1f}% ex={r[iex]:>10s} {r[isrc].strip()[:64]:64s} {st}")
hdr = None
for r in rows:
    if r and r[0] == "Kernel Name":
        flush(); data = []; kern = r[1][:90]
    elif r and r[0] == "Address":
        hdr = r
    elif hdr and len(r) >= len(hdr):
        data.append(r)
flush()
