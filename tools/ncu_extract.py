"""Compact per-kernel table from an ncu report: `ncu -i X.ncu-rep --page raw --csv | python tools/ncu_extract.py`.
Prints duration, DRAM / L2 bytes, tensor-pipe activity and the top warp-stall reasons for every profiled launch."""
import csv, sys, re

rows = list(csv.reader(sys.stdin))
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = i
        break
if hdr is None:
    sys.exit("no header")
names = rows[hdr]
units = rows[hdr + 1]
col = {n: j for j, n in enumerate(names)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
stall = [n for n in names if re.match(r"smsp__average_warps?_issue_stalled_.*_per_issue_active|smsp__average_warp_latency_issue_stalled", n)
         or re.match(r"smsp__warp_issue_stalled_.*_per_warp_active\.pct", n)]
for r in rows[hdr + 2:]:
    if len(r) < len(names):
        continue
    print("==", re.sub(r"\(.*", "", r[col["Kernel Name"]])[:100], " grid", r[col.get("Grid Size", 0)] if "Grid Size" in col else "")
    for w in want:
        if w in col:
            print("   %-72s %-10s %s" % (w, units[col[w]], r[col[w]]))
    sv = []
    for s_ in stall:
        try:
            sv.append((float(r[col[s_]].replace(",", "")), s_))
        except ValueError:
            pass
    for v, s_ in sorted(sv, reverse=True)[:4]:
        print("   stall %-66s %.2f" % (s_.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("smsp__warp_issue_stalled_", ""), v))
