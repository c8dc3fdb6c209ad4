"""Brief per-launch table from an .ncu-rep (raw page): duration, DRAM bytes / throughput, instructions, issue activity, occupancy,
top stall reasons.   python scripts/ncu_brief.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    g = lambda k: r[col[k]] if k in col else "?"
    dur = float(g("gpu__time_duration.sum"))
    rd, wr = float(g("dram__bytes_read.sum")), float(g("dram__bytes_write.sum"))
    unit_r = rows[1][col["dram__bytes_read.sum"]]
    mul = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(unit_r, 1e6)
    st = sorted(((float(r[col[s]]), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls), reverse=True)[:4]
    print("%-46s grid %6s blk %4s regs %3s | %7.1f us | dram %6.1f+%6.1f MB = %5.2f TB/s | inst %9.0f | issue %4.1f%% warps %4.1f%% | %s" % (
        g("Kernel Name")[:46], g("launch__grid_size"), g("launch__block_size"), g("launch__registers_per_thread"), dur,
        rd * mul / 1e6, wr * mul / 1e6, (rd + wr) * mul / dur / 1e6, float(g("smsp__inst_executed.sum")),
        float(g("smsp__issue_active.avg.pct_of_peak_sustained_active")), float(g("sm__warps_active.avg.pct_of_peak_sustained_active")),
        " ".join("%s %.1f" % (n, v) for v, n in st)))
