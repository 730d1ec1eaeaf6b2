#!/usr/bin/env python
"""Compact per-kernel summary of an ncu report (CSV of `ncu -i X.ncu-rep --page raw --csv`), for profiles/."""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = OrderedDict([
    ("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"), ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("smsp__inst_executed.sum", "warp_inst"),
])


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = OrderedDict()
    for r in data:
        name = r[idx["Kernel Name"]]
        name = name.replace("void <unnamed>::", "").replace("(<unnamed>::PairArgs)", "")
        seen.setdefault(name, []).append(r)
    for name, rs in seen.items():
        r = rs[-1]
        print("## %s   (%d captured launches; last shown)" % (name, len(rs)))
        out = []
        for k, short in KEYS.items():
            if k in idx and r[idx[k]] != "":
                v = r[idx[k]]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                out.append("%s=%s%s" % (short, v, (" " + units[idx[k]]) if units[idx[k]] not in ("%", "", "inst", "register/thread") else ""))
        st = []
        for h, i in idx.items():
            m = __import__("re").match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
            if m and r[i] != "":
                st.append((float(r[i]), m.group(1)))
        st.sort(reverse=True)
        out.append("stalls(warps per issue): " + " ".join("%s=%.2f" % (n, v) for v, n in st[:6]))
        for i in range(0, len(out), 6):
            print("   " + "  ".join(out[i:i + 6]))


if __name__ == "__main__":
    main()
