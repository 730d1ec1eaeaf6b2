#!/usr/bin/env python
"""profiles/fp64_ops.json from an ncu report (`--set full`): executed FP64 thread-instructions of the pair kernels, MEASURED.

  python scripts/ncu_ops.py <report.ncu-rep> <particles> [out.json]

ncu's roofline section carries smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on.sum.per_cycle_elapsed; times
smsp__cycles_elapsed.max that is the number of executed (predicated-on) thread instructions of the launch. bench.py turns them into
flop per particle (dadd + dmul + 2 dfma) for the FP64 side of the roofline -- replacing SURVEY.md's flop-per-pair estimates.
"""
import csv
import json
import subprocess
import sys

FAMILY = {"k_accel": "accel", "k_rate": "rate_stress", "k_post": "post"}


def main():
    rep, particles = sys.argv[1], int(sys.argv[2])
    out_path = sys.argv[3] if len(sys.argv) > 3 else "profiles/fp64_ops.json"
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, data = rows[0], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in data:
        name = r[idx["Kernel Name"]]
        fam = next((v for k, v in FAMILY.items() if k in name), None)
        if not fam:
            continue
        cyc = float(r[idx["smsp__cycles_elapsed.max"]].replace(",", ""))
        rec = {"particles": particles, "kernel": name.replace("void <unnamed>::", "").replace("(<unnamed>::PairArgs)", ""), "report": rep.split("/")[-1]}
        for op in ("dadd", "dmul", "dfma"):
            rec[op] = float(r[idx["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op]].replace(",", "")) * cyc
        rec["pipe_fp64_pct"] = float(r[idx["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]].replace(",", ""))
        rec["ms"] = float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(rows[1][idx["gpu__time_duration.sum"]], 1.0)
        rec["dram_bytes_per_launch"] = (float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(rows[1][idx["dram__bytes_read.sum"]], 1.0) +
                                        float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(rows[1][idx["dram__bytes_write.sum"]], 1.0))
        out[fam] = rec                      # last captured launch of the family
    json.dump(out, open(out_path, "w"), indent=1)
    if len(sys.argv) > 4:                   # profiles/traffic.json: measured DRAM bytes per launch, the `traffic` key of bench.py's roofline
        json.dump({fam: {"particles": particles, "dram_bytes_per_launch": rec["dram_bytes_per_launch"], "source": "%s %s: dram__bytes_read.sum + dram__bytes_write.sum" % (rec["report"], rec["kernel"])}
                   for fam, rec in out.items()}, open(sys.argv[4], "w"), indent=1)
    for fam, rec in out.items():
        print("%-12s %s: dadd %.3g dmul %.3g dfma %.3g -> %.0f flop/particle, FP64 pipe %.1f %%, %.2f ms" %
              (fam, rec["kernel"], rec["dadd"], rec["dmul"], rec["dfma"], (rec["dadd"] + rec["dmul"] + 2 * rec["dfma"]) / particles, rec["pipe_fp64_pct"], rec["ms"]))


if __name__ == "__main__":
    main()
