#!/usr/bin/env python
"""Summarise an ncu report (--set full) into a small text file for profiles/ and
update profiles/traffic.json (dram bytes per launch of each hot kernel).

    python tools/ncu_summary.py gpurun_out/r01_kernels.ncu-rep profiles/r01_kernels_summary.txt [n_local]
"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    nloc = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu --set full --clock-control none; source report: %s" % os.path.basename(rep)]
    traffic = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        lines.append("")
        lines.append("kernel: " + name)
        for w in WANT:
            if w in idx:
                lines.append("  %-68s %s %s" % (w, r[idx[w]], units[idx[w]]))
        st = sorted(((float(r[idx[h]]), h.replace("smsp__pcsamp_warps_issue_stalled_", ""))
                     for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and r[idx[h]]),
                    reverse=True)[:8]
        lines.append("  warp stall samples: " + ", ".join("%s=%d" % (h, v) for v, h in st))
        rd = float(r[idx["dram__bytes_read.sum"]]) * UNIT[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * UNIT[units[idx["dram__bytes_write.sum"]]]
        lines.append("  dram bytes per launch (read + write): %.4e" % (rd + wr))
        key = "gram" if "k_gram" in name else "direction" if "k_direction" in name else \
            "trial" if "k_trial" in name else "spg" if "k_spg" in name else None
        if key and nloc:
            traffic[key] = {"kernel": name, "n": nloc, "dram_bytes_per_launch": rd + wr,
                            "source": os.path.basename(out)}
    with open(out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if traffic:
        tj = os.path.join(os.path.dirname(out), "traffic.json")
        old = {}
        if os.path.exists(tj):
            old = json.load(open(tj))
        old.update(traffic)
        json.dump(old, open(tj, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
