#!/usr/bin/env python
"""Summarise an .ncu-rep (from gpurun_out/) into profiles/: per-launch table of the metrics the roofline
uses.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_full.md [profiles/traffic.json]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy%"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "alu%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_longsb"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    lines = ["| kernel | " + " | ".join(k[1] for k in KEYS) + " |", "|---|" + "---|" * len(KEYS)]
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "")
        cells = []
        for key, _ in KEYS:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[i]}".strip())
            else:
                cells.append("-")
        lines.append(f"| {name} | " + " | ".join(cells) + " |")
    with open(out, "w") as f:
        f.write(f"ncu --set full --clock-control none summary of {rep} (one row per captured launch)\n\n")
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if len(sys.argv) > 3:
        # DRAM bytes (read + write) per launch, averaged per kernel class: bench.py's roofline.traffic
        import json
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        acc = {}
        for r in rows[2:]:
            name = r[ki].split("(")[0].replace("void ", "").replace("b200he::", "").split("<")[0]
            b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
            acc.setdefault(name, []).append(b)
        with open(sys.argv[3], "w") as f:
            json.dump({"source": rep, "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), mean over captured launches",
                       **{k: sum(v) / len(v) for k, v in acc.items()}}, f, indent=1)


if __name__ == "__main__":
    main()
