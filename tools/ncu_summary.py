#!/usr/bin/env python3
"""Turn an ncu report into the small per-kernel summary kept under profiles/.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/NAME.csv [--images N --traffic profiles/NAME_traffic.json]

Reads the report with `ncu -i REP --page raw --csv`, keeps the metrics the design discussion uses (duration, DRAM bytes,
issue slots, pipes, shared-memory wavefronts and bank conflicts, stall reasons, registers, occupancy limits) and writes
them one metric per row, one kernel launch per column.  With --images it also writes dram read + write bytes per image
for every kernel (the `roofline.traffic` source of bench.py)."""
import argparse
import csv
import json
import subprocess
import sys

KEEP = [
    "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--images", type=int, default=0)
    ap.add_argument("--traffic", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")] for r in data]
    cols = [i for i, h in enumerate(hdr) if h in KEEP or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio"))]
    with open(a.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        for i in cols:
            w.writerow([hdr[i], units[i]] + [r[i] for r in data])
    if a.images and a.traffic:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        per = {}
        for n, r in zip(names, data):
            b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
            key = n.split("(")[0]
            while key in per:
                key += "+"
            per[key] = b / a.images
        json.dump({"source": "ncu --set full --clock-control none, %s" % a.report, "images": a.images,
                   "dram_bytes_per_image": per}, open(a.traffic, "w"), indent=1)
    print("wrote", a.out, names, file=sys.stderr)


if __name__ == "__main__":
    main()
