#!/usr/bin/env python3
"""Turn an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a short markdown summary
under profiles/, and update profiles/roofline_traffic.json (dram bytes per launch).

    python tools/summarize_ncu.py gpurun_out/prof_x.ncu-rep profiles/r1_x.md [traffic_key]
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary: `{os.path.basename(rep)}`", "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun (1x B200);",
             "per-launch values (cold-cache, serialised replays: compare shares, not absolutes).", ""]
    traffic = None
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines += [f"## `{name[:160]}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                lines.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if v > 0.05:
                    stalls.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        lines += ["", "warp stall reasons (avg warps stalled per issue-active cycle): " +
                  ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)), ""]
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        traffic = rd + wr
        lines += [f"dram traffic per launch = {rd + wr:.4g} B (read {rd:.4g} + write {wr:.4g})", ""]
    open(out, "w").write("\n".join(lines) + "\n")
    if key and traffic:
        p = os.path.join(os.path.dirname(out), "roofline_traffic.json")
        d = json.load(open(p)) if os.path.exists(p) else {}
        d[key] = traffic
        json.dump(d, open(p, "w"), indent=1, sort_keys=True)
    print("wrote", out)


if __name__ == "__main__":
    main()
