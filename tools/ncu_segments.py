#!/usr/bin/env python3
"""Split an ncu source page (`ncu -i rep --page source --csv`) of one kernel into the segments
between barriers and print stall samples per segment + the hottest instructions.

    python tools/ncu_segments.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    print(rows[0][1][:110])
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    g = lambda r, k: int(r[ix[k]] or 0)
    tot = sum(g(r, "# Samples") for r in data)
    print("samples", tot, "sass instructions", len(data), "warp-instr executed", sum(g(r, "Instructions Executed") for r in data))
    keys = ["stall_barrier", "stall_short_sb", "stall_long_sb", "stall_wait", "stall_math", "stall_mio", "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_no_inst"]
    cur = {"start": 0, "samples": 0, "insts": 0, **{k: 0 for k in keys}}
    segs = []
    for i, r in enumerate(data):
        cur["samples"] += g(r, "# Samples")
        cur["insts"] += g(r, "Instructions Executed")
        for k in keys:
            cur[k] += g(r, k)
        src = r[ix["Source"]]
        if "BAR.SYNC" in src or "SYNCS.PHASECHK" in src:
            cur["end"], cur["at"] = i, src.strip()[:40]
            segs.append(cur)
            cur = {"start": i + 1, "samples": 0, "insts": 0, **{k: 0 for k in keys}}
    cur["end"], cur["at"] = len(data), "end"
    segs.append(cur)
    for s in segs:
        if s["samples"] * 200 < tot:
            continue
        print(f"[{s['start']:5d}-{s['end']:5d}] {100 * s['samples'] / tot:5.1f}% samples  {s['insts'] / 1e6:8.1f}M inst  " +
              " ".join(f"{k[6:]}={100 * s[k] / tot:.1f}" for k in keys if s[k] * 100 >= tot) + f"  -> {s['at']}")
    for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:topn]:
        print(f"{100 * g(r, '# Samples') / tot:5.1f}%  {r[ix['Source']].strip()[:60]:60s} " +
              " ".join(f"{k[6:]}={g(r, k)}" for k in keys if g(r, k) * 10 >= max(1, g(r, '# Samples'))))


if __name__ == "__main__":
    main()
