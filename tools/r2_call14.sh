#!/bin/bash
# round 2, GPU call 14: validation of the round's final state: all GPU tests, default bench + reference arm,
# rows table, launch list, memcheck of the kernels added this round
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2_bench_n1_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_reference_final.json 2>/dev/null; echo "ref rc=$?"
timeout 300 python bench.py --workload hbf --layout 0 --steps 10 --profile > gpurun_out/r2_bench_hbf_fm.json 2>/dev/null
timeout 300 python bench.py --workload plumbing > gpurun_out/r2_bench_plumbing.json 2>/dev/null
timeout 900 python tools/bench_rows.py --out gpurun_out/r2_rows_final.json > gpurun_out/r2_rows_table_final.log 2>&1; echo "rows rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_biquad.csv python bench.py --steps 2 --warmup 1 --no-extra > /dev/null 2>&1; echo "ncu list rc=$? lines=$(wc -l < gpurun_out/r2_launches_biquad.csv)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_nco.py tests/test_gpu_hbf.py tests/test_gpu_fm_disc.py -m gpu -q -x -k "phase or lo_ or taps or chain_host or misaligned or fm_disc" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2_sanitizer_memcheck.log
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n1_final.json", "gpurun_out/r2_bench_reference_final.json", "gpurun_out/r2_bench_hbf_fm.json", "gpurun_out/r2_bench_plumbing.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "frac_pcie", (d.get("e2e") or {}).get("frac_of_pcie"))
    r = d.get("roofline") or {}
    print("  roofline", r.get("frac"), "sustained", (r.get("sustained") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("one_core"), "clocks", d.get("clocks"))
    for k, v in (d.get("extra") or {}).items():
        if "error" in v: print("  ", k, "ERROR", v["error"]); continue
        print("  ", k, v.get("value"), (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"), (v.get("e2e") or {}).get("frac_of_pcie"))
PY
