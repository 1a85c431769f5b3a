#!/bin/bash
# round 2, GPU call 30: default bench with the frame-major HBF leg + ncu of the tensor-map frame-major kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_n1_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1_final.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["frac_of_pcie"])
for k, v in d["extra"].items():
    print("  ", k, v.get("value"), (v.get("roofline") or {}).get("frac"), v.get("kernel"), v.get("error"))
PY
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {
  local name=$1 k=$2 s=$3; shift 3
  timeout 600 $NCU -k "regex:$k" -s $s -c 1 -o gpurun_out/r2f_$name "$@" > gpurun_out/r2f_$name.log 2>&1
  echo "$name rc=$?"
  python tools/summarize_ncu.py gpurun_out/r2f_$name.ncu-rep gpurun_out/r2_${name}_final_ncu.md > /dev/null 2>&1
  python tools/ncu_segments.py gpurun_out/r2f_$name.ncu-rep 12 > gpurun_out/r2_${name}_final_segments.txt 2>&1
  rm -f gpurun_out/r2f_$name.ncu-rep
}
cap hbf_dec16_fm_tensormap hbf_dec_fast_kernel 1 python bench.py --workload hbf --layout 0 --steps 2 --warmup 1 --profile
cap hbf_int16_fm_tensormap hbf_int_fast_kernel 1 python tools/bench_rows.py --only "a14 HbfInt x16 cascade f32 frame-major" --reps 2 --out gpurun_out/x.json
