#!/bin/bash
# round 2, GPU call 23: ncu --set full of the final lock-in, chain and HBF kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {
  local name=$1 k=$2 s=$3; shift 3
  timeout 600 $NCU -k "regex:$k" -s $s -c 1 -o gpurun_out/r2f_$name "$@" > gpurun_out/r2f_$name.log 2>&1
  echo "$name rc=$?"
  python tools/summarize_ncu.py gpurun_out/r2f_$name.ncu-rep gpurun_out/r2_${name}_final_ncu.md > /dev/null 2>&1
  python tools/ncu_segments.py gpurun_out/r2f_$name.ncu-rep 12 > gpurun_out/r2_${name}_final_segments.txt 2>&1
}
cap lockin_fm tma_lanes_kernel 1 python bench.py --workload lockin --steps 2 --warmup 1 --profile
cap hbf_dec16_lm hbf_dec_fast_kernel 1 python bench.py --workload hbf --steps 2 --warmup 1 --profile
cap hbf_int16_lm hbf_int_fast_kernel 1 python tools/bench_rows.py --only "a14 HbfInt x16 cascade f32 lane-major" --reps 2 --out gpurun_out/x.json
rm -f gpurun_out/r2f_hbf_int16_lm.ncu-rep
du -sh gpurun_out
