#!/bin/bash
# round 2, GPU call 10: ncu --set full captures of the dominant kernels + launch list of the default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, launch-skip, command...
  local name=$1 k=$2 s=$3; shift 3
  timeout 600 $NCU -k "regex:$k" -s $s -c 1 -o gpurun_out/r2_$name "$@" > gpurun_out/r2_$name.log 2>&1
  echo "$name rc=$? $(ls -la gpurun_out/r2_$name.ncu-rep 2>/dev/null | awk '{print $5}')"
  python tools/summarize_ncu.py gpurun_out/r2_$name.ncu-rep gpurun_out/r2_${name}_ncu.md > /dev/null 2>&1
}
drop() { rm -f gpurun_out/r2_$1.ncu-rep; }  # gpurun_out/ comes back only below 64 MiB: keep the summaries
cap hbf_int16_lm hbf_int_fast_kernel 1 python tools/bench_rows.py --only "a14 HbfInt x16 cascade f32 lane-major" --reps 2 --out gpurun_out/x.json
cap hbf_dec16_lm hbf_dec_fast_kernel 1 python tools/bench_rows.py --only "a13 HbfDec /16 cascade f32 lane-major" --reps 2 --out gpurun_out/x.json
cap hbf_dec16_fm hbf_dec_fast_kernel 1 python tools/bench_rows.py --only "a13 HbfDec /16 cascade f32 frame-major" --reps 2 --out gpurun_out/x.json
drop hbf_dec16_fm
cap chain_int_bq hbf_int_fast_kernel 1 python tools/bench_rows.py --only "cfg5 chain" --reps 2 --out gpurun_out/x.json
cap lockin_fm tma_lanes_kernel 1 python bench.py --workload lockin --steps 2 --warmup 1 --profile
cap biquad_fm tma_lanes_kernel 1 python bench.py --steps 2 --warmup 1 --profile --no-extra
drop biquad_fm
cap cascade4_fm tma_lanes_kernel 1 python tools/bench_rows.py --only "a7 Cascade<4> i32 frame-major" --reps 2 --out gpurun_out/x.json
drop cascade4_fm
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_default.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2_launches_bench_default.csv)"
du -sh gpurun_out; ls -la gpurun_out | tail -24
