#!/bin/bash
# round 2, GPU call 3: tile-shape sweep of the 8-byte-sample TMA ops, occupancy experiment on the tiled HBF /16
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export IDSP_B200_LIB=$PWD/idsp_b200/variants/tune.so
for cfg in x 0 1 3 4 5 6 8 9 10 11; do
  if [ "$cfg" = "x" ]; then unset IDSP_OUT8_CFG; else export IDSP_OUT8_CFG=$cfg; fi
  echo "== IDSP_OUT8_CFG=$cfg"
  timeout 120 python tools/bench_rows.py --quick --only "frame-major" --out /tmp/rows_$cfg.json 2>&1 | grep -E "DF1 i64|DF1 f64|FM disc|phase"
done > gpurun_out/r2c3_sweep_out8.log 2>&1
unset IDSP_OUT8_CFG
for extra in 0 20000 60000; do
  echo "== IDSP_HBF_EXTRA_SMEM=$extra"
  IDSP_HBF_EXTRA_SMEM=$extra timeout 200 python bench.py --workload hbf --profile --steps 16 2>&1 | tail -1
done > gpurun_out/r2c3_hbf_occupancy.log 2>&1
cat gpurun_out/r2c3_sweep_out8.log gpurun_out/r2c3_hbf_occupancy.log
