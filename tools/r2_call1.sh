#!/bin/bash
# round 2, GPU call 1: full -m gpu parity suite, rows table (policy 0 and generic-only for the 8-byte rows),
# lock-in tile-shape sweep.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
tail -5 gpurun_out/r2c1_pytest.log
timeout 600 python tools/bench_rows.py --out gpurun_out/r2c1_rows.json > gpurun_out/r2c1_rows.log 2>&1
timeout 300 python tools/bench_rows.py --policy 1 --only "i64|f64|FM disc|phase|Lockin" --out gpurun_out/r2c1_rows_generic.json > gpurun_out/r2c1_rows_generic.log 2>&1
IDSP_B200_LIB=$PWD/idsp_b200/variants/tune.so timeout 400 python tools/sweep_lockin.py > gpurun_out/r2c1_sweep_lockin.log 2>&1
IDSP_B200_LIB=$PWD/idsp_b200/variants/tune_dupk.so IDSP_SWEEP_ONLY=1 timeout 200 python tools/sweep_lockin.py > gpurun_out/r2c1_sweep_lockin_dupk.log 2>&1
grep -E "i64|f64|FM disc|Lockin|single stage|TAPS_98|/16|chain" gpurun_out/r2c1_rows.log | head -40
cat gpurun_out/r2c1_rows_generic.log | grep GSa | head -20
cat gpurun_out/r2c1_sweep_lockin.log
head -4 gpurun_out/r2c1_sweep_lockin_dupk.log
