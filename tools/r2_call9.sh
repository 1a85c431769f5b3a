#!/bin/bash
# round 2, GPU call 9: integer multiply issue rates + 8/16-bit SOS order check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools/ubench3 | tee gpurun_out/r2_ubench_int.log
timeout 900 python -m pytest tests/test_gpu_biquad.py tests/test_golden.py tests/test_gpu_fm_disc.py -m gpu -x -q 2>&1 | tail -2
timeout 400 python tools/bench_rows.py --only "Biquad DF1 i|Cascade|BiquadClamp DF1 i32" --out gpurun_out/r2c9_rows.json 2>&1 | grep GSa
