#!/bin/bash
# round 2, GPU call 7: full-circle cossin table in the lock-in / cossin kernels: parity, rows, WPC sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nco.py tests/test_golden.py tests/test_gpu_cpp.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/bench_rows.py --only "Lockin|cossin|phase" --out gpurun_out/r2c7_rows.json 2>&1 | grep GSa
timeout 200 python bench.py --workload lockin --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench lockin', d['value'], d['roofline']['frac'])"
if [ -f idsp_b200/variants/tune.so ]; then
for w in 1 2 4 8; do
  echo "== IDSP_LOCKIN_WPC=$w"; IDSP_B200_LIB=$PWD/idsp_b200/variants/tune.so IDSP_LOCKIN_WPC=$w IDSP_SWEEP_ONLY=1 timeout 120 python tools/sweep_lockin.py 2>&1 | head -2
done
fi
