#!/bin/bash
# round 2, GPU call 19: high-word products that ptxas cannot fold into IMAD.HI + 64-bit addend pairs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nco.py tests/test_golden.py tests/test_gpu_cpp.py tests/test_gpu_fm_disc.py -m gpu -x -q 2>&1 | tail -2
for v in default mixonly; do
  echo "== $v"
  if [ $v = default ]; then unset IDSP_B200_LIB; else export IDSP_B200_LIB=$PWD/idsp_b200/variants/$v.so; fi
  timeout 300 python tools/bench_rows.py --only "Lockin|cossin|phase" --out gpurun_out/r2c19_rows_$v.json 2>&1 | grep GSa
  timeout 200 python bench.py --workload lockin --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench lockin', d['value'], d['roofline']['frac'])"
done
