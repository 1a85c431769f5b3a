#!/bin/bash
# round 2, GPU call 24: HBF decimator with 256-sample tiles (more CTAs per SM)
cd "$(dirname "$0")/.."
for v in tt256m4 tt256m5 tt256m6; do
  export IDSP_B200_LIB=$PWD/idsp_b200/variants/$v.so
  echo "== $v"
  timeout 600 python -m pytest tests/test_gpu_hbf.py -m gpu -x -q -k "dec" 2>&1 | tail -1
  timeout 300 python bench.py --workload hbf --steps 10 --profile 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench hbf', d['value'])"
  timeout 300 python tools/bench_rows.py --only "HbfDec /(4|8|16|32) cascade f32 lane-major|chain" --out gpurun_out/x.json 2>&1 | grep GSa
done
