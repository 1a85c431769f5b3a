#!/bin/bash
# round 2, GPU call 25: frame-major /16 decimator with tensor-map input tiles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hbf.py tests/test_golden.py tests/test_gpu_float_edges.py tests/test_gpu_cpp.py -m gpu -x -q 2>&1 | tail -8
for v in tma ldgsts; do
  if [ $v = ldgsts ]; then export IDSP_HBF_FM_LDGSTS=1; else unset IDSP_HBF_FM_LDGSTS; fi
  echo "== $v"
  timeout 300 python bench.py --workload hbf --layout 0 --steps 10 --profile 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench hbf frame-major', d['value'], d.get('parity_check'))"
  timeout 300 python tools/bench_rows.py --only "HbfDec /16 cascade.*frame-major" --out gpurun_out/x.json 2>&1 | grep GSa
done
