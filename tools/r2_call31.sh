#!/bin/bash
# round 2, GPU call 31: lane-major decimator input tiles as tensor-map boxes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hbf.py tests/test_golden.py tests/test_gpu_float_edges.py tests/test_gpu_cpp.py -m gpu -x -q 2>&1 | tail -6
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hbf.py -m gpu -x -q -k "dec_cascade_tiled_kernel_streaming" 2>&1 | tail -2
for v in tma bulk; do
  if [ $v = bulk ]; then export IDSP_HBF_LM_BULK=1; else unset IDSP_HBF_LM_BULK; fi
  echo "== $v"
  timeout 300 python bench.py --workload hbf --steps 10 --profile 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench hbf lane-major', d['value'])"
  timeout 300 python tools/bench_rows.py --only "HbfDec /(2|4|8|16|32) cascade.*f32 lane-major|chain" --out gpurun_out/x.json 2>&1 | grep GSa
done
