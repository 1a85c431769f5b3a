#!/bin/bash
# round 2, GPU call 29: tensor-map tiles for frame-major /4 .. /32 and x4 .. x32
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hbf.py tests/test_golden.py tests/test_gpu_float_edges.py tests/test_gpu_cpp.py -m gpu -x -q 2>&1 | tail -6
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hbf.py -m gpu -x -q -k "tensor_map" 2>&1 | tail -2
for v in tma ldgsts; do
  if [ $v = ldgsts ]; then export IDSP_HBF_FM_LDGSTS=1; else unset IDSP_HBF_FM_LDGSTS; fi
  echo "== $v"
  timeout 300 python tools/bench_rows.py --only "Hbf(Int x|Dec /)(4|8) cascade.*f32 frame-major" --out gpurun_out/x.json 2>&1 | grep GSa
done
