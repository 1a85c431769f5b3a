#!/bin/bash
# round 2, GPU call 18: speculative saturation in the lock-in tile kernels: parity (incl. saturating states) + rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nco.py tests/test_golden.py tests/test_gpu_cpp.py tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_rows.py --only "Lockin|phase|Lowpass" --out gpurun_out/r2c18_rows.json 2>&1 | grep GSa
timeout 200 python bench.py --workload lockin --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench lockin', d['value'], d['roofline']['frac'])"
