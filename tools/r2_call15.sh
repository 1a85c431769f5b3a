#!/bin/bash
# round 2, GPU call 15: batched state loads in the tiled HBF kernels: parity + rows at two call lengths
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_hbf.py tests/test_golden.py tests/test_gpu_float_edges.py tests/test_gpu_cpp.py -m gpu -x -q 2>&1 | tail -2
timeout 400 python tools/bench_rows.py --only "cascade f32|chain|TAPS_98" --out gpurun_out/r2c15_rows.json 2>&1 | grep GSa
timeout 300 python bench.py --workload hbf --steps 10 --profile 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench hbf', d['value'])"
timeout 300 python bench.py --workload chain --steps 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chain', d['value'], [(p['lanes_per_gpu'], round(p['GSa/s'],1)) for p in d['sweep']])"
