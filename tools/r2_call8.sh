#!/bin/bash
# round 2, GPU call 8: integer SOS sums re-ordered (earliest operands first): parity + rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_biquad.py tests/test_golden.py tests/test_gpu_fm_disc.py -m gpu -x -q 2>&1 | tail -4
timeout 400 python tools/bench_rows.py --only "Biquad|Cascade|DirectForm|FM disc" --out gpurun_out/r2c8_rows.json 2>&1 | grep GSa
timeout 200 python bench.py --steps 20 --no-extra --profile 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench biquad', d['value'], d['roofline']['frac'])"
