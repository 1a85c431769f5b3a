#!/bin/bash
# round 2, GPU call 13: chain tile shapes (8 / 16 / 32 lanes per CTA) with the fused feed-forward half
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in 1 2; do
echo "== IDSP_CHAIN_WIDE=$w"
IDSP_CHAIN_WIDE=$w timeout 600 python -m pytest tests/test_gpu_hbf.py -m gpu -x -q -k chain 2>&1 | tail -1
IDSP_CHAIN_WIDE=$w timeout 300 python bench.py --workload chain --steps 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chain', d['value'], [(p['lanes_per_gpu'], round(p['GSa/s'],1)) for p in d['sweep']])"
done
