#!/bin/bash
# round 2, GPU call 17: chain tile shapes after the deferred store wait (tune build) + ncu of the default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export IDSP_B200_LIB=$PWD/idsp_b200/variants/tune.so
for w in 0 1 2; do
echo "== IDSP_CHAIN_WIDE=$w"
IDSP_CHAIN_WIDE=$w timeout 300 python bench.py --workload chain --steps 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chain', d['value'], [(p['lanes_per_gpu'], round(p['GSa/s'],1)) for p in d['sweep'][:5]])"
done
unset IDSP_B200_LIB
ncu --set full --clock-control none --import-source on -f -k regex:hbf_int_fast_kernel -s 1 -c 1 -o gpurun_out/r2_chain_deferred python tools/bench_rows.py --only "cfg5 chain" --reps 2 --out gpurun_out/x.json > gpurun_out/r2_chain_deferred.log 2>&1
python tools/summarize_ncu.py gpurun_out/r2_chain_deferred.ncu-rep gpurun_out/r2_chain_int_bq_deferred_ncu.md
