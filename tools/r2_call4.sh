#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --workload lockin_sharded --steps 5 > gpurun_out/r2c4_sharded_n2.json 2> gpurun_out/r2c4_sharded_n2.err
echo "rc=$?"; tail -c 1200 gpurun_out/r2c4_sharded_n2.err; cat gpurun_out/r2c4_sharded_n2.json
timeout 300 python tools/bench_rows.py --only "i64|f64|FM disc|phase|Lockin" --out gpurun_out/r2c4_rows.json 2>&1 | grep GSa
