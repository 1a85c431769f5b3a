#!/bin/bash
# round 2, GPU call 32 (8 GPUs): the default bench at N = 8 on the final kernels (frame-major HBF leg included)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err
echo "bench n8 rc=$?"; tail -c 200 gpurun_out/r2_bench_n8_final.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n8_final.json").read().strip().splitlines()[-1])
    print(d.get("value"), d["roofline"]["frac"], "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("frac_of_pcie"))
    for k, v in (d.get("extra") or {}).items():
        if "error" in v: print("  ", k, "ERROR", v["error"]); continue
        print("  ", k, v.get("value"), (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"), (v.get("pipelined_scatter_compute_store") or {}).get("value"))
except Exception as e:
    print("unparsable", e)
PY
