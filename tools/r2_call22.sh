#!/bin/bash
# round 2, GPU call 22 (8 GPUs): host-memory flavours under contention + the default bench at N = 8 on the final kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/pcie_probe.py 2>&1 | grep -E "^N=|unavailable" | tee gpurun_out/r2_pcie_probe_n8.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err
echo "bench n8 rc=$?"; tail -c 300 gpurun_out/r2_bench_n8_final.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n8_final.json").read().strip().splitlines()[-1])
    print(d.get("value"), "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("frac_of_pcie"), d.get("e2e", {}).get("pcie_gbs"))
    for k, v in (d.get("extra") or {}).items():
        if "error" in v: print("  ", k, "ERROR", v["error"]); continue
        print("  ", k, v.get("value"), (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"))
        if k == "lockin_sharded": print("     ", json.dumps({q: v.get(q) for q in ("resident_GSa/s", "ms", "nvlink_GBs", "fused_peer_store", "pipelined")})[:900])
except Exception as e:
    print("unparsable", e)
PY
