#!/bin/bash
# round 2, GPU call 2 (2 GPUs): multi-GPU tests (C-ABI communicator, NCCL scatter/gather, peer stores),
# bench at N = 1 and N = 2 with the sharded lock-in leg, reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -s > gpurun_out/r2c2_pytest_dist.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c2_pytest_dist.log
tail -8 gpurun_out/r2c2_pytest_dist.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c2_bench_n1.json 2> gpurun_out/r2c2_bench_n1.err
echo "bench n1 rc=$?"; tail -c 600 gpurun_out/r2c2_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c2_bench_n2.json 2> gpurun_out/r2c2_bench_n2.err
echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/r2c2_bench_n2.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/r2c2_bench_ref.json 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2c2_bench_n1.json", "gpurun_out/r2c2_bench_n2.json", "gpurun_out/r2c2_bench_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, d.get("value"), "e2e", d.get("e2e", {}).get("value"), "frac_pcie", d.get("e2e", {}).get("frac_of_pcie"), "pcie", d.get("e2e", {}).get("pcie_gbs"))
    r = d.get("roofline", {})
    print("  roofline", r.get("frac"), "sustained", (r.get("sustained") or {}).get("frac"), "cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("one_core"))
    for k, v in (d.get("extra") or {}).items():
        if "error" in v: print("  ", k, "ERROR", v["error"]); continue
        print("  ", k, v.get("value"), (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"))
        if k == "lockin_sharded": print("     ", json.dumps({q: v.get(q) for q in ("resident_GSa/s", "ms", "nvlink_GBs", "fused_peer_store")}))
        if k == "chain_f32": print("     ", [(p["lanes_per_gpu"], round(p["GSa/s"], 1)) for p in v["sweep"]])
PY
