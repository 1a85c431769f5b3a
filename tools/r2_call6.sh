#!/bin/bash
# round 2, GPU call 6 (8 GPUs): the default bench at N = 8 (headline + extras incl. the sharded lock-in data plane on
# 1 048 576 lanes), the multi-GPU tests, and the host-streaming chunk size under 8-way PCIe / host-memory contention.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c6_topo.txt 2>&1
lscpu | head -25 >> gpurun_out/r2c6_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2c6_bench_n8.json 2> gpurun_out/r2c6_bench_n8.err
echo "bench n8 rc=$?"; tail -c 800 gpurun_out/r2c6_bench_n8.err
for mb in 8 64 128; do
  IDSP_HOST_CHUNK_MB=$mb timeout 300 $TR --master-port 2953$((mb % 10)) bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/r2c6_e2e_chunk$mb.json 2> gpurun_out/r2c6_e2e_chunk$mb.err
  echo "chunk $mb rc=$?"
done
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -s -k "c_abi or root_buffer" > gpurun_out/r2c6_pytest_dist.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c6_pytest_dist.log; tail -6 gpurun_out/r2c6_pytest_dist.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c6_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    e = d.get("e2e", {})
    print(f, round(d.get("value", 0), 1), "e2e", e.get("value"), "frac_pcie", e.get("frac_of_pcie"), "pcie", (e.get("pcie_gbs") or {}).get("h2d"), (e.get("pcie_gbs") or {}).get("d2h"))
    for k, v in (d.get("extra") or {}).items():
        if "error" in v: print("  ", k, "ERROR", v["error"]); continue
        print("  ", k, v.get("value"), (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"))
        if k == "lockin_sharded": print("     ", json.dumps({q: v.get(q) for q in ("resident_GSa/s", "ms", "nvlink_GBs", "fused_peer_store", "pipelined_scatter_compute_store")}))
        if k == "chain_f32": print("     ", [(p["lanes_per_gpu"], round(p["GSa/s"], 1)) for p in v["sweep"]])
PY
