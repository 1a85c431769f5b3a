#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c5_pytest.log; tail -4 gpurun_out/r2c5_pytest.log
timeout 400 python tools/bench_rows.py --only "Hbf|chain" --out gpurun_out/r2c5_rows_hbf.json > gpurun_out/r2c5_rows_hbf.log 2>&1; grep GSa gpurun_out/r2c5_rows_hbf.log
timeout 200 python bench.py --workload hbf --layout 0 --profile --steps 16 2>&1 | tail -1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_hbf.py tests/test_gpu_nco.py tests/test_gpu_biquad.py -m gpu -x -q -k "caller_taps or misaligned or phase_stream or lo_stream or 8byte or chain_host or fm_disc" > gpurun_out/r2c5_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2c5_memcheck.log; tail -6 gpurun_out/r2c5_memcheck.log
