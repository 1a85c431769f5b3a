#!/bin/bash
# tools/run_variants.sh [packed|scalar]:name ...   (on the GPU box) bench + parity of each variant .so
cd "$(dirname "$0")/.."
for spec in "$@"; do
  var="${spec%%:*}"; name="${spec#*:}"
  export IDSP_HBF_VARIANT=$var IDSP_B200_LIB=$PWD/idsp_b200/variants/$name.so
  t=$(python -m pytest tests/test_gpu_hbf.py -m gpu -x -q 2>&1 | tail -1)
  b=$(python bench.py --workload hbf --profile --steps 24 2>&1 | tail -1)
  echo "$var $name | $t | $b"
done
