#!/bin/bash
# tools/build_variants.sh name=DEF1,DEF2 ...  -> idsp_b200/variants/<name>.so (parallel nvcc builds)
cd "$(dirname "$0")/.."
mkdir -p idsp_b200/variants
for spec in "$@"; do
  name="${spec%%=*}"; defs="${spec#*=}"
  ( IDSP_DEFS="$defs" python -m idsp_b200.build --out="$PWD/idsp_b200/variants/$name.so" > /tmp/build_$name.log 2>&1 || echo "FAILED $name" ) &
done
wait
ls -la idsp_b200/variants/
