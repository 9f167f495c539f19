#!/bin/bash
# Build libseir_b200.so with different compile-time switches ON THE GPU BOX and time the C2 day loop of each.
#   bash profiles/variants.sh "<flags of variant 1>" "<flags of variant 2>" ...
out=gpurun_out/variants.log
: > $out
for v in "$@"; do
  echo "=== variant: [$v]" >> $out
  SEIR_NVCC_EXTRA="$v" python covid19-demography_b200/build.py --force > /dev/null 2>> $out || { echo "BUILD FAILED" >> $out; continue; }
  grep -A2 "persistent" covid19-demography_b200/build_ptxas.log | grep -E "Used|spill" >> $out
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2 >> $out
  python profiles/day_timeline.py > gpurun_out/variant_tl.log 2>&1
  head -1 gpurun_out/variant_tl.log >> $out
  sed -n '11p;31p;41p;51p;56p' gpurun_out/variant_tl.log >> $out
done
python covid19-demography_b200/build.py --force > /dev/null 2>&1
