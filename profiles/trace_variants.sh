#!/bin/bash
# Build libseir_b200.so with different compile-time switches ON THE GPU BOX; parity of the tracing tests and the
# tracer's phase timeline (C2, tracing from day 37) of each.   bash profiles/trace_variants.sh "<flags 1>" "<flags 2>" ...
out=gpurun_out/trace_variants.log
: > $out
for v in "$@"; do
  echo "=== variant: [$v]" >> $out
  SEIR_NVCC_EXTRA="$v" python covid19-demography_b200/build.py --force > /dev/null 2>> $out || { echo "BUILD FAILED" >> $out; continue; }
  timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tracing or more_replicas" 2>&1 | tail -1 >> $out
  timeout 300 python profiles/trace_timeline.py > gpurun_out/trace_variant_tl.log 2>&1
  head -1 gpurun_out/trace_variant_tl.log >> $out
  sed -n '3p;9p;11p;15p;20p' gpurun_out/trace_variant_tl.log >> $out
done
python covid19-demography_b200/build.py --force > /dev/null 2>&1
