# usage: bash profiles/geometry_variants.sh "<nvcc flags>" ...   (each variant: C2 x4 runs + stress)
for v in "$@"; do
  echo "=== variant: $v"
  SEIR_NVCC_EXTRA="$v" python covid19-demography_b200/build.py --force > /dev/null 2>&1 || echo "BUILD FAILED"
  python profiles/run_case.py c2 4 2>&1 | tail -3
  python profiles/run_case.py stress 3 2>&1 | tail -1
done
