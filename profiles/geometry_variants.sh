for v in "" "-DSEIR_THREADS=1024 -DSEIR_MIN_BLOCKS=1" "-DSEIR_THREADS=256 -DSEIR_MIN_BLOCKS=4" "-DSEIR_THREADS=384 -DSEIR_MIN_BLOCKS=2" "-DSEIR_THREADS=768 -DSEIR_MIN_BLOCKS=1"; do
  echo "=== variant: $v"
  SEIR_NVCC_EXTRA="$v" python covid19-demography_b200/build.py --force > /dev/null 2>&1
  python profiles/run_case.py c2 4 2>&1 | tail -3
  python profiles/run_case.py stress 3 2>&1 | tail -1
done
