#!/bin/bash
# The parity tests on a build with our own bounds checks (-DSEIR_CHECKS: every index that reaches a list, a record, a queue or a
# contact target is tested; a violation fails seir_run) and with the per-CTA queues shrunk 256-fold, so that their overflow paths
# (hits / contacts handled in place, direct appends to a remote owner's list) run on the small fixtures.  Run on a GPU box (2 GPUs:
# the sharded test is included); the product build is restored afterwards.
SEIR_NVCC_EXTRA="-DSEIR_CHECKS -DSEIR_QUEUE_DIV=256 -DSEIR_TRACE_LOGCAP=120" python covid19-demography_b200/build.py --force > /dev/null 2>&1 || { echo BUILD_FAILED; exit 1; }
python -c "
import sys; sys.path.insert(0, '.')
from covid19_demography_b200 import engine
print(engine.load_library().seir_build_info().decode())"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_boundary.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -3
python covid19-demography_b200/build.py --force > /dev/null 2>&1
