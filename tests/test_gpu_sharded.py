"""-m gpu : ONE population sharded by whole households over 2 GPUs (BASELINE.json config C3)
against the CPU oracle, bit-exact.  The ranks are real processes on real devices (a persistent kernel of one rank
waits for the other's at the day barrier: they must run at the same time), so the test spawns
`python -m torch.distributed.run` with two ranks and needs a box with at least two GPUs; on a one-GPU box it is
skipped with that reason (run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_run_equals_the_oracle_on_two_gpus(gpu_engine_module):
    n_dev = gpu_engine_module.load_library().seir_device_count()
    if n_dev < 2:
        pytest.skip("needs 2 GPUs on one box (found %d): the sharded path exchanges through peer memory" % n_dev)
    world = 4 if n_dev >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "sharded_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.split("\n") if l.startswith("sharded x")]
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert len(lines) == 8 and all("bit-exact vs oracle" in l for l in lines), out.stdout[-3000:]
