"""GPU, >= 2 devices: spawns tests/dist_gpu_worker.py under torchrun (skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("extra", [[], ["301", "177", "mag"], ["24", "12", "tet"], ["301", "77", "persist"]])
@pytest.mark.parametrize("world", [2])
def test_distributed_solve_matches_single_gpu(world, extra):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    if extra and extra[-1] == "persist":       # the persistent PCG kernel (default from 4 ranks up) on 2 ranks
        env["FE_B200_PERSIST"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_gpu_worker.py")] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and "DIST-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
