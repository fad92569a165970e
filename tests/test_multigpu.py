"""Landmark-sharded multi-GPU LM loop (NCCL all-reduce of [S | rhs] inside the engine -- of the band only when the
reduced system is banded) against the single-GPU solve; hetero = shards with different co-visibility envelopes.  Needs >= 2 GPUs (gpurun --gpus 2); the CPU-side sharding logic is covered by the gloo
test in test_host_cpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("hetero", [0, 1])
@pytest.mark.parametrize("world", [2])
def test_sharded_solve_matches_single_gpu(world, hetero):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1", BA_MGPU_HETERO=str(hetero)))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU_OK" in out.stdout
