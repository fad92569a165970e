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


@pytest.mark.parametrize("hetero,xchg", [(0, 1), (1, 1), (0, 2), (0, 0)])
@pytest.mark.parametrize("world", [2])
def test_sharded_solve_matches_single_gpu(world, hetero, xchg):
    """xchg: 1 one-shot / 2 two-shot band exchange through peer memory, 0 = no peer memory at all (NCCL all-reduces)."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1", BA_MGPU_HETERO=str(hetero),
                                  **({"BA_B200_BAND_XCHG": str(xchg)} if xchg else {"BA_B200_NO_PEER_EXCHANGE": "1"})))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU_OK" in out.stdout


def test_two_devices_in_one_process():
    """Kernel attributes (dynamic shared-memory opt-ins) apply per device: a solver on device 1 created after one on
    device 0 in the SAME process must launch its 200 KB kernels too (the opt-in used to be a process-wide
    flag).  Both solve the same problem; results agree to the last bits that the by-point reds allow."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    sys.path.insert(0, ROOT)
    from bundle_adjustment_solver_b200 import capi, scenes
    from bundle_adjustment_solver_b200 import solver as S
    sc = scenes.scene_trajectory(150, 12_000, 8, stereo=True, seed=5, n_fixed=2)     # tile build + partitioned solve
    costs = []
    for dev in (0, 1, 0):
        e = S.load_scene(S.FullBundleAdjustmentSolver(device=dev), sc)
        summ = S.Summary()
        e.solve(capi.default_options(max_num_iterations=8, threshold_cost_change=0.0, threshold_step_size=0.0), summ)
        costs.append([i.cost for i in summ.optimization_info_list])
        pb = scenes.scene_poseonly_batch(n_frames=32, n_points=200, seed=2)
        out = S.PoseOnlyBundleAdjustmentSolver(device=dev).solve_batched(
            pb.kind, pb.offsets, pb.points, pb.px_left, pb.px_right, pb.intr_left, pb.intr_right, pb.poses_init,
            capi.PoseOnlyOptions(1e-6, 1e-6, 1.5, 2.5, 100), left_to_right=pb.left_to_right)
        assert np.abs(out["poses"] - pb.poses_true).max() < 1e-3
    assert len(costs[0]) == len(costs[1]) == 8
    np.testing.assert_allclose(costs[0], costs[1], rtol=1e-12)
    np.testing.assert_allclose(costs[0], costs[2], rtol=1e-12)
