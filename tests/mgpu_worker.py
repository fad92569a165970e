"""torchrun worker for the multi-GPU parity test: every rank holds all poses and its landmark shard,
joins the NCCL communicator through the C-ABI, and runs the LM loop; rank 0 also solves the unsharded
problem on its own GPU and compares."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bundle_adjustment_solver_b200 import capi, scenes, sharding  # noqa: E402
from bundle_adjustment_solver_b200 import solver as S  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L = capi.lib()
    sc = scenes.scene_trajectory(60, 3000, 8, stereo=True, seed=21, n_fixed=2)
    if os.environ.get("BA_MGPU_HETERO") == "1":
        # short tracks in the first half of the landmarks, long ones in the second: the shards see different
        # co-visibility envelopes, so the ranks only agree on the banded plan through ba_comm_init
        first = np.full(len(sc.points_init), 10 ** 9, dtype=np.int64)
        np.minimum.at(first, sc.obs_point, sc.obs_pose)
        keep = (sc.obs_point >= len(sc.points_init) // 2) | (sc.obs_pose - first[sc.obs_point] < 3)
        sc.obs_cam, sc.obs_pose, sc.obs_point, sc.obs_uv = sc.obs_cam[keep], sc.obs_pose[keep], sc.obs_point[keep], sc.obs_uv[keep]
    sh = sharding.shard_scene(sc, rank, world)
    e = S.load_scene(S.FullBundleAdjustmentSolver(device=local), sh)
    e._upload()
    sz = e.sizes()
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = (C.c_ubyte * 128)()
        assert L.ba_comm_get_unique_id(raw) == 0
        idbuf = torch.tensor(list(raw), dtype=torch.uint8)
    idbuf = idbuf.to(dev)
    dist.broadcast(idbuf, 0)
    tot = torch.tensor([sz["M"], sz["n_obs"]], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
    rc = L.ba_comm_init(e.h, raw, rank, world, int(tot[0]), int(tot[1]))
    assert rc == 0, L.ba_last_error(e.h)
    opt = capi.default_options(max_num_iterations=8, threshold_cost_change=1e-6, threshold_step_size=1e-6)
    summ = S.Summary()
    e.solve(opt, summ)
    costs = np.array([i.cost for i in summ.optimization_info_list])
    lams = np.array([i.damping_term for i in summ.optimization_info_list])
    poses = e.get_poses()
    # every rank must hold identical poses and LM history (replicated reduced solve)
    buf = torch.from_numpy(np.concatenate([costs, lams, poses.reshape(-1)])).to(dev)
    ref = buf.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(buf, ref), "ranks diverged"
    if rank == 0:
        full = S.load_scene(S.FullBundleAdjustmentSolver(device=local), sc)
        s1 = S.Summary()
        full.solve(capi.default_options(max_num_iterations=8, threshold_cost_change=1e-6, threshold_step_size=1e-6), s1)
        c1 = np.array([i.cost for i in s1.optimization_info_list])
        assert len(c1) == len(costs), (len(c1), len(costs))
        np.testing.assert_allclose(costs, c1, rtol=1e-8)
        np.testing.assert_allclose(poses, full.get_poses(), atol=1e-8)
        lo, hi = sh.meta["landmark_range"]
        np.testing.assert_allclose(e.get_points(), full.get_points()[lo:hi], atol=1e-8)
    # a second, larger problem in the same process ATTACHES to the communicator (no second NCCL set-up); its band
    # needs a larger exchange buffer than the first one's (the old set is retired, not freed) and its reduced system
    # goes through the partitioned solve
    sc2 = scenes.scene_trajectory(150, 9000, 8, stereo=True, seed=22, n_fixed=2)
    sh2 = sharding.shard_scene(sc2, rank, world)
    e2 = S.load_scene(S.FullBundleAdjustmentSolver(device=local), sh2)
    e2._upload()
    sz2 = e2.sizes()
    tot2 = torch.tensor([sz2["M"], sz2["n_obs"]], dtype=torch.int64, device=dev)
    dist.all_reduce(tot2)
    rc = L.ba_comm_attach(e2.h, int(tot2[0]), int(tot2[1]))
    assert rc == 0, L.ba_last_error(e2.h)
    s2 = S.Summary()
    e2.solve(capi.default_options(max_num_iterations=6, threshold_cost_change=0.0, threshold_step_size=0.0), s2)
    costs2 = np.array([i.cost for i in s2.optimization_info_list])
    buf2 = torch.from_numpy(np.concatenate([costs2, e2.get_poses().reshape(-1)])).to(dev)
    ref2 = buf2.clone()
    dist.broadcast(ref2, 0)
    assert torch.equal(buf2, ref2), "ranks diverged on the attached solver"
    # the first solver still works after the exchange buffers were replaced
    e.solve(capi.default_options(max_num_iterations=2, threshold_cost_change=0.0, threshold_step_size=0.0), S.Summary())
    if rank == 0:
        full2 = S.load_scene(S.FullBundleAdjustmentSolver(device=local), sc2)
        s3 = S.Summary()
        full2.solve(capi.default_options(max_num_iterations=6, threshold_cost_change=0.0, threshold_step_size=0.0), s3)
        np.testing.assert_allclose(costs2, [i.cost for i in s3.optimization_info_list], rtol=1e-8)
        np.testing.assert_allclose(e2.get_poses(), full2.get_poses(), atol=1e-8)
        print("MGPU_OK", world, costs[-1], costs2[-1])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
