"""Generates the golden fixtures under tests/golden/ from the CPU oracle (oracle/ba_oracle.cpp).

The reference holds no golden vectors for its solvers and cannot be compiled in the
build container, so these fixtures pin the ORACLE's behaviour on seeded scenes (regression + the
expected values the GPU tests compare against when the oracle is not re-run).  Re-run with
    python tests/golden/make_golden.py
after any deliberate change of the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from bundle_adjustment_solver_b200 import scenes  # noqa: E402
from bundle_adjustment_solver_b200 import solver as S  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def full_ba(seed, accum, max_it=300):
    sc = scenes.scene_test_ba(seed=seed)
    o = S.load_scene(oracle.FullBAOracle(), sc)
    o.build_only(thres_huber=1.0, lam=100.0, b_accumulate=accum, do_solve=True)
    blocks = {k: o.dump(k) for k in ("A", "a", "C", "b", "B", "rhs", "x", "y")}
    Sm = o.dump("S")
    o = S.load_scene(oracle.FullBAOracle(), sc)
    opt = oracle.default_full_options(max_num_iterations=max_it, threshold_cost_change=1e-6,
                                      threshold_step_size=1e-6, b_accumulate=accum)
    infos, conv = o.solve(opt)
    out = dict(
        initial_cost=o.initial_cost(), converged=conv, n_iterations=len(infos),
        cost=np.array([i.cost for i in infos]), lam=np.array([i.damping_term for i in infos]),
        status=np.array([i.iteration_status for i in infos]), step=np.array([i.abs_step for i in infos]),
        poses=o.get_poses(), points=o.get_points(),
        # block checksums of the first linearisation (lambda = 100): sums and Frobenius norms
        **{f"sum_{k}": v.sum() for k, v in blocks.items()},
        **{f"nrm_{k}": np.linalg.norm(v) for k, v in blocks.items()},
        sum_S=Sm.sum(), nrm_S=np.linalg.norm(Sm), x0=blocks["x"], a0=blocks["a"],
        sizes=np.array(list(o.sizes().values())))
    return out


def poseonly(kind_name):
    if kind_name == "stereo6":
        pb = scenes.scene_poseonly_batch(n_frames=8, n_points=300, seed=11, pixel_sigma=0.5, stereo=True,
                                         right_invalid_fraction=0.1)
    elif kind_name == "mono6":
        pb = scenes.scene_poseonly_batch(n_frames=8, n_points=300, seed=12, pixel_sigma=0.5, stereo=False)
    elif kind_name == "stereo3":
        pb = scenes.scene_poseonly_planar_batch(n_frames=8, n_points=300, seed=13, pixel_sigma=0.3, stereo=True)
    else:
        pb = scenes.scene_poseonly_planar_batch(n_frames=8, n_points=300, seed=14, pixel_sigma=0.3, stereo=False)
    ref = oracle.poseonly_solve_batched(pb.kind, pb.offsets, pb.points, pb.px_left, pb.px_right, pb.intr_left,
                                        pb.intr_right, pb.poses_init, oracle.PoseOnlyOptions(1e-6, 1e-6, 1.5, 2.5, 100),
                                        left_to_right=pb.left_to_right, base_to_camera=pb.base_to_camera,
                                        world_to_last=pb.world_to_last)
    return dict(poses=ref["poses"], n_iterations=np.array([r.n_iterations for r in ref["results"]]),
                converged=np.array([r.converged for r in ref["results"]]),
                final_error=np.array([r.final_error for r in ref["results"]]),
                mask_left=np.packbits(ref["mask_left"]), mask_right=np.packbits(ref["mask_right"]))


if __name__ == "__main__":
    for seed, accum in ((0, 0), (1, 0), (0, 1)):
        np.savez_compressed(os.path.join(HERE, f"full_ba_c1_seed{seed}_accum{accum}.npz"), **full_ba(seed, accum))
    for k in ("stereo6", "mono6", "stereo3", "mono3"):
        np.savez_compressed(os.path.join(HERE, f"poseonly_{k}.npz"), **poseonly(k))
    print("golden fixtures written to", HERE)
