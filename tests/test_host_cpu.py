"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol (no compute
without a GPU), it fails loudly without a device, the registration semantics mirror the reference,
landmark sharding is exact (world_size-2 gloo run), and the scene generators are deterministic."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle
from bundle_adjustment_solver_b200 import capi, scenes, sharding
from bundle_adjustment_solver_b200 import solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(engine_lib):
    hdr = open(os.path.join(ROOT, "include", "ba_b200.h")).read()
    declared = set(re.findall(r"\b(ba_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ba_last_error"} - set(capi.SYMBOLS)
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(engine_lib, s)]
    assert not missing, missing
    assert sorted(declared) == sorted(set(capi.SYMBOLS)), (sorted(declared ^ set(capi.SYMBOLS)))
    assert b"sm_100a" in engine_lib.ba_version()


def test_struct_layouts_match_header(tmp_path):
    # the ctypes mirrors against the C compiler's view of include/ba_b200.h
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ba_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu\\n", sizeof(ba_options), offsetof(ba_options, inverse_scaler),'
                   ' offsetof(ba_options, method), sizeof(ba_iter_info));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    size, off_inv, off_method, size_info = map(int, subprocess.check_output([str(exe)]).split())
    assert C.sizeof(capi.Options) == size == 64
    assert capi.Options.inverse_scaler.offset == off_inv and capi.Options.method.offset == off_method
    assert C.sizeof(capi.IterInfo) == size_info
    assert C.sizeof(capi.IterInfo) == 64
    assert C.sizeof(capi.PoseOnlyOptions) == 20 and C.sizeof(capi.PoseOnlyResult) == 24
    assert C.sizeof(oracle.IterInfo) == C.sizeof(capi.IterInfo)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure path")
def test_engine_fails_loudly_without_gpu(engine_lib):
    h = C.c_void_p()
    assert engine_lib.ba_create(C.byref(h), 0) == -2       # BA_ERR_CUDA, no silent CPU fallback
    with pytest.raises(capi.BaError):
        S.FullBundleAdjustmentSolver(device=0)
    pb = scenes.scene_poseonly_batch(n_frames=2, n_points=16, seed=0)
    with pytest.raises(capi.BaError):
        S.PoseOnlyBundleAdjustmentSolver().solve_batched(pb.kind, pb.offsets, pb.points, pb.px_left, pb.px_right,
                                                         pb.intr_left, pb.intr_right, pb.poses_init,
                                                         capi.PoseOnlyOptions(1e-6, 1e-6, 1.5, 2.5, 10),
                                                         left_to_right=pb.left_to_right)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bundle_adjustment_solver_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "ba_oracle" not in txt and "orc_" not in txt, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read().lower().replace("oracle's layout", "")


def test_scene_generators_are_deterministic():
    a, b = scenes.scene_test_ba(seed=2), scenes.scene_test_ba(seed=2)
    assert np.array_equal(a.obs_uv, b.obs_uv) and np.array_equal(a.points_init, b.points_init)
    c = scenes.scene_c3(seed=1, scale=0.02)
    d = scenes.scene_c3(seed=1, scale=0.02)
    assert np.array_equal(c.obs_uv, d.obs_uv)
    assert 0.8e6 * 0.02 < c.n_obs < 1.2e6 * 0.02
    e = scenes.scene_c5(seed=0, scale=0.01)
    assert len(e.cam_ids) == 1 and e.obs_cam.max() == 0
    deg = np.bincount(e.obs_point)
    assert deg.min() >= 2 and deg.max() > 3 * np.median(deg)   # heavy tail


def test_landmark_ranges_balance_and_cover():
    sc = scenes.scene_c3(seed=0, scale=0.05)
    for world in (1, 2, 3, 8):
        rg = sharding.landmark_ranges(sc.obs_point, len(sc.points_init), world)
        assert rg[0][0] == 0 and rg[-1][1] == len(sc.points_init)
        assert all(rg[k][1] == rg[k + 1][0] for k in range(world - 1))
        cnt = [np.count_nonzero((sc.obs_point >= lo) & (sc.obs_point < hi)) for lo, hi in rg]
        assert sum(cnt) == sc.n_obs
        assert max(cnt) - min(cnt) <= 0.05 * sc.n_obs / world + 64
    # ragged / empty edge cases
    assert sharding.landmark_ranges(np.zeros(0, dtype=np.int64), 0, 2) == [(0, 0), (0, 0)]
    assert sharding.landmark_ranges(np.array([0, 0, 0]), 1, 2)[-1][1] == 1


_GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["BA_ROOT"]); sys.path.insert(0, os.path.join(os.environ["BA_ROOT"], "tests"))
import oracle
from bundle_adjustment_solver_b200 import scenes, sharding
from bundle_adjustment_solver_b200 import solver as S
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sc = scenes.scene_trajectory(24, 300, 6, stereo=True, seed=4, n_fixed=2)
sh = sharding.shard_scene(sc, rank, world)
o = S.load_scene(oracle.FullBAOracle(), sh)
o.build_only(thres_huber=1.0, lam=10.0, b_accumulate=0, do_solve=False)
n = 6 * o.sizes()["N"]
buf = torch.from_numpy(np.concatenate([o.dump("S"), o.dump("rhs"), [o.cost(), float(o.sizes()["M"]), float(o.sizes()["n_obs"])]]))
dist.all_reduce(buf)                       # the engine's per-iteration exchange: [S | rhs] and LM scalars
if rank == 0:
    full = S.load_scene(oracle.FullBAOracle(), sc)
    full.build_only(thres_huber=1.0, lam=10.0, b_accumulate=0, do_solve=False)
    ref = np.concatenate([full.dump("S"), full.dump("rhs"), [full.cost(), float(full.sizes()["M"]), float(full.sizes()["n_obs"])]])
    err = np.abs(buf.numpy() - ref).max() / np.abs(ref).max()
    assert err < 1e-12, err
    # replicated reduced solve + local back-substitution == the single-rank solution
    x_full = np.linalg.solve(ref[:n * n].reshape(n, n), ref[n * n:n * n + n])
    x_sum = np.linalg.solve(buf.numpy()[:n * n].reshape(n, n), buf.numpy()[n * n:n * n + n])
    assert np.abs(x_full - x_sum).max() < 1e-9 * np.abs(x_full).max()
    print("GLOO_OK", err)
dist.barrier()
dist.destroy_process_group()
'''


def test_sharded_partial_systems_sum_to_the_full_system_gloo(tmp_path):
    """world_size-2 gloo run on CPU: every rank builds [S | rhs] from its landmark shard (all poses
    replicated); the all-reduced system equals the single-rank system (the N>1 path of the engine)."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, BA_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29500 + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_shard_scene_preserves_relative_insertion_order():
    sc = scenes.scene_trajectory(12, 80, 5, stereo=True, seed=9)
    parts = [sharding.shard_scene(sc, r, 3) for r in range(3)]
    assert sum(p.n_obs for p in parts) == sc.n_obs
    assert sum(len(p.points_init) for p in parts) == len(sc.points_init)
    lo, hi = parts[1].meta["landmark_range"]
    k = (sc.obs_point >= lo) & (sc.obs_point < hi)
    assert np.array_equal(parts[1].obs_uv, sc.obs_uv[k])       # same relative order => same last-writer pairs
    assert parts[1].obs_point.min() == 0


_GLOO_POSEONLY_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["BA_ROOT"])
import oracle
from bundle_adjustment_solver_b200 import scenes, sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
pb = scenes.scene_poseonly_batch(n_frames=37, n_points=60, seed=11, pixel_sigma=0.3, stereo=True, ragged=True)
sh = sharding.shard_poseonly_batch(pb, rank, world)
opt = oracle.PoseOnlyOptions(1e-6, 1e-6, 1.5, 2.5, 100)
mine = oracle.poseonly_solve_batched(sh.kind, sh.offsets, sh.points, sh.px_left, sh.px_right, sh.intr_left, sh.intr_right,
                                     sh.poses_init, opt, left_to_right=sh.left_to_right)
# frames are independent: no collective on the data path; the gather below only serves the check
lo, hi = sharding.frame_ranges(pb.n_frames, world)[rank]
buf = torch.zeros(pb.n_frames, 12, dtype=torch.float32)
buf[lo:hi] = torch.from_numpy(mine["poses"])
dist.all_reduce(buf)
if rank == 0:
    full = oracle.poseonly_solve_batched(pb.kind, pb.offsets, pb.points, pb.px_left, pb.px_right, pb.intr_left, pb.intr_right,
                                         pb.poses_init, opt, left_to_right=pb.left_to_right)
    assert np.array_equal(buf.numpy(), full["poses"])          # bit-identical: same frames, same arithmetic
    print("GLOO_POSEONLY_OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_poseonly_frame_sharding_gloo(tmp_path):
    """Config C2 over several GPUs = frames split evenly, no communication: world_size-2 gloo run
    of the split on the CPU oracle; the shards' poses put together are bit-identical to the unsharded batch."""
    assert sharding.frame_ranges(4096, 8) == [(512 * r, 512 * (r + 1)) for r in range(8)]
    assert sharding.frame_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharding.frame_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]          # more ranks than frames: empty shards
    script = tmp_path / "worker_po.py"
    script.write_text(_GLOO_POSEONLY_WORKER)
    env = dict(os.environ, BA_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 30500 + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "GLOO_POSEONLY_OK" in out.stdout


def load_oracle(sc):
    import oracle
    from bundle_adjustment_solver_b200 import solver as S
    return S.load_scene(oracle.FullBAOracle(), sc)


def test_bal_format_round_trip_and_camera_model(tmp_path):
    """bal.py: a mono scene written in the 'Bundle Adjustment in the Large' layout and read back is
    the same problem (same oracle cost); a file with radial distortion and per-image focal lengths projects, after
    the loader's undistortion / focal normalisation, exactly where the format's camera model says."""
    from bundle_adjustment_solver_b200 import bal
    sc = scenes.scene_trajectory(9, 60, 4, stereo=False, seed=2, n_fixed=2)
    sc.cam_intr[0, 2:] = 0.0                                 # the format has no principal point
    # regenerate the pixels for cx = cy = 0 through the oracle-independent projection of the scene generator
    T_cw = scenes.inv_T(sc.poses_true)
    Xc = np.einsum("nij,nj->ni", T_cw[sc.obs_pose, :3, :3], sc.points_true[sc.obs_point]) + T_cw[sc.obs_pose, :3, 3]
    f0 = sc.cam_intr[0, 0]
    sc.obs_uv = np.stack([f0 * Xc[:, 0] / Xc[:, 2], f0 * Xc[:, 1] / Xc[:, 2]], axis=1)
    p = tmp_path / "problem.txt"
    bal.save_bal(sc, p)
    back = bal.load_bal(p, n_fixed=2)
    assert back.n_obs == sc.n_obs and len(back.points_init) == len(sc.points_init)
    assert np.abs(back.poses_init - sc.poses_init).max() < 1e-12
    assert np.abs(back.obs_uv - sc.obs_uv).max() < 1e-9
    c0 = load_oracle(sc); c0.sizes()
    c1 = load_oracle(back); c1.sizes()
    assert abs(c0.cost() - c1.cost()) <= 1e-9 * c0.cost()
    # distortion + per-image focal: write the file by hand with the format's own model
    rng = np.random.default_rng(5)
    nc, npt = 3, 40
    w = rng.normal(size=(nc, 3)) * 0.1
    t = rng.normal(size=(nc, 3)) * 0.2 + [0, 0, -6.0]        # points end up in front of the -z looking cameras
    fk = np.stack([rng.uniform(900, 1400, nc), rng.uniform(-0.05, 0.0, nc), rng.uniform(0.0, 0.01, nc)], axis=1)
    X = rng.uniform(-1, 1, size=(npt, 3))
    from scipy.spatial.transform import Rotation
    rows = []
    for c in range(nc):
        P = Rotation.from_rotvec(w[c]).apply(X) + t[c]
        q = -P[:, :2] / P[:, 2:3]
        r2 = (q * q).sum(axis=1)
        px = fk[c, 0] * (1 + fk[c, 1] * r2 + fk[c, 2] * r2 * r2)[:, None] * q
        rows += [(c, i, px[i, 0], px[i, 1]) for i in range(npt)]
    with open(tmp_path / "dist.txt", "w") as fh:
        fh.write(f"{nc} {npt} {len(rows)}\n")
        fh.writelines(f"{a} {b} {x:.17g} {y:.17g}\n" for a, b, x, y in rows)
        fh.writelines(f"{v:.17g}\n" for c in range(nc) for v in (*w[c], *t[c], *fk[c]))
        fh.writelines(f"{v:.17g}\n" for v in X.reshape(-1))
    sd = bal.load_bal(tmp_path / "dist.txt")
    od = load_oracle(sd); od.sizes()
    assert od.cost() < 1e-9                                   # exact data: zero reprojection error after loading


def test_cpp_bal_loader_matches_python_loader(tmp_path):
    """include/ba_b200/utility/bal_loader.h against bal.py on a file with distortion and per-image focal lengths."""
    from bundle_adjustment_solver_b200 import bal
    rng = np.random.default_rng(8)
    nc, npt = 4, 25
    rows = [(c, i, *rng.uniform(-400, 400, 2)) for c in range(nc) for i in range(npt) if (c + i) % 3]
    cams = np.column_stack([rng.normal(size=(nc, 3)) * 0.3, rng.normal(size=(nc, 3)), rng.uniform(800, 1500, nc),
                            rng.uniform(-0.04, 0.0, nc), rng.uniform(0, 0.005, nc)])
    X = rng.normal(size=(npt, 3))
    path = tmp_path / "p.txt"
    with open(path, "w") as fh:
        fh.write(f"{nc} {npt} {len(rows)}\n")
        fh.writelines(f"{a} {b} {x:.17g} {y:.17g}\n" for a, b, x, y in rows)
        fh.writelines(f"{v:.17g}\n" for v in cams.reshape(-1))
        fh.writelines(f"{v:.17g}\n" for v in X.reshape(-1))
    exe = tmp_path / "bal"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-DBA_B200_FORCE_EIGEN_SHIM", "-o", str(exe),
                        os.path.join(ROOT, "tests", "cpp", "test_bal_loader.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([str(exe), str(path)], capture_output=True, text=True, check=True).stdout.splitlines()
    sc = bal.load_bal(path)
    assert out[0].split()[1:] == [str(nc), str(npt), str(len(rows))]
    assert abs(float(out[1].split()[1]) - sc.meta["f0"]) < 1e-12
    poses = np.array([[float(v) for v in ln.split()[1:]] for ln in out if ln.startswith("POSE")]).reshape(nc, 4, 4).transpose(0, 2, 1)
    assert np.abs(poses - sc.poses_init).max() < 1e-12
    obs = np.array([[float(v) for v in ln.split()[1:]] for ln in out if ln.startswith("OBS")])
    assert np.array_equal(obs[:, 0], sc.obs_pose) and np.array_equal(obs[:, 1], sc.obs_point)
    assert np.abs(obs[:, 2:] - sc.obs_uv).max() < 1e-9
    assert subprocess.run([str(exe), str(tmp_path / "missing.txt")], capture_output=True, text=True).returncode == 2


def test_bench_byte_accounting_moves_with_the_pose_side_forms(monkeypatch):
    """bench.py's algorithmic bytes per phase follow the engine's switches: with the speculative pose side the pass
    over the observations is charged to the trial-cost phase, with the storing reduce (banded plans) the linearize
    phase holds no kernel at all -- and no form counts the pass over the observations twice."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import importlib
    import bench
    sz = dict(n_obs=1_000_484, N_total=200, M_total=50_000, N=198, M=50_000, P=499_479)
    banded = dict(kernel="k_nd_persistent<14,11> (partitioned banded Cholesky)", bw=71, band_clear=False, alg_bytes=7.0e5)
    dense = dict(banded, kernel="k_chol_cluster (blocked Cholesky in one thread-block-cluster launch)")
    free_obs = 990_000
    out = {}
    for name, spec, stores, info in (("separate", "0", "1", banded), ("spec", "1", "0", banded),
                                     ("spec_stores", "1", "1", banded), ("spec_dense", "1", "1", dense)):
        monkeypatch.setenv("BA_B200_SPEC_LIN", spec)
        monkeypatch.setenv("BA_B200_REDUCE_STORES", stores)
        importlib.reload(bench)
        out[name] = bench.algorithmic_bytes(sz, free_obs, info)
    monkeypatch.delenv("BA_B200_SPEC_LIN"); monkeypatch.delenv("BA_B200_REDUCE_STORES")
    importlib.reload(bench)
    assert out["separate"]["linearize"] > 28 * free_obs and out["separate"]["update_cost"] >= 28 * sz["n_obs"]
    for k in ("spec", "spec_stores", "spec_dense"):
        assert out[k]["linearize"] < 28 * free_obs          # no pass over the observations in the linearize phase
        assert out[k]["update_cost"] > out["separate"]["update_cost"]
    assert out["spec_stores"]["linearize"] == 0 and out["spec_dense"]["linearize"] > 0
    assert out["spec_stores"]["schur"] > out["spec"]["schur"]          # the storing reduce writes the pose-side blocks
    total = lambda b: b["linearize"] + b["schur"] + b["backsub"] + b["update_cost"]
    assert total(out["spec_stores"]) < total(out["spec"]) < total(out["separate"])
