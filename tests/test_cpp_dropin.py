"""The drop-in C++ classes (include/ba_b200/core/*.h) compiled with g++ against libba_b200.so:
registration semantics on the CPU, and -- on the GPU -- the reference's own call sequence
(test/test_ba.cpp:235-297) checked against the oracle on the same scene."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle
from bundle_adjustment_solver_b200 import capi, scenes
from helpers import load_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "bundle_adjustment_solver_b200")


def _compile(src, out):
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include", "ba_b200"), "-o", str(out),
           os.path.join(ROOT, "tests", "cpp", src), "-L", LIBDIR, "-lba_b200", f"-Wl,-rpath,{LIBDIR}"]
    subprocess.check_call(cmd)
    return str(out)


def test_registration_semantics_cpu(tmp_path, engine_lib):
    exe = _compile("test_dropin_registration.cpp", tmp_path / "reg")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "REGISTRATION_OK" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]
    assert "Invalid camera index." in out.stderr and "Nonexisting pose." in out.stderr and "Nonexisting point." in out.stderr
    assert "Cannot enroll parameter" in out.stderr and "Empty pointer is conveyed" in out.stderr


def test_refactor_registration_semantics_cpu(tmp_path, engine_lib):
    """FullBundleAdjustmentSolverRefactor throws where the reference does (refactor.cpp:99-153, 223-229, 851)."""
    exe = _compile("test_ba_refactor_dropin.cpp", tmp_path / "refac")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "REFACTOR_REGISTRATION_OK" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]
    assert "existing camera" in out.stderr


def _write_scene(sc, path):
    with open(path, "wb") as f:
        f.write(struct.pack("<5i", len(sc.cam_ids), len(sc.poses_init), len(sc.points_init), len(sc.fixed_poses), sc.n_obs))
        for k, cid in enumerate(sc.cam_ids):
            f.write(struct.pack("<i", cid))
            f.write(np.asarray(sc.cam_intr[k], dtype="<f8").tobytes())
            f.write(np.ascontiguousarray(sc.cam_T[k].T, dtype="<f8").tobytes())
        for T in sc.poses_init:
            f.write(np.ascontiguousarray(T.T, dtype="<f8").tobytes())
        f.write(np.ascontiguousarray(sc.points_init, dtype="<f8").tobytes())
        f.write(np.asarray(sc.fixed_poses, dtype="<i4").tobytes())
        rec = np.zeros(sc.n_obs, dtype=[("c", "<i4"), ("j", "<i4"), ("i", "<i4"), ("u", "<f8"), ("v", "<f8")])
        rec["c"], rec["j"], rec["i"] = sc.obs_cam, sc.obs_pose, sc.obs_point
        rec["u"], rec["v"] = sc.obs_uv[:, 0], sc.obs_uv[:, 1]
        f.write(rec.tobytes())


@pytest.mark.gpu
@pytest.mark.parametrize("accum", [0, 1])
def test_cpp_dropin_full_ba_matches_oracle(tmp_path, accum, engine_lib):
    sc = scenes.scene_test_ba(seed=6)
    exe = _compile("test_ba_dropin.cpp", tmp_path / "ba")
    _write_scene(sc, tmp_path / "scene.bin")
    out = subprocess.run([exe, str(tmp_path / "scene.bin"), str(tmp_path / "res.bin"), "300", str(accum)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "Analytic Solver Report" in out.stdout and "Bundle Adjustment Statistics" in out.stdout
    raw = open(tmp_path / "res.bin", "rb").read()
    n_it, conv = struct.unpack_from("<2i", raw, 0)
    costs = np.frombuffer(raw, dtype="<f8", count=n_it, offset=8)
    off = 8 + 8 * n_it
    poses = np.frombuffer(raw, dtype="<f8", count=16 * len(sc.poses_init), offset=off).reshape(-1, 4, 4).transpose(0, 2, 1)
    pts = np.frombuffer(raw, dtype="<f8", count=3 * len(sc.points_init), offset=off + 128 * len(sc.poses_init)).reshape(-1, 3)
    o = load_oracle(sc)
    infos, conv_o = o.solve(oracle.default_full_options(max_num_iterations=300, threshold_cost_change=1e-6,
                                                        threshold_step_size=1e-6, b_accumulate=accum))
    assert bool(conv) == conv_o and abs(n_it - len(infos)) <= 1
    assert abs(costs[-1] - infos[-1].cost) <= 1e-6 * abs(infos[-1].cost)
    assert np.abs(poses - o.get_poses()).max() < 1e-6 and np.abs(pts - o.get_points()).max() < 1e-6
    assert np.array_equal(poses[:5], sc.poses_init[:5])     # fixed poses untouched


@pytest.mark.gpu
@pytest.mark.parametrize("mode,method,iters", [("lm", 0, 300), ("gn", 1, 30), ("gd", 2, 20)])
def test_cpp_refactor_dropin_matches_oracle(tmp_path, mode, method, iters, engine_lib):
    """The reference's test/test_ba_refactor.cpp call sequence through the refactored front-end, LM / Gauss-Newton /
    gradient descent, against the oracle on the same scene."""
    sc = scenes.scene_test_ba(seed=8)
    exe = _compile("test_ba_refactor_dropin.cpp", tmp_path / "refac")
    _write_scene(sc, tmp_path / "scene.bin")
    out = subprocess.run([exe, str(tmp_path / "scene.bin"), str(tmp_path / "res.bin"), str(iters), mode],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    raw = open(tmp_path / "res.bin", "rb").read()
    n_it, conv = struct.unpack_from("<2i", raw, 0)
    costs = np.frombuffer(raw, dtype="<f8", count=n_it, offset=8)
    off = 8 + 8 * n_it
    poses = np.frombuffer(raw, dtype="<f8", count=16 * len(sc.poses_init), offset=off).reshape(-1, 4, 4).transpose(0, 2, 1)
    pts = np.frombuffer(raw, dtype="<f8", count=3 * len(sc.points_init), offset=off + 128 * len(sc.poses_init)).reshape(-1, 3)
    o = load_oracle(sc)
    infos, conv_o = o.solve(oracle.default_full_options(max_num_iterations=iters, threshold_cost_change=1e-6,
                                                        threshold_step_size=1e-6, method=method))
    assert bool(conv) == conv_o and abs(n_it - len(infos)) <= 1
    assert abs(costs[-1] - infos[-1].cost) <= 1e-6 * abs(infos[-1].cost)
    assert np.abs(poses - o.get_poses()).max() < 1e-6 and np.abs(pts - o.get_points()).max() < 1e-6


@pytest.mark.gpu
def test_cpp_dropin_poseonly(tmp_path, engine_lib):
    exe = _compile("test_poseonly_dropin.cpp", tmp_path / "po")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "POSEONLY_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
