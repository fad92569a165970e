"""utility/geometry_library.cpp: the oracle's restatement against independent SciPy / NumPy
derivations, the host drop-in header against the oracle (CPU), and the batched device entry point
ba_geometry_batched against the oracle (GPU).  Tolerances: double 1e-12 absolute, float 2e-5."""
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg
from scipy.spatial.transform import Rotation as Rot

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(seed=0, n=400):
    rng = np.random.default_rng(seed)
    w = rng.normal(size=(n, 3)) * 0.9
    nrm = np.linalg.norm(w, axis=1, keepdims=True)
    w = np.where(nrm > 2.9, w * 2.9 / nrm, w)          # the logarithm returns the rotation angle in [0, pi]
    w[:5] *= 1e-12                                     # the series branches (theta < 1e-9 / trace test)
    w[5] = [np.pi - 1e-3, 0, 0]                        # close to the pi singularity of the logarithm
    xi = np.concatenate([rng.normal(size=(n, 3)), w], axis=1)
    return rng, w, xi


def test_oracle_restatement_against_scipy():
    rng, w, xi = _inputs()
    R = oracle.geometry("so3_exp", w).reshape(-1, 3, 3)
    assert np.abs(R - Rot.from_rotvec(w).as_matrix()).max() < 1e-13
    big = (np.linalg.norm(w, axis=1) > 1e-4) & (np.linalg.norm(w, axis=1) < 3.0)   # below its trace threshold the reference returns 0
    assert np.abs(oracle.geometry("so3_log", R)[big] - w[big]).max() < 1e-9
    assert np.abs(oracle.geometry("so3_log", R)[:5]).max() == 0.0
    T = oracle.geometry("se3_exp", xi)
    for k in range(0, len(xi), 7):
        M = np.zeros((4, 4))
        M[:3, :3] = [[0, -xi[k, 5], xi[k, 4]], [xi[k, 5], 0, -xi[k, 3]], [-xi[k, 4], xi[k, 3], 0]]
        M[:3, 3] = xi[k, :3]
        E = scipy.linalg.expm(M)
        assert np.abs(E[:3, :3] - T[k, :9].reshape(3, 3)).max() < 1e-12 and np.abs(E[:3, 3] - T[k, 9:]).max() < 1e-12
    assert np.abs(oracle.geometry("se3_log", T)[big] - xi[big]).max() < 1e-8
    q = oracle.geometry("r2q", R)
    assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < 1e-12
    assert np.abs(oracle.geometry("q2r", q).reshape(-1, 3, 3) - R).max() < 1e-12
    qs = Rot.from_rotvec(w[big]).as_quat()[:, [3, 0, 1, 2]]
    qo = oracle.geometry("rotvec2q", w[big])
    assert np.abs(qo * np.sign(qo[:, :1]) - qs * np.sign(qs[:, :1])).max() < 1e-12
    rpy = rng.uniform(-1.2, 1.2, size=(100, 3))
    Ra = oracle.geometry("a2r", rpy)
    assert np.abs(Ra.reshape(-1, 3, 3) - Rot.from_euler("xyz", rpy).as_matrix()).max() < 1e-13
    assert np.abs(oracle.geometry("r2euler", Ra) - rpy).max() < 1e-12
    Ti = oracle.geometry("inverse_se3", T)
    for k in range(0, len(T), 11):
        A, B = np.eye(4), np.eye(4)
        A[:3, :3], A[:3, 3] = T[k, :9].reshape(3, 3), T[k, 9:]
        B[:3, :3], B[:3, 3] = Ti[k, :9].reshape(3, 3), Ti[k, 9:]
        assert np.abs(A @ B - np.eye(4)).max() < 1e-12
    q1, q2 = oracle.geometry("rotvec2q", w[big][:50]), oracle.geometry("rotvec2q", w[big][50:100])
    qm = oracle.geometry("q_mult", q1, q2)
    Rm = oracle.geometry("q2r", qm).reshape(-1, 3, 3)
    assert np.abs(Rm - oracle.geometry("q2r", q1).reshape(-1, 3, 3) @ oracle.geometry("q2r", q2).reshape(-1, 3, 3)).max() < 1e-12


def test_host_dropin_header_matches_oracle(tmp_path):
    exe = tmp_path / "geom"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-DBA_B200_FORCE_EIGEN_SHIM", "-o", str(exe),
                        os.path.join(ROOT, "tests", "cpp", "test_geometry_dropin.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = {ln.split()[0]: np.array([float(v) for v in ln.split()[1:]]) for ln in out.splitlines()}
    xi = np.array([[0.3, -0.2, 0.5, 0.4, -0.7, 0.25]])
    w = xi[:, 3:]
    cm4 = lambda T12: np.vstack([np.c_[T12[:9].reshape(3, 3), T12[9:]], [0, 0, 0, 1]]).T.reshape(-1)   # column-major 4 x 4
    cm3 = lambda R9: R9.reshape(3, 3).T.reshape(-1)
    T = oracle.geometry("se3_exp", xi)[0]
    R = oracle.geometry("so3_exp", w)[0]
    tol = 1e-14
    assert np.abs(got["se3Exp"] - cm4(T)).max() < tol
    assert np.abs(got["SE3Log"] - oracle.geometry("se3_log", T[None])[0]).max() < tol
    assert np.abs(got["so3Exp"] - cm3(R)).max() < tol
    assert np.abs(got["SO3Log"] - oracle.geometry("so3_log", R[None])[0]).max() < tol
    q = oracle.geometry("r2q", R[None])[0]
    assert np.abs(got["r2q"] - q).max() < tol
    assert np.abs(got["q2r"] - cm3(oracle.geometry("q2r", q[None])[0])).max() < tol
    assert np.abs(got["rotvec2q"] - oracle.geometry("rotvec2q", w)[0]).max() < tol
    Ra = oracle.geometry("a2r", [[0.1, -0.4, 0.9]])[0]
    assert np.abs(got["a2r"] - cm3(Ra)).max() < tol
    assert np.abs(got["r2euler"] - oracle.geometry("r2euler", Ra[None])[0]).max() < tol
    assert np.abs(got["inverseSE3"] - cm4(oracle.geometry("inverse_se3", T[None])[0])).max() < tol
    assert np.abs(got["addFrontse3"] - oracle.geometry("add_front_se3", xi, [[0.01, 0.02, -0.03, 0.05, 0.02, -0.04]])[0]).max() < tol
    q1, q2 = oracle.geometry("rotvec2q", w), oracle.geometry("rotvec2q", [[-0.2, 0.1, 0.6]])
    qm = oracle.geometry("q_mult", q1, q2)[0]
    assert np.abs(got["q1_mult_q2"] - qm).max() < tol
    # q_left_mult(q1) q2 == q1 * q2 == q_right_mult(q2) q1  (utility/geometry_library.cpp:23-59)
    assert np.abs(got["q_left_mult"].reshape(4, 4).T @ q2[0] - qm).max() < 1e-14
    assert np.abs(got["q_conj"] - q1[0] * [1, -1, -1, -1]).max() == 0
    assert np.array_equal(got["skewMat"].reshape(3, 3).T, [[0, -3, 2], [3, 0, -1], [-2, 1, 0]])
    assert np.abs(got["so3Exp_f"] - cm3(R)).max() < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_device_batched_geometry_matches_oracle(dtype, tol):
    from bundle_adjustment_solver_b200 import capi
    L = capi.lib()
    f = L.ba_geometry_batched if dtype == np.float64 else L.ba_geometry_batched_f
    rng, w, xi = _inputs(seed=3, n=5000)
    T = oracle.geometry("se3_exp", xi)
    R = oracle.geometry("so3_exp", w)
    q = oracle.geometry("r2q", R)
    rpy = rng.uniform(-1.2, 1.2, size=(len(w), 3))
    cases = {"se3_exp": (xi, None), "se3_log": (T, None), "so3_exp": (w, None), "so3_log": (R, None), "q2r": (q, None),
             "r2q": (R, None), "rotvec2q": (w, None), "r2euler": (R, None), "a2r": (rpy, None), "inverse_se3": (T, None),
             "add_front_se3": (xi, 0.05 * rng.normal(size=xi.shape)), "q_mult": (q, q[::-1].copy())}
    for name, (a, b) in cases.items():
        if dtype == np.float32 and name in ("se3_log", "so3_log", "add_front_se3"):
            keep = np.linalg.norm(w, axis=1) > 1e-2          # float: stay away from the trace threshold and from pi
            keep &= np.linalg.norm(w, axis=1) < 3.0
            a, b = a[keep], (None if b is None else b[keep])
        op, si, s2, so = oracle.GEOM_OPS[name]
        a_ = np.ascontiguousarray(a, dtype=dtype)
        b_ = None if b is None else np.ascontiguousarray(b, dtype=dtype)
        out = np.zeros((len(a_), so), dtype=dtype)
        assert f(0, op, len(a_), capi.ptr(a_), capi.ptr(b_), capi.ptr(out)) == 0
        ref = oracle.geometry(name, a_, b_, dtype=dtype)
        assert np.abs(out - ref).max() < tol, (name, np.abs(out - ref).max())
    # empty batch and bad op
    assert f(0, 0, 0, None, None, None) == 0
    assert f(0, 99, 1, capi.ptr(np.zeros(6, dtype=dtype)), None, capi.ptr(np.zeros(12, dtype=dtype))) != 0
