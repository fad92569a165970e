"""ctypes bindings for the CPU oracle (oracle/ba_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from the product
package.  PARITY UNPINNED: see the header of ba_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libba_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "ba_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return _LIB_PATH


class FullOptions(C.Structure):
    """Mirrors Options (core/solver_option_and_summary.h:55-71); floats stay float."""
    _fields_ = [
        ("solver_type", C.c_int),
        ("threshold_step_size", C.c_float),
        ("threshold_cost_change", C.c_float),
        ("threshold_huber_loss", C.c_float),
        ("threshold_outlier_rejection", C.c_float),
        ("max_num_iterations", C.c_int),
        ("initial_lambda", C.c_float),
        ("decrease_ratio_lambda", C.c_float),
        ("increase_ratio_lambda", C.c_float),
        ("b_accumulate", C.c_int),
        ("method", C.c_int),   # 0 LM (Solve), 1 Gauss-Newton / 2 gradient descent of the refactor class
    ]


def default_full_options(**kw):
    o = FullOptions(1, 1e-5, 1e-5, 1.0, 2.0, 50, 100.0, 0.33, 3.0, 0, 0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class IterInfo(C.Structure):
    _fields_ = [
        ("cost", C.c_double),
        ("cost_change", C.c_double),
        ("average_reprojection_error", C.c_double),
        ("abs_gradient", C.c_double),
        ("abs_step", C.c_double),
        ("damping_term", C.c_double),
        ("iter_time", C.c_double),
        ("iteration_status", C.c_int),
        ("_pad", C.c_int),
    ]


class PoseOnlyOptions(C.Structure):
    _fields_ = [
        ("threshold_step_size", C.c_float),
        ("threshold_cost_change", C.c_float),
        ("threshold_huber_loss", C.c_float),
        ("threshold_outlier_rejection", C.c_float),
        ("max_num_iterations", C.c_int),
    ]


class PoseOnlyResult(C.Structure):
    _fields_ = [
        ("n_iterations", C.c_int),
        ("converged", C.c_int),
        ("success", C.c_int),
        ("n_summary", C.c_int),
        ("final_error", C.c_float),
        ("final_step", C.c_float),
    ]


_lib = None


def build_native():
    """The TIMING build of the oracle: the reference's own flags (CMakeLists.txt:6, -O2 -march=native, FMA contraction
    on).  -march=native binds the library to the CPU it was compiled on, so it is built where it runs (a temporary
    directory keyed by the CPU's flags), never shipped.  Only bench.py's CPU legs use it; parity tests use the
    portable build, whose arithmetic is the plain IEEE sequence."""
    import hashlib
    import tempfile
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except Exception:
        flags = "unknown"
    src = os.path.join(_HERE, "ba_oracle.cpp")
    key = hashlib.sha1((flags + str(os.path.getmtime(src))).encode()).hexdigest()[:12]
    out = os.path.join(tempfile.gettempdir(), f"libba_oracle_native_{key}.so")
    if not os.path.exists(out):
        tmp = out + f".{os.getpid()}"
        subprocess.check_call([os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-march=native", "-fPIC", "-shared",
                               "-o", tmp, src])
        os.replace(tmp, out)
    return out


_native = None


def lib(native=False):
    global _lib, _native
    if native:
        if _native is None:
            _native = _declare(C.CDLL(build_native()))
        return _native
    if _lib is None:
        build()
        _lib = _declare(C.CDLL(_LIB_PATH))
    return _lib


def _declare(L):
    L.orc_full_create.restype = C.c_void_p
    L.orc_full_destroy.argtypes = [C.c_void_p]
    L.orc_full_add_camera.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 4 + [C.c_void_p]
    L.orc_full_add_poses.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_full_add_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_full_make_pose_fixed.argtypes = [C.c_void_p, C.c_int]
    L.orc_full_make_point_fixed.argtypes = [C.c_void_p, C.c_int]
    L.orc_full_add_observation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    L.orc_full_add_observations.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 4
    L.orc_full_add_observations.restype = C.c_longlong
    L.orc_full_solve.argtypes = [C.c_void_p, C.POINTER(FullOptions), C.c_void_p, C.c_int]
    L.orc_full_converged.argtypes = [C.c_void_p]
    L.orc_full_initial_cost.argtypes = [C.c_void_p]
    L.orc_full_initial_cost.restype = C.c_double
    L.orc_full_get_pose.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_full_get_point.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_full_get_internal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_full_sizes.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_full_dump.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_full_dump.restype = C.c_longlong
    L.orc_full_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_full_opt_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_full_build_only.argtypes = [C.c_void_p, C.c_float, C.c_double, C.c_int, C.c_int]
    L.orc_full_build_only.restype = C.c_double
    L.orc_full_cost.argtypes = [C.c_void_p]
    L.orc_full_cost.restype = C.c_double
    L.orc_ldlt_solve_f64.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_ldlt_solve_f32.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_se3_exp_f64.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_poseonly_solve.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 11 + [
        C.POINTER(PoseOnlyOptions), C.POINTER(PoseOnlyResult)] + [C.c_void_p] * 3
    L.orc_poseonly_solve_batched.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 12 + [
        C.POINTER(PoseOnlyOptions), C.c_void_p]
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def T44_colmajor(R, t):
    """(R 3x3, t 3) -> 16 doubles, Eigen::Transform::data() order."""
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return np.ascontiguousarray(T.T.reshape(-1))


DUMP = dict(A=0, a=1, C=2, b=3, Cinv=4, B=5, S=6, rhs=7, x=8, y=9, scalars=10)


class FullBAOracle:
    """Integer-id mirror of FullBundleAdjustmentSolver (core/full_bundle_adjustment_solver.h:127-146).

    Poses are user-facing camera-to-world transforms (4x4), points are 3-vectors,
    exactly what the reference's AddPose/AddPoint receive; ids are insertion indices.
    """

    def __init__(self, native=False):
        self.L = lib(native)
        self.h = C.c_void_p(self.L.orc_full_create())
        self.n_poses = 0
        self.n_points = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_full_destroy(self.h)
            self.h = None

    def add_camera(self, cam_id, fx, fy, cx, cy, T_this_to_cam0_4x4):
        T = np.ascontiguousarray(np.asarray(T_this_to_cam0_4x4, dtype=np.float64).T.reshape(-1))
        self.L.orc_full_add_camera(self.h, cam_id, fx, fy, cx, cy, _p(T))

    def add_poses(self, T_wc):  # (n,4,4) row-major numpy
        T = np.ascontiguousarray(np.asarray(T_wc, dtype=np.float64).transpose(0, 2, 1).reshape(-1))
        self.L.orc_full_add_poses(self.h, len(T_wc), _p(T))
        self.n_poses += len(T_wc)

    def add_points(self, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        self.L.orc_full_add_points(self.h, len(X), _p(X))
        self.n_points += len(X)

    def make_pose_fixed(self, j):
        self.L.orc_full_make_pose_fixed(self.h, int(j))

    def make_point_fixed(self, i):
        self.L.orc_full_make_point_fixed(self.h, int(i))

    def add_observations(self, cam, pose, point, uv):
        cam = np.ascontiguousarray(cam, dtype=np.int32)
        pose = np.ascontiguousarray(pose, dtype=np.int32)
        point = np.ascontiguousarray(point, dtype=np.int32)
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        return self.L.orc_full_add_observations(self.h, len(cam), _p(cam), _p(pose), _p(point), _p(uv))

    def solve(self, options):
        cap = max(1, options.max_num_iterations)
        infos = (IterInfo * cap)()
        n = self.L.orc_full_solve(self.h, C.byref(options), infos, cap)
        return [infos[k] for k in range(min(n, cap))], bool(self.L.orc_full_converged(self.h))

    def initial_cost(self):
        return self.L.orc_full_initial_cost(self.h)

    def cost(self):
        return self.L.orc_full_cost(self.h)

    def build_only(self, thres_huber=1.0, lam=100.0, b_accumulate=0, do_solve=False):
        self.L.orc_full_build_only(self.h, thres_huber, lam, b_accumulate, int(do_solve))

    def sizes(self):
        out = np.zeros(6, dtype=np.int64)
        self.L.orc_full_sizes(self.h, _p(out))
        return dict(N=int(out[0]), M=int(out[1]), P=int(out[2]), n_obs=int(out[3]),
                    N_total=int(out[4]), M_total=int(out[5]))

    def dump(self, name):
        n = self.L.orc_full_dump(self.h, DUMP[name], None)
        buf = np.zeros(n, dtype=np.float64)
        self.L.orc_full_dump(self.h, DUMP[name], _p(buf))
        return buf

    def pairs(self):
        P = self.sizes()["P"]
        pj = np.zeros(P, dtype=np.int32)
        pi = np.zeros(P, dtype=np.int32)
        self.L.orc_full_pairs(self.h, _p(pj), _p(pi))
        return pj, pi

    def opt_ids(self):
        s = self.sizes()
        oj = np.zeros(s["N"], dtype=np.int32)
        oi = np.zeros(s["M"], dtype=np.int32)
        self.L.orc_full_opt_ids(self.h, _p(oj), _p(oi))
        return oj, oi

    def get_poses(self):
        out = np.zeros((self.n_poses, 16))
        for j in range(self.n_poses):
            self.L.orc_full_get_pose(self.h, j, _p(out[j]))
        return out.reshape(-1, 4, 4).transpose(0, 2, 1).copy()

    def get_points(self):
        out = np.zeros((self.n_points, 3))
        for i in range(self.n_points):
            self.L.orc_full_get_point(self.h, i, _p(out[i]))
        return out

    def get_internal(self):
        T = np.zeros((self.n_poses, 12))
        X = np.zeros((self.n_points, 3))
        self.L.orc_full_get_internal(self.h, _p(T), _p(X))
        return T, X


def pose12(R, t):
    return np.concatenate([np.asarray(R, dtype=np.float32).reshape(-1), np.asarray(t, dtype=np.float32).reshape(-1)])


def poseonly_solve(kind, Xw, pxl, pxr, intr_l, intr_r, pose_io, options, left_to_right=None,
                   base_to_camera=None, world_to_last=None, want_history=False, native=False):
    """kind: 0 mono-6dof, 1 stereo-6dof, 2 mono-planar3dof, 3 stereo-planar3dof.  Poses: 12 float32."""
    L = lib(native)
    Xw = np.ascontiguousarray(Xw, dtype=np.float32)
    pxl = np.ascontiguousarray(pxl, dtype=np.float32)
    pxr = None if pxr is None else np.ascontiguousarray(pxr, dtype=np.float32)
    n = len(Xw)
    intr_l = np.ascontiguousarray(intr_l, dtype=np.float32)
    intr_r = np.ascontiguousarray(intr_l if intr_r is None else intr_r, dtype=np.float32)
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    l2r, b2c, w2l = f(left_to_right), f(base_to_camera), f(world_to_last)
    pose = np.ascontiguousarray(pose_io, dtype=np.float32).copy()
    ml = np.zeros(n, dtype=np.uint8)
    mr = np.zeros(n, dtype=np.uint8)
    res = PoseOnlyResult()
    K = max(1, options.max_num_iterations)
    hc = np.zeros(K, dtype=np.float32) if want_history else None
    hs = np.zeros(K, dtype=np.float32) if want_history else None
    dbg = np.zeros(12 * K, dtype=np.float32) if want_history else None
    L.orc_poseonly_solve(kind, n, _p(Xw), _p(pxl), _p(pxr), _p(intr_l), _p(intr_r), _p(l2r), _p(b2c),
                         _p(w2l), _p(pose), _p(ml), _p(mr), C.byref(options), C.byref(res), _p(hc), _p(hs),
                         _p(dbg))
    out = dict(pose=pose, mask_left=ml.astype(bool), mask_right=mr.astype(bool), result=res)
    if want_history:
        out.update(hist_cost=hc[:res.n_summary], hist_step=hs[:res.n_summary],
                   debug_poses=dbg.reshape(K, 12)[:res.n_iterations])
    return out


def poseonly_solve_batched(kind, offsets, Xw, pxl, pxr, intr_l, intr_r, poses_io, options,
                           left_to_right=None, base_to_camera=None, world_to_last=None, native=False):
    L = lib(native)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    nf = len(offsets) - 1
    Xw = np.ascontiguousarray(Xw, dtype=np.float32)
    pxl = np.ascontiguousarray(pxl, dtype=np.float32)
    pxr = None if pxr is None else np.ascontiguousarray(pxr, dtype=np.float32)
    intr_l = np.ascontiguousarray(intr_l, dtype=np.float32)
    intr_r = np.ascontiguousarray(intr_l if intr_r is None else intr_r, dtype=np.float32)
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    l2r, b2c, w2l = f(left_to_right), f(base_to_camera), f(world_to_last)
    poses = np.ascontiguousarray(poses_io, dtype=np.float32).copy()
    n = len(Xw)
    ml = np.zeros(n, dtype=np.uint8)
    mr = np.zeros(n, dtype=np.uint8)
    results = (PoseOnlyResult * nf)()
    L.orc_poseonly_solve_batched(kind, nf, _p(offsets), _p(Xw), _p(pxl), _p(pxr), _p(intr_l), _p(intr_r),
                                 _p(l2r), _p(b2c), _p(w2l), _p(poses), _p(ml), _p(mr), C.byref(options),
                                 results)
    return dict(poses=poses, mask_left=ml.astype(bool), mask_right=mr.astype(bool), results=results)


GEOM_OPS = dict(se3_exp=(0, 6, 0, 12), se3_log=(1, 12, 0, 6), so3_exp=(2, 3, 0, 9), so3_log=(3, 9, 0, 3), q2r=(4, 4, 0, 9),
                r2q=(5, 9, 0, 4), rotvec2q=(6, 3, 0, 4), r2euler=(7, 9, 0, 3), a2r=(8, 3, 0, 9), inverse_se3=(9, 12, 0, 12),
                add_front_se3=(10, 6, 6, 6), q_mult=(11, 4, 4, 4))


def geometry(name, a, b=None, dtype=np.float64):
    """utility/geometry_library.cpp restated (geom_ref in ba_oracle.cpp): rows of `a` (and `b`) -> rows of the result."""
    op, si, s2, so = GEOM_OPS[name]
    a = np.ascontiguousarray(a, dtype=dtype).reshape(-1, si)
    b = None if not s2 else np.ascontiguousarray(b, dtype=dtype).reshape(-1, s2)
    out = np.zeros((len(a), so), dtype=dtype)
    f = lib().orc_geometry if dtype == np.float64 else lib().orc_geometry_f
    assert f(op, len(a), _p(a), _p(b), _p(out)) == 0
    return out
