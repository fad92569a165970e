"""Reader / writer of the "Bundle Adjustment in the Large" text format (the on-disk form of the Venice data set that
BASELINE config C5 is shaped after; the reference has no on-disk format at all -- SURVEY.md 8f rank 4).

File layout:
    <num_cameras> <num_points> <num_observations>
    <camera_index> <point_index> <x> <y>            x num_observations
    <camera>: 9 values, one per line: Rodrigues rotation (3), translation (3), focal f, radial k1, k2
    <point>:  3 values, one per line
Camera model of the format:  P = R X + t,  p = -P / P.z,  pixel = f (1 + k1 |p|^2 + k2 |p|^4) p.

The solver's model (core/full_bundle_adjustment_solver.cpp:744-760) is an ideal pinhole u = fx x / z + cx looking down
+z with the intrinsics of a camera RIG shared by all poses.  The loader therefore
  * turns every camera frame by pi about x (D = diag(1, -1, -1): P' = D P looks down +z) and flips the sign of v,
  * removes the radial distortion from the observed pixels (fixed-point inversion of the radial polynomial),
  * registers ONE rig camera with the median focal length f0 and rescales the pixels of camera c by f0 / f_c, which
    is exact for the geometry and weights the residuals of camera c by f0 / f_c relative to the file's pixel units.
Poses are returned camera-to-world like the rest of this package (what AddPose receives)."""
import numpy as np

from .scenes import FullScene, inv_T, make_T


def _rodrigues(w):
    th = np.linalg.norm(w, axis=1)
    K = np.zeros((len(w), 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -w[:, 2], w[:, 1], w[:, 2], -w[:, 0], -w[:, 1], w[:, 0]
    small = th < 1e-12
    ths = np.where(small, 1.0, th)
    a = np.where(small, 1.0, np.sin(ths) / ths)
    b = np.where(small, 0.5, (1.0 - np.cos(ths)) / (ths * ths))
    return np.eye(3)[None] + a[:, None, None] * K + b[:, None, None] * (K @ K)


def _rotvec(R):
    from scipy.spatial.transform import Rotation
    return Rotation.from_matrix(R).as_rotvec()


def load_bal(path, n_fixed=2):
    """-> scenes.FullScene (mono rig, camera id 0).  The first `n_fixed` poses are held fixed (gauge)."""
    with open(path) as f:
        tok = f.read().split()
    nc, npt, nobs = int(tok[0]), int(tok[1]), int(tok[2])
    obs = np.asarray(tok[3:3 + 4 * nobs], dtype=np.float64).reshape(nobs, 4)
    off = 3 + 4 * nobs
    cams = np.asarray(tok[off:off + 9 * nc], dtype=np.float64).reshape(nc, 9)
    off += 9 * nc
    pts = np.asarray(tok[off:off + 3 * npt], dtype=np.float64).reshape(npt, 3)
    if len(tok) != off + 3 * npt:
        raise ValueError(f"{path}: {len(tok)} values, expected {off + 3 * npt}")
    ci, pi = obs[:, 0].astype(np.int32), obs[:, 1].astype(np.int32)
    if ci.min() < 0 or ci.max() >= nc or pi.min() < 0 or pi.max() >= npt:
        raise ValueError(f"{path}: observation indices out of range")
    f_c, k1, k2 = cams[:, 6], cams[:, 7], cams[:, 8]
    # undistort: pixel / f = r(|p|) p  ->  p
    d = obs[:, 2:4] / f_c[ci, None]
    p = d.copy()
    for _ in range(20):
        r2 = (p * p).sum(axis=1)
        p = d / (1.0 + k1[ci] * r2 + k2[ci] * r2 * r2)[:, None]
    f0 = float(np.median(f_c))
    uv = np.empty((nobs, 2))
    uv[:, 0] = f0 * p[:, 0]          # u = -f x / z_file = f x' / z'
    uv[:, 1] = -f0 * p[:, 1]         # v flips with the frame
    D = np.diag([1.0, -1.0, -1.0])
    R = D[None] @ _rodrigues(cams[:, :3])
    t = cams[:, 3:6] @ D.T
    T_cw = np.tile(np.eye(4), (nc, 1, 1))
    T_cw[:, :3, :3], T_cw[:, :3, 3] = R, t
    T_wc = inv_T(T_cw)
    return FullScene(cam_ids=[0], cam_intr=np.array([[f0, f0, 0.0, 0.0]]), cam_T=np.eye(4)[None].copy(),
                     poses_true=T_wc.copy(), poses_init=T_wc, fixed_poses=np.arange(min(n_fixed, nc)),
                     points_true=pts.copy(), points_init=pts, fixed_points=np.zeros(0, dtype=np.int64),
                     obs_cam=np.zeros(nobs, dtype=np.int32), obs_pose=ci, obs_point=pi, obs_uv=uv,
                     name=f"BAL {path}", meta=dict(bal_focal=f_c, bal_k1=k1, bal_k2=k2, f0=f0))


def save_bal(sc, path, use_init=True):
    """Writes a mono FullScene (one rig camera, identity extrinsic, cx = cy = 0) in the format, k1 = k2 = 0."""
    if len(sc.cam_ids) != 1 or not np.allclose(sc.cam_T[0], np.eye(4)):
        raise ValueError("the format holds one pinhole camera per image: mono scenes only")
    fx, fy, cx, cy = sc.cam_intr[0]
    if fx != fy:
        raise ValueError("the format has a single focal length per camera")
    T_wc = sc.poses_init if use_init else sc.poses_true
    X = sc.points_init if use_init else sc.points_true
    T_cw = inv_T(T_wc)
    D = np.diag([1.0, -1.0, -1.0])
    R = D[None] @ T_cw[:, :3, :3]                   # back to the file's frame (D is its own inverse)
    t = T_cw[:, :3, 3] @ D.T
    w = _rotvec(R)
    with open(path, "w") as f:
        f.write(f"{len(T_wc)} {len(X)} {sc.n_obs}\n")
        u = sc.obs_uv[:, 0] - cx
        v = -(sc.obs_uv[:, 1] - cy)
        for a, b, x, y in zip(sc.obs_pose, sc.obs_point, u, v):
            f.write(f"{a} {b} {x:.17g} {y:.17g}\n")
        for k in range(len(T_wc)):
            for val in (*w[k], *t[k], fx, 0.0, 0.0):
                f.write(f"{val:.17g}\n")
        for p in X:
            for val in p:
                f.write(f"{val:.17g}\n")
