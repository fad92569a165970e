"""Seeded synthetic scenes for the BASELINE.json configs (numpy, host side).

The reference seeds every test from std::random_device, so nothing in it is
reproducible; these generators follow the reference recipes with a fixed seed:
  * scene_test_ba        -- test/test_ba.cpp:53-232 (config C1)
  * scene_trajectory     -- configs C3/C4/C5 (SURVEY.md 8d "Synthetic inputs")
  * scene_poseonly_batch -- test/test_6dof_stereo_poseonly_ba.cpp:15-107 (config C2)
Poses are user-facing camera-to-world 4x4 transforms (what AddPose receives).
"""
from dataclasses import dataclass, field

import numpy as np


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def make_T(R, t):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def inv_T(T):
    """Isometry inverse (R^T, -R^T t); works on (...,4,4)."""
    T = np.asarray(T)
    R = T[..., :3, :3]
    t = T[..., :3, 3]
    out = np.zeros_like(T)
    Rt = np.swapaxes(R, -1, -2)
    out[..., :3, :3] = Rt
    out[..., :3, 3] = -np.einsum("...ij,...j->...i", Rt, t)
    out[..., 3, 3] = 1.0
    return out


@dataclass
class FullScene:
    cam_ids: list
    cam_intr: np.ndarray          # (n_cam, 4) fx fy cx cy
    cam_T: np.ndarray             # (n_cam, 4, 4) pose_this_to_cam0
    poses_true: np.ndarray        # (N, 4, 4) camera-to-world
    poses_init: np.ndarray
    fixed_poses: np.ndarray       # int ids
    points_true: np.ndarray       # (M, 3)
    points_init: np.ndarray
    fixed_points: np.ndarray
    obs_cam: np.ndarray           # int32, insertion order
    obs_pose: np.ndarray
    obs_point: np.ndarray
    obs_uv: np.ndarray            # (n_obs, 2) float64
    name: str = ""
    meta: dict = field(default_factory=dict)

    @property
    def n_obs(self):
        return len(self.obs_cam)


def _stereo_rig(fx, fy, cx, cy, baseline):
    """test_ba.cpp:79-98: cam1.pose_this_to_cam0 = (I, +baseline x)^-1."""
    intr = np.array([[fx, fy, cx, cy], [fx, fy, cx, cy]], dtype=np.float64)
    T0 = np.eye(4)
    T1 = inv_T(make_T(np.eye(3), [baseline, 0.0, 0.0]))
    return intr, np.stack([T0, T1])


def scene_test_ba(seed=0, pixel_sigma=0.0, num_total_poses=60, num_fixed_poses=5,
                  point_error_level=0.5, pose_translation_error_level=0.1):
    """Config C1.  Follows test/test_ba.cpp line by line (float32 where it uses float)."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    # GenerateWorldPosition (:53-77): float loop counters
    pts = []
    z = f32(1.7)
    while z <= f32(5.7):
        y = f32(0.0)
        while y <= f32(26.0):
            pts.append([float(f32(8.5)), float(y), float(z)])
            y = f32(y + f32(0.4))
        z = f32(z + f32(0.4))
    points_true = np.array(pts, dtype=np.float64)
    M = len(points_true)
    intr, cam_T = _stereo_rig(525.0, 525.0, 320.0, 240.0, 0.12)
    # camera poses (:132-171)
    base_to_camera = make_T(rot_y(np.pi / 2) @ rot_z(-np.pi / 2), [0, 0, 0])
    R_wb = rot_z(-0.1)
    t_wb = np.array([-4.0, -2.5, 0.0])
    x_step, y_step, yaw_step = float(f32(0.005)), float(f32(0.2)), float(f32(0.005))
    poses_true = []
    for _ in range(num_total_poses):
        R_wb = R_wb @ rot_z(yaw_step)
        t_wb = t_wb + np.array([x_step, y_step, 0.0])
        poses_true.append(make_T(R_wb, t_wb) @ base_to_camera)
    poses_true = np.array(poses_true)
    poses_init = poses_true.copy()
    lvl = pose_translation_error_level
    for j in range(num_fixed_poses, num_total_poses):  # (:174-178)
        poses_init[j, :3, 3] += rng.uniform(-lvl, lvl, 3).astype(f32).astype(np.float64)
    # projections with the TRUE poses (:180-232); inverse depth is a float in the reference
    cams, poses, points, uvs = [], [], [], []
    for j in range(num_total_poses):
        T_cw = inv_T(poses_true[j])
        local = points_true @ T_cw[:3, :3].T + T_cw[:3, 3]
        per_cam = []
        for c in range(2):
            Xc = local @ cam_T[c][:3, :3].T + cam_T[c][:3, 3]
            invz = (1.0 / Xc[:, 2]).astype(f32).astype(np.float64)
            noise = rng.normal(0, pixel_sigma, (M, 2)) if pixel_sigma > 0 else np.zeros((M, 2))
            u = intr[c, 0] * Xc[:, 0] * invz + intr[c, 2] + noise[:, 0]
            v = intr[c, 1] * Xc[:, 1] * invz + intr[c, 3] + noise[:, 1]
            seen = (u < 640) & (u > 0) & (v < 480) & (v > 0)
            ids = np.nonzero(seen)[0]
            per_cam.append((ids, u[ids], v[ids]))
        for c in range(2):  # AddObservation order (:254-274): left list then right list per frame
            ids, u, v = per_cam[c]
            cams.append(np.full(len(ids), c)); poses.append(np.full(len(ids), j)); points.append(ids)
            uvs.append(np.stack([u, v], axis=1))
    lv = point_error_level
    points_init = points_true + rng.uniform(-lv, lv, (M, 3)).astype(f32).astype(np.float64)
    return FullScene(
        cam_ids=[0, 1], cam_intr=intr, cam_T=cam_T, poses_true=poses_true, poses_init=poses_init,
        fixed_poses=np.arange(num_fixed_poses), points_true=points_true, points_init=points_init,
        fixed_points=np.zeros(0, dtype=np.int64),
        obs_cam=np.concatenate(cams).astype(np.int32), obs_pose=np.concatenate(poses).astype(np.int32),
        obs_point=np.concatenate(points).astype(np.int32), obs_uv=np.concatenate(uvs).astype(np.float64),
        name="C1_test_ba", meta=dict(seed=seed, pixel_sigma=pixel_sigma))


def scene_trajectory(n_poses, n_points, mean_track, stereo=True, seed=0, n_fixed=2, pixel_sigma=0.0,
                     heavy_tail=False, loop_fraction=0.0, point_error_level=0.5,
                     pose_translation_error_level=0.1, name="traj", pose_noise_seed=None):
    """Configs C3/C4/C5: poses on a smooth trajectory, each landmark seen by a run of
    consecutive poses (in both cameras when stereo).  heavy_tail draws the track length from a
    geometric distribution clipped to [2, 100] (Venice-like); loop_fraction adds far-away
    revisits so the reduced camera system is not purely banded.  Vectorised."""
    rng = np.random.default_rng(seed)
    step = 0.2
    j = np.arange(n_poses)
    # smooth trajectory: forward along world x, gentle lateral sway and yaw/pitch wiggle
    t = np.stack([step * j, 0.3 * np.sin(j * 0.05), 0.1 * np.cos(j * 0.03)], axis=1)
    poses_true = np.tile(np.eye(4), (n_poses, 1, 1))
    yaw = 0.05 * np.sin(j * 0.02)
    pitch = 0.02 * np.cos(j * 0.04)
    cy_, sy_ = np.cos(yaw), np.sin(yaw)
    cp_, sp_ = np.cos(pitch), np.sin(pitch)
    # camera looks along world +z; R = Ry(yaw) Rx(pitch)
    Ry = np.zeros((n_poses, 3, 3)); Rx = np.zeros((n_poses, 3, 3))
    Ry[:, 0, 0] = cy_; Ry[:, 0, 2] = sy_; Ry[:, 1, 1] = 1; Ry[:, 2, 0] = -sy_; Ry[:, 2, 2] = cy_
    Rx[:, 0, 0] = 1; Rx[:, 1, 1] = cp_; Rx[:, 1, 2] = -sp_; Rx[:, 2, 1] = sp_; Rx[:, 2, 2] = cp_
    poses_true[:, :3, :3] = Ry @ Rx
    poses_true[:, :3, 3] = t
    if heavy_tail:
        L = np.clip(rng.geometric(1.0 / max(mean_track - 1.0, 1.0), n_points) + 1, 2, min(100, n_poses))
    else:
        lo, hi = max(2, mean_track - 2), mean_track + 2
        L = rng.integers(lo, hi + 1, n_points)
        L = np.minimum(L, n_poses)
    start = (rng.random(n_points) * (n_poses - L + 1)).astype(np.int64)
    # sort landmarks by first observing pose (ids follow the trajectory, like a SLAM map)
    order = np.argsort(start, kind="stable")
    start, L = start[order], L[order]
    mid = start + (L - 1) / 2.0
    points_true = np.stack([step * mid + rng.uniform(-1.0, 1.0, n_points),
                            rng.uniform(-1.5, 1.5, n_points),
                            rng.uniform(4.0, 12.0, n_points)], axis=1)
    # (pose, point) incidence
    pt = np.repeat(np.arange(n_points), L)
    offs = np.arange(len(pt)) - np.repeat(np.cumsum(L) - L, L)
    ps = np.repeat(start, L) + offs
    if loop_fraction > 0:
        n_loop = int(loop_fraction * n_points)
        lp = rng.choice(n_points, n_loop, replace=False)
        lj = rng.integers(0, n_poses, n_loop)
        # keep only revisits that are not already in the run
        ok = (lj < start[lp]) | (lj >= start[lp] + L[lp])
        pt = np.concatenate([pt, lp[ok]]); ps = np.concatenate([ps, lj[ok]])
    if stereo:
        intr, cam_T = _stereo_rig(525.0, 525.0, 320.0, 240.0, 0.12)
        cam_ids = [0, 1]
    else:
        intr = np.array([[525.0, 525.0, 320.0, 240.0]]); cam_T = np.eye(4)[None]; cam_ids = [0]
    n_cam = len(cam_ids)
    cam = np.tile(np.arange(n_cam), len(pt))
    pt = np.repeat(pt, n_cam); ps = np.repeat(ps, n_cam)
    # insertion order of test_ba.cpp: per frame, the left list then the right list
    o = np.lexsort((pt, cam, ps))
    cam, pt, ps = cam[o], pt[o], ps[o]
    T_cw = inv_T(poses_true)
    Xb = np.einsum("nij,nj->ni", T_cw[ps, :3, :3], points_true[pt]) + T_cw[ps, :3, 3]
    Xc = np.einsum("nij,nj->ni", cam_T[cam, :3, :3], Xb) + cam_T[cam, :3, 3]
    uv = np.stack([intr[cam, 0] * Xc[:, 0] / Xc[:, 2] + intr[cam, 2],
                   intr[cam, 1] * Xc[:, 1] / Xc[:, 2] + intr[cam, 3]], axis=1)
    if pixel_sigma > 0:
        uv = uv + rng.normal(0, pixel_sigma, uv.shape)
    poses_init = poses_true.copy()
    lvl = pose_translation_error_level
    # pose_noise_seed: landmark shards of one problem (multi-GPU) must share the same initial poses
    prng = rng if pose_noise_seed is None else np.random.default_rng(pose_noise_seed)
    poses_init[n_fixed:, :3, 3] += prng.uniform(-lvl, lvl, (n_poses - n_fixed, 3))
    points_init = points_true + rng.uniform(-point_error_level, point_error_level, (n_points, 3))
    return FullScene(
        cam_ids=cam_ids, cam_intr=intr, cam_T=cam_T, poses_true=poses_true, poses_init=poses_init,
        fixed_poses=np.arange(n_fixed), points_true=points_true, points_init=points_init,
        fixed_points=np.zeros(0, dtype=np.int64), obs_cam=cam.astype(np.int32),
        obs_pose=ps.astype(np.int32), obs_point=pt.astype(np.int32), obs_uv=uv.astype(np.float64),
        name=name, meta=dict(seed=seed, pixel_sigma=pixel_sigma, mean_track=mean_track))


def scene_c3(seed=0, pixel_sigma=0.0, scale=1.0, pose_noise_seed=None):
    """stereo full BA, 200 poses / 50k landmarks / ~1M observations."""
    return scene_trajectory(200, int(50_000 * scale), 10, stereo=True, seed=seed, pixel_sigma=pixel_sigma,
                            name="C3_200p_50k_1M", pose_noise_seed=pose_noise_seed)


def scene_c4(seed=0, pixel_sigma=0.0, scale=1.0):
    """stereo full BA, 2000 poses / 1M landmarks / ~8M observations."""
    return scene_trajectory(int(2000 * scale), int(1_000_000 * scale), 4, stereo=True, seed=seed,
                            pixel_sigma=pixel_sigma, name="C4_2000p_1M_8M")


def scene_c5(seed=0, pixel_sigma=0.0, scale=1.0):
    """BAL-Venice-shaped mono full BA, 1778 poses / ~1M points / ~5M observations."""
    return scene_trajectory(int(1778 * scale), int(1_000_000 * scale), 5, stereo=False, seed=seed,
                            pixel_sigma=pixel_sigma, heavy_tail=True, loop_fraction=0.02,
                            name="C5_venice_shaped")


def restrict_poses(sc, lo, hi):
    """Sub-scene of a FullScene: poses [lo, hi) re-indexed from 0, the observations made from them in their original
    relative (insertion) order, and the landmarks those observations see (re-indexed in id order).  A window of a
    trajectory scene keeps the structure of the whole (track lengths, band of the reduced system) at a size the CPU
    oracle solves in seconds."""
    import copy
    keep = (sc.obs_pose >= lo) & (sc.obs_pose < hi)
    pts = np.unique(sc.obs_point[keep])
    remap = np.full(len(sc.points_init), -1, dtype=np.int64)
    remap[pts] = np.arange(len(pts))
    out = copy.copy(sc)
    out.poses_true = sc.poses_true[lo:hi]
    out.poses_init = sc.poses_init[lo:hi]
    fp = np.asarray(sc.fixed_poses, dtype=np.int64)
    out.fixed_poses = fp[(fp >= lo) & (fp < hi)] - lo
    out.points_true = sc.points_true[pts]
    out.points_init = sc.points_init[pts]
    fx = np.asarray(sc.fixed_points, dtype=np.int64)
    out.fixed_points = remap[fx[remap[fx] >= 0]] if len(fx) else fx
    out.obs_cam = sc.obs_cam[keep]
    out.obs_pose = (sc.obs_pose[keep] - lo).astype(np.int32)
    out.obs_point = remap[sc.obs_point[keep]].astype(np.int32)
    out.obs_uv = sc.obs_uv[keep]
    out.name = f"{sc.name}[poses {lo}:{hi}]"
    out.meta = dict(sc.meta, pose_window=(lo, hi))
    return out


@dataclass
class PoseOnlyBatch:
    kind: int                   # 0 mono-6dof, 1 stereo-6dof, 2 mono-planar3dof, 3 stereo-planar3dof
    offsets: np.ndarray         # (n_frames+1,) int32
    points: np.ndarray          # (n, 3) float32, reference-frame positions
    px_left: np.ndarray         # (n, 2) float32
    px_right: np.ndarray        # (n, 2) float32 (stereo) or None
    intr_left: np.ndarray       # (4,) float32
    intr_right: np.ndarray
    left_to_right: np.ndarray   # (12,) float32 R row-major | t
    poses_true: np.ndarray      # (n_frames, 12) float32
    poses_init: np.ndarray
    base_to_camera: np.ndarray = None
    world_to_last: np.ndarray = None

    @property
    def n_frames(self):
        return len(self.offsets) - 1


def _pose12(R, t):
    return np.concatenate([np.asarray(R).reshape(-1), np.asarray(t).reshape(-1)]).astype(np.float32)


def scene_poseonly_batch(n_frames=4096, n_points=300, seed=0, pixel_sigma=0.0, stereo=True,
                         right_invalid_fraction=0.0, ragged=False):
    """Config C2 (test/test_6dof_stereo_poseonly_ba.cpp:15-107), one independent problem per frame."""
    rng = np.random.default_rng(seed)
    fx = fy = 338.0
    cx, cy = 320.0, 240.0
    baseline = 0.05
    if ragged:
        counts = rng.integers(max(8, n_points // 3), n_points + 1, n_frames)
    else:
        counts = np.full(n_frames, n_points)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    n = int(offsets[-1])
    X = np.stack([rng.uniform(-1.7, 1.7, n), rng.uniform(-1.3, 1.3, n), 1.2 + rng.uniform(0, 5.0, n)], axis=1)
    frame = np.repeat(np.arange(n_frames), counts)
    poses_true = np.zeros((n_frames, 12), dtype=np.float32)
    Rs = np.zeros((n_frames, 3, 3)); ts = np.zeros((n_frames, 3))
    for f in range(n_frames):
        w = rng.normal(0, 0.02, 3)
        R = rot_y(-0.12) @ rot_z(w[2]) @ rot_y(w[1]) @ rot_x(w[0])
        t = np.array([0.4, 0.012, -0.5]) + rng.normal(0, 0.05, 3)
        Rs[f], ts[f] = R, t
        poses_true[f] = _pose12(R, t)
    # left local = T_wc^-1 X ; right local = left_to_right^-1 * left local
    Xl = np.einsum("nji,nj->ni", Rs[frame], X - ts[frame])
    Xr = Xl - np.array([baseline, 0, 0])
    pl = np.stack([fx * Xl[:, 0] / Xl[:, 2] + cx, fy * Xl[:, 1] / Xl[:, 2] + cy], axis=1)
    pr = np.stack([fx * Xr[:, 0] / Xr[:, 2] + cx, fy * Xr[:, 1] / Xr[:, 2] + cy], axis=1)
    if pixel_sigma > 0:
        pl = pl + rng.normal(0, pixel_sigma, pl.shape)
        pr = pr + rng.normal(0, pixel_sigma, pr.shape)
    if right_invalid_fraction > 0:
        bad = rng.random(n) < right_invalid_fraction
        pr[bad] = -1.0
    init = _pose12(np.eye(3), [-0.2, -0.5, 0.0])
    intr = np.array([fx, fy, cx, cy], dtype=np.float32)
    return PoseOnlyBatch(kind=1 if stereo else 0, offsets=offsets, points=X.astype(np.float32),
                         px_left=pl.astype(np.float32), px_right=pr.astype(np.float32) if stereo else None,
                         intr_left=intr, intr_right=intr.copy(),
                         left_to_right=_pose12(np.eye(3), [baseline, 0, 0]),
                         poses_true=poses_true, poses_init=np.tile(init, (n_frames, 1)))


def scene_poseonly_planar_batch(n_frames=64, n_points=300, seed=0, pixel_sigma=0.0, stereo=True):
    """Planar 3-DoF variant (test/test_3dof_stereo_poseonly_ba.cpp recipe, reduced): a ground robot
    whose base frame moves by (x, y, psi); world points are expressed in the previous base frame b1
    and world_to_last = base_to_camera (last camera frame == b1's camera)."""
    rng = np.random.default_rng(seed)
    fx = fy = 338.0
    cx, cy = 320.0, 240.0
    baseline = 0.05
    # base: x forward, y left, z up ; camera: z forward, x right, y down
    R_bc = np.array([[0.0, 0, 1], [-1, 0, 0], [0, -1, 0]])
    T_bc = make_T(R_bc, [0.1, 0.0, 0.3])  # base_to_camera (camera pose expressed in base)
    counts = np.full(n_frames, n_points)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    n = int(offsets[-1])
    # points in base frame b1, in front of the robot
    Xb = np.stack([1.5 + rng.uniform(0, 5.0, n), rng.uniform(-1.7, 1.7, n), rng.uniform(-0.2, 1.5, n)], axis=1)
    frame = np.repeat(np.arange(n_frames), counts)
    pl = np.zeros((n, 2)); pr = np.zeros((n, 2))
    poses_true = np.zeros((n_frames, 12), dtype=np.float32)
    poses_init = np.zeros((n_frames, 12), dtype=np.float32)
    w2l = np.zeros((n_frames, 12), dtype=np.float32)
    T_cb = inv_T(T_bc)
    for f in range(n_frames):
        x, y, psi = 0.3 + rng.normal(0, 0.05), rng.normal(0, 0.05), rng.normal(0, 0.05)
        T_b1b2 = make_T(rot_z(psi), [x, y, 0.0])          # current base expressed in last base
        T_b2b1 = inv_T(T_b1b2)
        sl = slice(offsets[f], offsets[f + 1])
        Xc = (Xb[sl] @ T_b2b1[:3, :3].T + T_b2b1[:3, 3]) @ T_cb[:3, :3].T + T_cb[:3, 3]
        Xr = Xc - np.array([baseline, 0, 0])
        pl[sl] = np.stack([fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy], axis=1)
        pr[sl] = np.stack([fx * Xr[:, 0] / Xr[:, 2] + cx, fy * Xr[:, 1] / Xr[:, 2] + cy], axis=1)
        # world := b1 ; pose_world_to_current = T_b1b2 * T_bc ; pose_world_to_last = T_bc
        T_wc = T_b1b2 @ T_bc
        poses_true[f] = _pose12(T_wc[:3, :3], T_wc[:3, 3])
        poses_init[f] = _pose12(T_bc[:3, :3], T_bc[:3, 3])   # initial guess: no motion
        w2l[f] = _pose12(T_bc[:3, :3], T_bc[:3, 3])
    if pixel_sigma > 0:
        pl = pl + rng.normal(0, pixel_sigma, pl.shape)
        pr = pr + rng.normal(0, pixel_sigma, pr.shape)
    intr = np.array([fx, fy, cx, cy], dtype=np.float32)
    return PoseOnlyBatch(kind=3 if stereo else 2, offsets=offsets, points=Xb.astype(np.float32),
                         px_left=pl.astype(np.float32), px_right=pr.astype(np.float32) if stereo else None,
                         intr_left=intr, intr_right=intr.copy(),
                         left_to_right=_pose12(np.eye(3), [baseline, 0, 0]),
                         poses_true=poses_true, poses_init=poses_init,
                         base_to_camera=_pose12(T_bc[:3, :3], T_bc[:3, 3]), world_to_last=w2l)
