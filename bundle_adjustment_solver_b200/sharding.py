"""Landmark sharding for the multi-GPU full-BA path.

Landmarks (with all their observations) are split into `world` contiguous id ranges balanced by
observation count; poses, cameras and the fixed sets are replicated.  Each rank builds its partial
[S | rhs] (its A/a contributions included) and the engine all-reduces it over NCCL; the reduced solve is
replicated and every rank back-substitutes its own landmarks.  No observation crosses ranks.
"""
import copy

import numpy as np


def landmark_ranges(obs_point, n_points, world):
    """Contiguous landmark id ranges [lo, hi) per rank, balanced on observation counts."""
    counts = np.bincount(np.asarray(obs_point), minlength=n_points).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        bounds.append(int(np.searchsorted(cum, target, side="left")))
    bounds.append(n_points)
    bounds = np.maximum.accumulate(np.asarray(bounds))
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]


def shard_scene(sc, rank, world):
    """The rank's view of a scenes.FullScene: all poses, its landmarks re-indexed from 0, its observations
    in their original relative (insertion) order."""
    lo, hi = landmark_ranges(sc.obs_point, len(sc.points_init), world)[rank]
    keep = (sc.obs_point >= lo) & (sc.obs_point < hi)
    out = copy.copy(sc)
    out.points_true = sc.points_true[lo:hi]
    out.points_init = sc.points_init[lo:hi]
    fp = np.asarray(sc.fixed_points, dtype=np.int64)
    out.fixed_points = fp[(fp >= lo) & (fp < hi)] - lo
    out.obs_cam = sc.obs_cam[keep]
    out.obs_pose = sc.obs_pose[keep]
    out.obs_point = (sc.obs_point[keep] - lo).astype(np.int32)
    out.obs_uv = sc.obs_uv[keep]
    out.name = f"{sc.name}[shard {rank}/{world}]"
    out.meta = dict(sc.meta, landmark_range=(lo, hi))
    return out


def frame_ranges(n_frames, world):
    """Batches of independent pose-only problems (config C2) split evenly over the GPUs, no communication
    (SURVEY.md 8e row 2): frame range [lo, hi) per rank, sizes differ by at most one."""
    base, extra = divmod(int(n_frames), int(world))
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < extra else 0))
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_poseonly_batch(pb, rank, world):
    """The rank's frames of a scenes.PoseOnlyBatch (offsets rebased, per-point arrays sliced; the rig is shared)."""
    lo, hi = frame_ranges(pb.n_frames, world)[rank]
    p0, p1 = int(pb.offsets[lo]), int(pb.offsets[hi])
    out = copy.copy(pb)
    out.offsets = (np.asarray(pb.offsets[lo:hi + 1]) - p0).astype(np.int32)
    out.points = pb.points[p0:p1]
    out.px_left = pb.px_left[p0:p1]
    out.px_right = None if pb.px_right is None else pb.px_right[p0:p1]
    out.poses_true = pb.poses_true[lo:hi]
    out.poses_init = pb.poses_init[lo:hi]
    for name in ("base_to_camera", "world_to_last"):      # per-frame (n_frames, 12) or shared (12,)
        a = getattr(pb, name)
        if a is not None and np.ndim(a) == 2 and len(a) == pb.n_frames:
            setattr(out, name, a[lo:hi])
    return out
