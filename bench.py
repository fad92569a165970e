#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 bundle-adjustment engine.

Metric (BASELINE.json): LM iters/s & Schur-build obs/s.  One "step" = one full Levenberg-Marquardt
iteration (linearise, Schur complement, reduced solve, back-substitution, trial cost, accept/lambda)
over the workload's observations.  `value` = observations x LM iterations per second (whole job, all
ranks), device-resident; `lm_iters_per_s` and `schur_build_obs_per_s` are reported beside it.
Workload at N=1: config C3 "stereo full BA, 200 poses / 50k landmarks / ~1M observations on 1 B200".
N>1 (weak scaling): every rank holds all 200 poses and its own 50k landmarks (+~1M observations);
[S | rhs] and the LM scalars are all-reduced over NCCL every iteration.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                     # the reference's algorithm on host cores (oracle)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line: library chatter written to fd 1 (e.g. NCCL's version banner) is sent
# to stderr for the whole run and the result goes to the saved descriptor
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


METRIC = "LM-iteration observation throughput (observations x LM iters/s; full iteration: linearise, Schur, solve, back-substitute, cost/accept)"
UNIT = "obs/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_scene(workload, rank, scale):
    from bundle_adjustment_solver_b200 import scenes
    if workload == "c3":
        return scenes.scene_c3(seed=100 + rank, scale=scale, pose_noise_seed=7)
    if workload == "c1":
        return scenes.scene_test_ba(seed=rank)
    if workload == "c4":
        return scenes.scene_c4(seed=100 + rank, scale=scale)
    if workload == "c5":
        return scenes.scene_c5(seed=100 + rank, scale=scale)
    raise SystemExit(f"unknown workload {workload}")


WORKLOAD_NAMES = {
    "c1": "C1 test_ba.cpp stereo full BA (60 poses / 660 landmarks / 34,019 observations)",
    "c3": "C3 stereo full BA, 200 poses / 50k landmarks / ~1M observations",
    "c4": "C4 stereo full BA, 2000 poses / 1M landmarks / ~8M observations",
    "c5": "C5 BAL-Venice-shaped mono full BA (1778 poses / ~1M points / ~5M observations)",
}


def algorithmic_bytes(sz, n_obs_free_pose):
    """Algorithmic bytes / flops per LM iteration of the implemented design (DESIGN.md 'Kernels').
    pose side  (k_linearize_by_pose + k_finish_poses): 28 B/obs in pose order, A/a and the S diagonal
    point side (k_build_tiles, fused K1+K3+K4): 20 B/obs (pixel, camera) + 24 B per (pose, landmark) incidence,
               pose/point gathers, B written once (144 B/pair), per-landmark blocks (144 B), S tile flush
    back-substitution: B read once (144 B/pair) + per-landmark blocks;   cost: 28 B/obs."""
    O, Nt, Mt, N, M, P = sz["n_obs"], sz["N_total"], sz["M_total"], sz["N"], sz["M"], sz["P"]
    n = 6 * N
    b = {}
    b["linearize"] = 28 * n_obs_free_pose + 96 * Nt + 24 * Mt + 2 * 336 * N + 8 * (n + 1) * (n + 1)  # pose side + S memset
    b["schur"] = 20 * O + 24 * P + 96 * Nt + 24 * Mt + 144 * P + 144 * M + 8 * n * (n + 1) // 2 + 8 * n
    b["backsub"] = 144 * P + 8 * P + 48 * N + 24 * M + (144 + 24 + 24 + 24 + 24) * Mt
    b["update_cost"] = 28 * O + 96 * Nt * 2 + 24 * Mt
    b["solve_flops"] = n ** 3 / 3.0 + 2.0 * n * n
    # build flops: ~330 per observation (projection, weight, Jacobians, C/b/B) + Schur 216 per block pair
    b["schur_flops"] = 330.0 * O
    return b


def poseonly_c2(device, stream, frames=4096, points=300, reps=5):
    """Config C2: batched 6-DoF stereo pose-only BA, `frames` independent frames x `points` observations,
    device-resident (ba_poseonly_upload / ba_poseonly_run), timed with CUDA events on the launching stream."""
    import ctypes as C
    import torch
    from bundle_adjustment_solver_b200 import capi, scenes
    L = capi.lib()
    pb = scenes.scene_poseonly_batch(n_frames=frames, n_points=points, seed=1)
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    off = np.ascontiguousarray(pb.offsets, dtype=np.int32)
    arrs = [f32(pb.points), f32(pb.px_left), f32(pb.px_right), f32(pb.intr_left), f32(pb.intr_right),
            f32(pb.left_to_right), None, None, f32(pb.poses_init)]
    h = C.c_void_p()
    rc = L.ba_poseonly_upload(C.byref(h), device, pb.kind, pb.n_frames, capi.ptr(off), *[capi.ptr(a) for a in arrs])
    if rc != 0:
        raise SystemExit("ba_poseonly_upload failed")
    opt = capi.PoseOnlyOptions(1e-6, 1e-6, 1.5, 2.5, 100)
    for _ in range(3):
        L.ba_poseonly_run(h, C.byref(opt), C.c_void_p(stream.cuda_stream))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        L.ba_poseonly_run(h, C.byref(opt), C.c_void_p(stream.cuda_stream))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res = (capi.PoseOnlyResult * pb.n_frames)()
    poses = np.zeros((pb.n_frames, 12), dtype=np.float32)
    L.ba_poseonly_download(h, capi.ptr(poses), None, None, res)
    L.ba_poseonly_free(h)
    iters = float(np.mean([res[k].n_iterations for k in range(pb.n_frames)]))
    n_pts = int(off[-1])
    return {"workload": f"C2 batched 6-DoF stereo pose-only BA, {frames} frames x {points} points (FP32)",
            "ms_per_batch": ms, "frames_per_s": pb.n_frames / (ms * 1e-3),
            "point_iterations_per_s": n_pts * iters / (ms * 1e-3), "mean_gn_iterations": iters,
            "max_pose_err_vs_truth": float(np.abs(poses - pb.poses_true).max()),
            "bytes_per_solve": int(28 * n_pts), "note": "one warp per frame, persistent over the GN iterations; working set is L2-resident"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, loaded = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
                loaded.append(float(f[7]) > 0)
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        under = [s for s, l in zip(sm, loaded) if l] or sm
        return {"sm_mhz": float(np.median(under)) if under else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": int(sum(loaded))}


def run_reference(args, rank, world):
    """--impl reference: the reference's own algorithm on the host cores.  The reference cannot be
    compiled here (no Eigen/Ceres/OpenCV), so this is the oracle port (kind 'port'), single thread like
    the reference (it has no threads), on a bounded number of LM iterations of the same workload."""
    if rank != 0:
        return
    import oracle
    from bundle_adjustment_solver_b200 import solver as S
    sc = make_scene(args.workload, 0, args.scale)
    o = S.load_scene(oracle.FullBAOracle(), sc)
    o.sizes()
    iters = max(1, min(args.steps, args.cpu_iters))
    warm = 1 if args.warmup > 0 else 0
    opt = oracle.default_full_options(max_num_iterations=max(1, warm), threshold_cost_change=0.0,
                                      threshold_step_size=0.0)
    if warm:
        o.solve(opt)
    opt.max_num_iterations = iters
    t0 = time.perf_counter()
    infos, _ = o.solve(opt)
    dt = time.perf_counter() - t0
    value = sc.n_obs * len(infos) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(infos),
        "warmup": warm, "ms_per_step": 1e3 * dt / len(infos), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.workload], "n_obs": sc.n_obs, "scale": args.scale},
        "lm_iters_per_s": len(infos) / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{len(infos)} LM iterations of the same workload, oracle/ba_oracle.cpp -O2, "
                                   f"host has {os.cpu_count()} cores; the reference is single-threaded"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOAD_NAMES))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cpu-iters", type=int, default=4, help="LM iterations of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-poseonly", action="store_true", help="skip the C2 pose-only sub-benchmark")
    ap.add_argument("--e2e-max-iters", type=int, default=300,
                    help="iteration cap of the end-to-end solve (SURVEY 8d: thresholds 1e-6f, max 300 iterations)")
    ap.add_argument("--quick", action="store_true", help="profiling run: no clock-settling loop, phases, e2e or CPU baseline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.steps = max(1, args.steps)
    args.warmup = max(3, args.warmup)

    import ctypes as C

    import torch
    import torch.distributed as dist

    from bundle_adjustment_solver_b200 import capi
    from bundle_adjustment_solver_b200 import solver as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)

    # N > 1: C3 (the default) is weak scaling -- every rank brings its own 50k landmarks seen from the SAME 200 poses
    # (pose_noise_seed fixed).  C4 / C5 are quoted in BASELINE.json as ONE problem sharded over the GPUs: every rank
    # generates the same scene and keeps its contiguous landmark range (strong scaling).
    strong = world > 1 and args.workload in ("c4", "c5")
    if strong:
        from bundle_adjustment_solver_b200 import sharding
        sc = sharding.shard_scene(make_scene(args.workload, 0, args.scale), rank, world)
    else:
        sc = make_scene(args.workload, rank, args.scale)
    L = capi.lib()

    def new_solver():
        s = S.FullBundleAdjustmentSolver(device=local_rank, stream=stream.cuda_stream)
        S.load_scene(s, sc)
        return s

    def join_comm(s, sz):
        if world == 1:
            return
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
        rc = L.ba_comm_init(s.h, raw, rank, world, int(tot[0]), int(tot[1]))
        if rc != 0:
            raise SystemExit(f"ba_comm_init failed: {L.ba_last_error(s.h)}")

    with torch.cuda.stream(stream):
        s = new_solver()
        s._upload()
        sz = s.sizes()
        join_comm(s, sz)
        T12_0, X_0 = s.internal_parameters()
        n_obs_free_pose = int(np.count_nonzero(~np.isin(sc.obs_pose, sc.fixed_poses)))

        def solve_n(n_it, check_every=None, profile=False):
            s.set_profile(profile)
            opt = capi.default_options(max_num_iterations=n_it, threshold_cost_change=0.0, threshold_step_size=0.0,
                                       check_every=check_every or n_it)
            res = capi.Result()
            rc = L.ba_solve(s.h, C.byref(opt), None, 0, C.byref(res))
            if rc != 0:
                raise SystemExit(f"ba_solve failed: {L.ba_last_error(s.h)}")
            assert res.n_iterations == n_it, (res.n_iterations, n_it)
            return res

        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        # ---- warm-up (untimed): W steps, then keep the GPU busy ~1 s so the clock samples see load
        solve_n(args.warmup)
        t_w = time.perf_counter()
        while not args.quick and time.perf_counter() - t_w < 1.0:
            solve_n(args.steps)
        # ---- timed: exactly K steps from the initial guess
        s.update_parameters_internal(T12_0, X_0)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        res = solve_n(args.steps)
        ev1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = ev0.elapsed_time(ev1)
        launches = int(res.kernel_launches)
        clocks = sampler.stop() if rank == 0 else None
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        nobs_all = torch.tensor([sz["n_obs"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(nobs_all)
        ms = float(tmax[0])
        total_obs = float(nobs_all[0])
        value = total_obs * args.steps / (ms * 1e-3)

        if args.quick:
            if rank == 0:
                emit({"quick": True, "ms_per_step": ms / args.steps, "value": value, "gpu_launches": launches})
            if world > 1:
                dist.destroy_process_group()
            return
        # ---- per-phase device times (CUDA events on the launching stream, plain launches)
        s.update_parameters_internal(T12_0, X_0)
        solve_n(args.warmup, profile=True)
        s.update_parameters_internal(T12_0, X_0)
        resp = solve_n(args.steps, profile=True)
        s.set_profile(False)
        ph = {"linearize": resp.t_linearize_ms / args.steps, "schur": resp.t_schur_ms / args.steps,
              "solve": resp.t_solve_ms / args.steps, "backsub": resp.t_backsub_ms / args.steps,
              "update_cost": resp.t_update_cost_ms / args.steps}

        # ---- end to end through the C-ABI from host buffers: register + finalize (H2D pack) +
        #      solve to convergence + read back.  Per-step bytes = totals / LM iterations.
        e2e = None
        if not args.no_e2e:
            # the timed solver above is done: release it, so that the end-to-end solve runs in the steady state of a
            # process that solves one problem after another (device buffers come back from the stream-ordered pool
            # instead of growing it next to a live 300 MB problem -- that growth alone varied between 10 and 150 ms)
            del s
            torch.cuda.synchronize()
            s2 = S.FullBundleAdjustmentSolver(device=local_rank, stream=stream.cuda_stream)
            join_comm_needed = world > 1
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            S.load_scene(s2, sc)
            t_reg = time.perf_counter()
            s2._upload()
            if join_comm_needed:
                join_comm(s2, sz)   # N > 1: communicator creation (NCCL set-up, ~0.5 s) is inside the timed region
            t_fin = time.perf_counter()
            opt = capi.default_options(max_num_iterations=args.e2e_max_iters, threshold_cost_change=1e-6,
                                       threshold_step_size=1e-6)
            summ = S.Summary()
            s2.solve(opt, summ)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            stages = {"register_ms": 1e3 * (t_reg - t0), "finalize_h2d_ms": 1e3 * (t_fin - t_reg),
                      "solve_writeback_ms": 1e3 * (t0 + dt - t_fin)}
            it = len(summ.optimization_info_list)
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0])
            h2d = sc.n_obs * (16 + 16 + 12 + 12) + sz["N_total"] * 96 * 2 + sz["M_total"] * 24 * 2 + sz["P"] * 12
            d2h = sz["N_total"] * 96 + sz["M_total"] * 24 + it * 64
            e2e = {"value": total_obs * it / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d / max(it, 1)),
                   "d2h_bytes_per_step": int(d2h / max(it, 1)), "lm_iterations": it, "converged": bool(summ.convergence_status),
                   "wall_s": dt, "stages_ms": stages, "final_cost": summ.optimization_info_list[-1].cost if it else None,
                   "includes": "host registration, FinalizeParameters (sorts, H2D pack), LM loop to convergence, D2H write-back"}
            del s2

    po2 = None
    if world == 1 and not args.no_poseonly:
        with torch.cuda.stream(stream):
            po2 = poseonly_c2(local_rank, stream)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = peaks()
    ab = algorithmic_bytes(sz, n_obs_free_pose)
    phases = {}
    for k in ("linearize", "schur", "backsub", "update_cost"):
        gbs = ab[k] / (ph[k] * 1e-3) / 1e9 if ph[k] > 0 else 0.0
        phases[k] = {"ms": ph[k], "algorithmic_bytes": ab[k], "achieved_gbs": gbs, "frac_hbm": gbs / hbm_peak}
    phases["solve"] = {"ms": ph["solve"], "algorithmic_flops": ab["solve_flops"],
                       "achieved_tflops": ab["solve_flops"] / (ph["solve"] * 1e-3) / 1e12 if ph["solve"] > 0 else 0.0}
    t_build = ph["linearize"] + ph["schur"]
    kernels = {"linearize": "k_linearize_by_pose+k_finish_poses", "schur": "k_build_tiles (linearise + C^-1 + Schur DMMA GEMM)",
               "backsub": "k_backsub_pairs+k_backsub_points", "update_cost": "k_cost+k_update_poses",
               "solve": "k_chol_banded_smem (DMMA window update; cluster / multi-kernel DMMA variants for other structures)"}
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the default C3 workload from the
    # committed `ncu --set full` capture profiles/r01_ncu_full_c3_summary.md; None for any other workload
    ncu_traffic = {"solve": 0.99e6, "schur": 34.66e6 + 27.96e6, "linearize": 25.38e6 + 0.25e6,
                   "backsub": 76.02e6 + 2.67e6 + 9.86e6, "update_cost": 29.24e6 + 0.11e6}
    traffic = (lambda k: ncu_traffic[k]) if (args.workload == "c3" and args.scale == 1.0 and world == 1) else (lambda k: None)
    # the dominant kernel of the step by device time
    dom = max(ph, key=lambda k: ph[k])
    fp64_tensor_peak = 37.1   # TFLOP/s, measured on this pool's B200 with profiles/micro/dmma_peak.cu (m8n8k4 DMMA)
    if dom == "solve":
        roofline = {"bound": "tensor", "kernel": kernels[dom], "achieved": phases["solve"]["achieved_tflops"],
                    "peak": fp64_tensor_peak, "unit": "TFLOP/s", "frac": phases["solve"]["achieved_tflops"] / fp64_tensor_peak,
                    "traffic": traffic("solve"),
                    "peak_source": "measured FP64 DMMA throughput (profiles/micro/dmma_peak.cu; MEASURED_PEAKS.json has no FP64 figure)",
                    "note": "achieved = dense-equivalent n^3/3 + 2n^2 flops / solve time; the banded factorisation is bound by "
                            "its chain of n/8 dependent panel steps (latency), not by FP64 throughput - see DESIGN.md 4 (K5)"}
    else:
        roofline = {"bound": "hbm", "kernel": kernels[dom], "achieved": phases[dom]["achieved_gbs"], "peak": hbm_peak,
                    "unit": "GB/s", "frac": phases[dom]["frac_hbm"], "traffic": traffic(dom), "peak_source": peak_src}
    # the Jacobian / Schur build (BASELINE metric "Schur-build obs/s"): both build phases together
    build_bytes = ab["linearize"] + ab["schur"]
    build_gbs = build_bytes / (t_build * 1e-3) / 1e9 if t_build > 0 else 0.0
    roofline_build = {"bound": "hbm", "kernel": kernels["linearize"] + " ; " + kernels["schur"], "achieved": build_gbs,
                      "peak": hbm_peak, "unit": "GB/s", "frac": build_gbs / hbm_peak, "algorithmic_bytes": build_bytes,
                      "traffic": (traffic("linearize") + traffic("schur")) if traffic("schur") is not None else None,
                      "peak_source": peak_src,
                      "fp64_tflops": (ab["schur_flops"] / (t_build * 1e-3) / 1e12) if t_build > 0 else 0.0}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        o = S.load_scene(oracle.FullBAOracle(), sc)
        o.sizes()
        it_cpu = max(1, args.cpu_iters)
        oo = oracle.default_full_options(max_num_iterations=it_cpu, threshold_cost_change=0.0, threshold_step_size=0.0)
        t0 = time.perf_counter()
        infos, _ = o.solve(oo)
        dtc = time.perf_counter() - t0
        cpu = {"value": sc.n_obs * len(infos) / dtc, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{len(infos)} LM iterations of the same workload on 1 of {os.cpu_count()} host cores "
                         f"(oracle/ba_oracle.cpp; the reference is single-threaded), {dtc:.1f} s",
               "ms_per_step": 1e3 * dtc / len(infos)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.workload], "per_gpu": {k: sz[k] for k in ("N", "M", "P", "n_obs")},
                   "scale": args.scale, "parallelism": f"landmark-sharded x{world}, S all-reduce" if world > 1 else "single GPU",
                   "l2": "no flush: per-iteration working set %.0f MB exceeds the 126 MB L2" % (
                       (ab["linearize"] + ab["schur"] + ab["backsub"] + ab["update_cost"]) / 1e6),
                   "cuda_graph": world == 1},
        "lm_iters_per_s": args.steps / (ms * 1e-3),
        "schur_build_obs_per_s": total_obs / (t_build * 1e-3) if t_build > 0 else None,
        "phases": phases, "roofline": roofline, "roofline_build": roofline_build, "poseonly_c2": po2, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
        "clocks": clocks,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
