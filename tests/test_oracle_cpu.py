"""CPU tests (no GPU): the oracle against the reference's only known-answer vectors, against
independent derivations (finite differences, SciPy), against its committed golden fixtures, and the
reference's behavioural quirks the device path must reproduce."""
import os

import numpy as np
import pytest
import scipy.linalg

import oracle
from bundle_adjustment_solver_b200 import scenes
from bundle_adjustment_solver_b200 import solver as S
from helpers import load_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_projection_known_answer_vectors():
    """test/test_projection_of_3d_point.cc:11-32: the 9 (left px, right px, 3-D point) triples, fx=fy=500,
    cx=200, cy=100, baseline 0.1.  With the rig convention of test_ba.cpp:79-98 the oracle's cost is 0."""
    left = np.array([[100, 50], [200, 50], [300, 50], [100, 100], [200, 100], [300, 100], [100, 150], [200, 150], [300, 150]], float)
    right = np.array([[90, 50], [190, 50], [290, 50], [75, 100], [175, 100], [275, 100], [50, 150], [150, 150], [250, 150]], float)
    X = np.array([[-1.0, -0.5, 5.0], [0.0, -0.5, 5.0], [1.0, -0.5, 5.0], [-0.4, 0.0, 2.0], [0.0, 0.0, 2.0], [0.4, 0.0, 2.0],
                  [-0.2, 0.1, 1.0], [0.0, 0.1, 1.0], [0.2, 0.1, 1.0]])
    o = oracle.FullBAOracle()
    o.add_camera(0, 500, 500, 200, 100, np.eye(4))
    o.add_camera(1, 500, 500, 200, 100, scenes.inv_T(scenes.make_T(np.eye(3), [0.1, 0, 0])))
    o.add_poses(np.eye(4)[None])
    o.add_points(X)
    n = len(X)
    o.add_observations(np.r_[np.zeros(n), np.ones(n)], np.zeros(2 * n), np.r_[np.arange(n), np.arange(n)],
                       np.vstack([left, right]))
    o.sizes()
    assert o.cost() < 1e-12
    # and a wrong convention is detected: flipping the baseline sign gives a large cost
    o2 = oracle.FullBAOracle()
    o2.add_camera(0, 500, 500, 200, 100, np.eye(4))
    o2.add_camera(1, 500, 500, 200, 100, scenes.make_T(np.eye(3), [0.1, 0, 0]))
    o2.add_poses(np.eye(4)[None]); o2.add_points(X)
    o2.add_observations(np.r_[np.zeros(n), np.ones(n)], np.zeros(2 * n), np.r_[np.arange(n), np.arange(n)],
                        np.vstack([left, right]))
    o2.sizes()
    assert o2.cost() > 1.0


def test_scene_c1_matches_survey_counts():
    sc = scenes.scene_test_ba(seed=0)
    assert len(sc.points_true) == 660 and len(sc.poses_true) == 60
    assert sc.n_obs == 34019                      # derived from the deterministic visibility test
    o = load_oracle(sc)
    assert o.sizes() == dict(N=55, M=660, P=16557, n_obs=34019, N_total=60, M_total=660)


@pytest.mark.parametrize("n", [1, 3, 6, 17])
def test_ldlt_restatement(n):
    rng = np.random.default_rng(n)
    L = oracle.lib()
    M = rng.normal(size=(n, n))
    A = M @ M.T + 1e-3 * np.eye(n)
    b = rng.normal(size=(n, 2))
    x = np.asfortranarray(b.copy())
    L.orc_ldlt_solve_f64(n, oracle._p(np.asfortranarray(A)), oracle._p(x), 2)
    np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-8)
    # indefinite but non-singular: LDLT (unlike LLT) still solves it
    A2 = A - 2.0 * np.trace(A) / n * np.eye(n)
    x = np.asfortranarray(b.copy())
    L.orc_ldlt_solve_f64(n, oracle._p(np.asfortranarray(A2)), oracle._p(x), 2)
    np.testing.assert_allclose(x, np.linalg.solve(A2, b), rtol=1e-6)
    # zero matrix -> zero solution (unobserved landmark: C = 0 => Cinv = 0, SURVEY trap 6)
    x = np.asfortranarray(b.copy())
    L.orc_ldlt_solve_f64(n, oracle._p(np.zeros((n, n))), oracle._p(x), 2)
    assert np.all(x == 0)
    # float version
    xf = np.asfortranarray(b.astype(np.float32))
    L.orc_ldlt_solve_f32(n, oracle._p(np.asfortranarray(A.astype(np.float32))), oracle._p(xf), 2)
    np.testing.assert_allclose(xf, np.linalg.solve(A, b), rtol=5e-2, atol=1e-3)


def test_se3_exp_matches_matrix_exponential():
    rng = np.random.default_rng(0)
    L = oracle.lib()
    for scale in (1.0, 1e-3, 1e-9, 0.0):
        xi = rng.normal(size=6) * scale
        out = np.zeros(12)
        L.orc_se3_exp_f64(oracle._p(xi), oracle._p(out))
        v, w = xi[:3], xi[3:]
        G = np.zeros((4, 4))
        G[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        G[:3, 3] = v
        E = scipy.linalg.expm(G)
        np.testing.assert_allclose(out[:9].reshape(3, 3), E[:3, :3], atol=1e-12)
        np.testing.assert_allclose(out[9:], E[:3, 3], atol=1e-12)


def _small_scene(seed=0, n_fixed=2):
    return scenes.scene_trajectory(8, 30, 4, stereo=True, seed=seed, n_fixed=n_fixed, name="small")


def _residuals(T12, X, sc, scaler=0.01):
    """Independent numpy restatement of r (scaled units) for finite differences."""
    R = T12[:, :9].reshape(-1, 3, 3)
    t = T12[:, 9:]
    Xb = np.einsum("nij,nj->ni", R[sc.obs_pose], X[sc.obs_point]) + t[sc.obs_pose]
    cT = sc.cam_T.copy()
    cT[:, :3, 3] *= scaler
    Xc = np.einsum("nij,nj->ni", cT[sc.obs_cam, :3, :3], Xb) + cT[sc.obs_cam, :3, 3]
    intr = sc.cam_intr * scaler
    u = intr[sc.obs_cam, 0] * Xc[:, 0] / Xc[:, 2] + intr[sc.obs_cam, 2] - sc.obs_uv[:, 0] * scaler
    v = intr[sc.obs_cam, 1] * Xc[:, 1] / Xc[:, 2] + intr[sc.obs_cam, 3] - sc.obs_uv[:, 1] * scaler
    return np.stack([u, v], axis=1)


def _se3_exp_np(xi):
    v, w = xi[:3], xi[3:]
    G = np.zeros((4, 4))
    G[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    G[:3, 3] = v
    return scipy.linalg.expm(G)


def test_blocks_against_finite_differences():
    """A, a, C, b and (in corrected mode) B equal J^T W J / -J^T W r with J from finite differences of an
    independent numpy projection -- pins the analytic Jacobians of full...cpp:770-829 incl. the left
    se3 update and the -[Xb]x convention."""
    sc = _small_scene(seed=3)
    o = load_oracle(sc)
    o.build_only(thres_huber=1.0, lam=0.0, b_accumulate=1, do_solve=False)
    sz = o.sizes()
    T12, X = o.get_internal()
    oj, oi = o.opt_ids()
    r0 = _residuals(T12, X, sc)
    w = np.where(np.abs(r0).sum(1) > 1.0, 1.0 / np.abs(r0).sum(1), 1.0)
    n_obs = sc.n_obs
    eps = 1e-7
    J = np.zeros((2 * n_obs, 6 * sz["N"] + 3 * sz["M"]))
    for jo, j in enumerate(oj):
        for k in range(6):
            xi = np.zeros(6); xi[k] = eps
            D = _se3_exp_np(xi)
            T2 = T12.copy()
            Tj = np.eye(4); Tj[:3, :3] = T12[j, :9].reshape(3, 3); Tj[:3, 3] = T12[j, 9:]
            Tn = D @ Tj
            T2[j, :9] = Tn[:3, :3].reshape(-1); T2[j, 9:] = Tn[:3, 3]
            J[:, 6 * jo + k] = ((_residuals(T2, X, sc) - r0) / eps).reshape(-1)
    for io, i in enumerate(oi):
        for k in range(3):
            X2 = X.copy(); X2[i, k] += eps
            J[:, 6 * sz["N"] + 3 * io + k] = ((_residuals(T12, X2, sc) - r0) / eps).reshape(-1)
    W = np.repeat(w, 2)
    H = J.T @ (W[:, None] * J)
    g = -J.T @ (W * r0.reshape(-1))
    nP = 6 * sz["N"]
    A = o.dump("A").reshape(-1, 6, 6); a = o.dump("a").reshape(-1, 6)
    C = o.dump("C").reshape(-1, 3, 3); b = o.dump("b").reshape(-1, 3)
    for jo in range(sz["N"]):
        np.testing.assert_allclose(A[jo], H[6 * jo:6 * jo + 6, 6 * jo:6 * jo + 6], rtol=2e-5, atol=1e-6 * np.abs(H).max())
        np.testing.assert_allclose(a[jo], g[6 * jo:6 * jo + 6], rtol=2e-5, atol=1e-6 * np.abs(g).max())
    for io in range(sz["M"]):
        s = nP + 3 * io
        np.testing.assert_allclose(C[io], H[s:s + 3, s:s + 3], rtol=2e-5, atol=1e-6 * np.abs(H).max())
        np.testing.assert_allclose(b[io], g[s:s + 3], rtol=2e-5, atol=1e-6 * np.abs(g).max())
    B = o.dump("B").reshape(-1, 6, 3)
    pj, pi = o.pairs()
    jmap = {j: k for k, j in enumerate(oj)}; imap = {i: k for k, i in enumerate(oi)}
    for p in range(len(pj)):
        jo, io = jmap[pj[p]], imap[pi[p]]
        np.testing.assert_allclose(B[p], H[6 * jo:6 * jo + 6, nP + 3 * io:nP + 3 * io + 3], rtol=2e-5,
                                   atol=1e-6 * np.abs(H).max())


def test_last_writer_wins_quirk():
    """full...cpp:826 assigns B_ji: with a stereo rig only the LAST inserted observation of a (pose,point)
    pair survives.  Reference-exact B == B built from the right-camera observations alone; corrected B ==
    left + right."""
    sc = _small_scene(seed=1)
    o_ref = load_oracle(sc); o_ref.build_only(1.0, 0.0, 0, False)
    o_acc = load_oracle(sc); o_acc.build_only(1.0, 0.0, 1, False)
    import copy
    def only_cam(c):
        s2 = copy.copy(sc)
        k = sc.obs_cam == c
        s2.obs_cam, s2.obs_pose, s2.obs_point, s2.obs_uv = sc.obs_cam[k], sc.obs_pose[k], sc.obs_point[k], sc.obs_uv[k]
        o = load_oracle(s2); o.build_only(1.0, 0.0, 1, False)
        return o
    o_l, o_r = only_cam(0), only_cam(1)
    assert o_ref.pairs()[0].tolist() == o_r.pairs()[0].tolist()
    np.testing.assert_allclose(o_ref.dump("B"), o_r.dump("B"), rtol=1e-13)
    np.testing.assert_allclose(o_acc.dump("B"), o_l.dump("B") + o_r.dump("B"), rtol=1e-12, atol=1e-14)
    # A, C, a, b always accumulate both cameras
    np.testing.assert_allclose(o_ref.dump("C"), o_l.dump("C") + o_r.dump("C"), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(o_ref.dump("A"), o_l.dump("A") + o_r.dump("A"), rtol=1e-12, atol=1e-14)


def test_schur_system_is_consistent():
    """S x = rhs and y = Cinv (b - B^T x) reproduce the solution of the full damped normal equations."""
    sc = _small_scene(seed=2)
    o = load_oracle(sc)
    lam = 0.5
    o.build_only(1.0, lam, 1, True)
    sz = o.sizes()
    N, M = sz["N"], sz["M"]
    A = o.dump("A").reshape(N, 6, 6); C = o.dump("C").reshape(M, 3, 3); B = o.dump("B").reshape(-1, 6, 3)
    a = o.dump("a"); b = o.dump("b")
    pj, pi = o.pairs(); oj, oi = o.opt_ids()
    jmap = {j: k for k, j in enumerate(oj)}; imap = {i: k for k, i in enumerate(oi)}
    H = np.zeros((6 * N + 3 * M, 6 * N + 3 * M))
    for j in range(N): H[6 * j:6 * j + 6, 6 * j:6 * j + 6] = A[j]
    for i in range(M): H[6 * N + 3 * i:6 * N + 3 * i + 3, 6 * N + 3 * i:6 * N + 3 * i + 3] = C[i]
    for p in range(len(pj)):
        j, i = jmap[pj[p]], imap[pi[p]]
        H[6 * j:6 * j + 6, 6 * N + 3 * i:6 * N + 3 * i + 3] = B[p]
        H[6 * N + 3 * i:6 * N + 3 * i + 3, 6 * j:6 * j + 6] = B[p].T
    sol = np.linalg.solve(H, np.r_[a, b])
    np.testing.assert_allclose(o.dump("x"), sol[:6 * N], rtol=1e-6, atol=1e-9 * np.abs(sol).max())
    np.testing.assert_allclose(o.dump("y"), sol[6 * N:], rtol=1e-6, atol=1e-9 * np.abs(sol).max())
    Sm = o.dump("S").reshape(6 * N, 6 * N)
    np.testing.assert_allclose(Sm, Sm.T, atol=0)  # exactly symmetric (mirrored, :874-876)


def test_lm_loop_quirks_and_convergence():
    """Un-squared cost, rho on damped blocks, previous_cost overwritten on every iteration, lambda clamps."""
    sc = scenes.scene_test_ba(seed=0)
    o = load_oracle(sc)
    infos, conv = o.solve(oracle.default_full_options(max_num_iterations=300, threshold_cost_change=1e-6,
                                                      threshold_step_size=1e-6))
    assert conv and len(infos) == 213                 # 213 / 95 / 96 for seeds 0 / 1 / 2
    lam = np.array([i.damping_term for i in infos])
    assert lam[0] == pytest.approx(100.0 * np.float32(0.33)) and lam.min() == 1e-10
    assert all(i.iteration_status == 1 for i in infos)  # every step UPDATE_TRUST_MORE, none rejected
    # forced NO_CONVERGENCE when the last allowed iteration is reached (:977-979)
    o = load_oracle(sc)
    infos, conv = o.solve(oracle.default_full_options(max_num_iterations=5, threshold_cost_change=1e9,
                                                      threshold_step_size=1e9))
    assert len(infos) == 1 and conv        # converges at once with huge thresholds
    o = load_oracle(sc)
    infos, conv = o.solve(oracle.default_full_options(max_num_iterations=1, threshold_cost_change=1e9,
                                                      threshold_step_size=1e9))
    assert len(infos) == 1 and not conv    # ...unless it is the last allowed iteration


REJECT_KW = dict(max_num_iterations=25, threshold_cost_change=1e-9, threshold_step_size=1e-9, initial_lambda=1e-10)


def test_rejected_step_reporting():
    """The reject branch of the LM loop (full_bundle_adjustment_solver.cpp:939-953, 995-1005, 457-482), forced by
    starting the test_ba.cpp scene undamped (initial_lambda 1e-10): Gauss-Newton steps from the noisy start
    overshoot, rho <= 0.25, the step is SKIPPED.  Every assertion is unconditional."""
    sc = scenes.scene_test_ba(seed=0)
    o = load_oracle(sc)
    infos, conv = o.solve(oracle.default_full_options(**REJECT_KW))
    st = [i.iteration_status for i in infos]
    n_obs = o.sizes()["n_obs"]
    assert len(infos) == 25 and not conv
    assert st.count(2) >= 8 and st.count(0) >= 2 and st.count(1) >= 8     # SKIPPED, UPDATE, UPDATE_TRUST_MORE all occur
    assert st[:5] == [1, 1, 1, 1, 2] and 2 in st[12:]
    lam = [1e-10] + [i.damping_term for i in infos]
    inc, dec = float(np.float32(3.0)), float(np.float32(0.33))
    for k, i in enumerate(infos):
        if st[k] == 2:
            # SKIPPED row: cost replaced by the previous cost, change 0, avg = sqrt(prev / n_obs) (:995-1000); lambda x3
            assert i.cost_change == 0.0
            assert i.average_reprojection_error == pytest.approx(np.sqrt(i.cost / n_obs), rel=1e-14)
            assert i.damping_term == pytest.approx(min(100.0, lam[k] * inc), rel=1e-15)
        elif st[k] == 1:
            assert i.damping_term == pytest.approx(max(1e-10, lam[k] * dec), rel=1e-15)
            assert i.average_reprojection_error == pytest.approx(i.cost / n_obs, rel=1e-14)
        else:
            assert i.damping_term == lam[k]                                # 0.25 < rho <= 0.5: lambda kept
    # `previous_cost = current_cost` is unconditional (:1005): the row after a SKIPPED row reports the REJECTED trial
    # cost as its previous cost, not the cost of the parameters that were kept
    k = st.index(2)
    assert st[k + 1] == 2
    assert infos[k].cost == infos[k - 1].cost                  # first SKIPPED row: the accepted cost
    assert infos[k + 1].cost > infos[k].cost * (1 + 1e-4)      # next row: the rejected trial's (higher) cost
    # the rejected trial cost of iteration k is what a run stopped there saw (debug dump: last trial cost)
    o2 = load_oracle(sc)
    kw = dict(REJECT_KW); kw["max_num_iterations"] = k + 1
    infos2, _ = o2.solve(oracle.default_full_options(**kw))
    assert [i.iteration_status for i in infos2] == st[:k + 1]
    assert o2.dump("scalars")[1] == infos[k + 1].cost
    # Revert (:468-482): after the SKIPPED iteration the parameters are the reserved ones -> their cost is the accepted one
    assert o2.cost() == pytest.approx(infos[k - 1].cost, rel=1e-13)
    # a later accepted step can even raise the cost of the kept parameters: rho is measured against the rejected cost
    kept = [(j, infos[j].cost) for j in range(len(st)) if st[j] != 2]
    assert any(c1 > c0 for (_, c0), (_, c1) in zip(kept, kept[1:]))


def test_corrected_mode_reaches_scipy_minimum():
    """Cross-check replacing the Ceres comparison (test_compare_ceres_vs_native.cpp): on a small noise-free
    problem the corrected solver converges to the least-squares minimiser that scipy finds."""
    from scipy.optimize import least_squares
    sc = _small_scene(seed=7)
    o = load_oracle(sc)
    infos, conv = o.solve(oracle.default_full_options(max_num_iterations=100, threshold_cost_change=1e-12,
                                                      threshold_step_size=1e-12, b_accumulate=1,
                                                      threshold_huber_loss=1e9))
    T12, X = o.get_internal()
    assert infos[-1].cost < 1e-6 * o.initial_cost()
    # scipy on the same internal parameterisation (points only + free poses via se3 left update)
    oj, oi = o.opt_ids()
    o0 = load_oracle(sc); o0.sizes()
    T0, X0 = o0.get_internal()

    def fun(p):
        T = T0.copy(); Xp = X0.copy()
        for k, j in enumerate(oj):
            D = _se3_exp_np(p[6 * k:6 * k + 6])
            Tj = np.eye(4); Tj[:3, :3] = T0[j, :9].reshape(3, 3); Tj[:3, 3] = T0[j, 9:]
            Tn = D @ Tj
            T[j, :9] = Tn[:3, :3].reshape(-1); T[j, 9:] = Tn[:3, 3]
        for k, i in enumerate(oi):
            Xp[i] += p[6 * len(oj) + 3 * k:6 * len(oj) + 3 * k + 3]
        return _residuals(T, Xp, sc).reshape(-1)

    res = least_squares(fun, np.zeros(6 * len(oj) + 3 * len(oi)), xtol=1e-14, ftol=1e-14, gtol=1e-14)
    assert np.abs(res.fun).max() < 1e-7
    assert np.abs(_residuals(T12, X, sc)).max() < 1e-6


@pytest.mark.parametrize("name", ["full_ba_c1_seed0_accum0", "full_ba_c1_seed1_accum0", "full_ba_c1_seed0_accum1"])
def test_oracle_reproduces_golden_full_ba(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    seed = int(name.split("seed")[1][0]); accum = int(name[-1])
    sc = scenes.scene_test_ba(seed=seed)
    o = load_oracle(sc)
    infos, conv = o.solve(oracle.default_full_options(max_num_iterations=300, threshold_cost_change=1e-6,
                                                      threshold_step_size=1e-6, b_accumulate=accum))
    assert conv == bool(g["converged"]) and len(infos) == int(g["n_iterations"])
    np.testing.assert_allclose([i.cost for i in infos], g["cost"], rtol=1e-9)
    np.testing.assert_allclose(o.get_points(), g["points"], atol=1e-9)
    np.testing.assert_allclose(o.get_poses(), g["poses"], atol=1e-9)


@pytest.mark.parametrize("name", ["stereo6", "mono6", "stereo3", "mono3"])
def test_oracle_reproduces_golden_poseonly(name):
    from golden.make_golden import poseonly
    g = np.load(os.path.join(GOLD, f"poseonly_{name}.npz"))
    cur = poseonly(name)
    np.testing.assert_allclose(cur["poses"], g["poses"], atol=1e-6)
    assert cur["n_iterations"].tolist() == g["n_iterations"].tolist()
    assert np.array_equal(cur["mask_left"], g["mask_left"])


def test_poseonly_oracle_recovers_true_pose_and_quirks():
    pb = scenes.scene_poseonly_batch(n_frames=3, n_points=300, seed=4, pixel_sigma=0.0, stereo=True)
    ref = oracle.poseonly_solve_batched(pb.kind, pb.offsets, pb.points, pb.px_left, pb.px_right, pb.intr_left,
                                        pb.intr_right, pb.poses_init, oracle.PoseOnlyOptions(1e-6, 1e-6, 1.5, 2.5, 100),
                                        left_to_right=pb.left_to_right)
    assert np.abs(ref["poses"] - pb.poses_true).max() < 1e-3
    for r in ref["results"]:
        assert r.success and r.converged and 3 <= r.n_iterations <= 12
        assert r.n_summary == r.n_iterations - 1     # the converging trip pushes no OptimizationInfo (:116-147)
    # invalid right pixels (x<0 or y<0) are skipped (:298): making ALL right pixels invalid == mono result
    pr = np.full_like(pb.px_right, -1.0)
    a = oracle.poseonly_solve(1, pb.points[:300], pb.px_left[:300], pr[:300], pb.intr_left, pb.intr_right, pb.poses_init[0],
                              oracle.PoseOnlyOptions(1e-6, 1e-6, 1.5, 2.5, 100), left_to_right=pb.left_to_right)
    b = oracle.poseonly_solve(0, pb.points[:300], pb.px_left[:300], None, pb.intr_left, None, pb.poses_init[0],
                              oracle.PoseOnlyOptions(1e-6, 1e-6, 1.5, 2.5, 100))
    np.testing.assert_allclose(a["pose"], b["pose"], atol=1e-6)
    # max iterations reached -> converged false
    c = oracle.poseonly_solve(0, pb.points[:300], pb.px_left[:300], None, pb.intr_left, None, pb.poses_init[0],
                              oracle.PoseOnlyOptions(0.0, 0.0, 1.5, 2.5, 4))
    assert c["result"].n_iterations == 4 and not c["result"].converged


def test_refactor_gauss_newton_and_gradient_descent_branches(oracle_mod):
    """Oracle restatement of FullBundleAdjustmentSolverRefactor (refactor.cpp:944-982, 1075-1367): Gauss-Newton keeps
    every step with lambda untouched (it stays a damped step with initial_lambda) and, lightly damped, reaches the LM
    minimum on a well-posed scene; gradient descent moves every block by at most 0.001 (scaled units) per iteration
    and never rejects."""
    import oracle
    from bundle_adjustment_solver_b200 import scenes
    from bundle_adjustment_solver_b200 import solver as S
    sc = scenes.scene_trajectory(12, 200, 6, stereo=True, seed=3, n_fixed=2)
    res = {}
    for method, iters, lam in ((0, 300, 100.0), (1, 300, 1e-3), (2, 15, 100.0)):
        o = S.load_scene(oracle.FullBAOracle(), sc)
        infos, conv = o.solve(oracle.default_full_options(max_num_iterations=iters, threshold_cost_change=1e-6,
                                                          threshold_step_size=1e-6, method=method, initial_lambda=lam))
        res[method] = (infos, conv)
    lm, gn, gd = res[0][0], res[1][0], res[2][0]
    assert all(i.iteration_status == 0 and abs(i.damping_term - 1e-3) < 1e-9 for i in gn)
    assert all(i.iteration_status == 0 and i.damping_term == 100.0 for i in gd)
    assert gn[-1].cost <= 1.01 * lm[-1].cost + 1e-6, (gn[-1].cost, lm[-1].cost)
    n_blocks = 10 + 200
    assert all(i.abs_step <= (0.02 + 0.001 * n_blocks) / n_blocks * (1 + 1e-12) for i in gd)
    assert gd[-1].cost < 0.5 * gd[0].cost                        # clipped steps along the negative gradient
    assert len(gd) == 15 and not res[2][1]      # max iterations reached -> convergence flag forced false
