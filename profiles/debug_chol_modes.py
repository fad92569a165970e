"""Debug helper: reduced-solve result x of the cluster path vs the multi-kernel path vs the oracle."""
import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

def run(n_poses, mode):
    os.environ["BA_B200_CHOL_MODE"] = str(mode)
    import oracle
    from bundle_adjustment_solver_b200 import capi, scenes
    from helpers import load_engine, load_oracle
    sc = scenes.scene_trajectory(n_poses, 40 * n_poses, 6, stereo=True, seed=1, n_fixed=2) if n_poses > 0 else scenes.scene_test_ba(seed=0)
    o = load_oracle(sc); o.build_only(1.0, 100.0, 0, True)
    e = load_engine(sc, identical_internal=o.get_internal())
    e.build_only(capi.default_options(), 100.0, do_solve=True)
    xo, xe = o.dump("x"), e.dump("x")
    err = np.abs(xe - xo).reshape(-1, 6).max(1) / np.abs(xo).max()
    print(f"poses={n_poses} n={len(xo)} mode={mode} max rel err={err.max():.3e} first bad block={int(np.argmax(err > 1e-6)) if (err > 1e-6).any() else -1} nan={np.isnan(xe).sum()}")

if __name__ == "__main__":
    if len(sys.argv) > 2:
        run(int(sys.argv[1]), int(sys.argv[2]))
    else:
        for npz in (8, 13, 14, 24, 40, 0):
            for mode in (0, 1):
                subprocess.call([sys.executable, __file__, str(npz), str(mode)])
