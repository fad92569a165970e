import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["BA_B200_CHOL_MODE"] = sys.argv[2]
import oracle
from bundle_adjustment_solver_b200 import capi, scenes
from helpers import load_engine, load_oracle
n_poses = int(sys.argv[1])
track = int(sys.argv[3]) if len(sys.argv) > 3 else 6
sc = scenes.scene_trajectory(n_poses, 40 * n_poses, track, stereo=True, seed=1, n_fixed=2)
o = load_oracle(sc); o.build_only(1.0, 100.0, 0, True)
e = load_engine(sc, identical_internal=o.get_internal())
e.set_debug(True)
e.build_only(capi.default_options(), 100.0, do_solve=True)
n = len(o.dump("x"))
S = e.dump("S").reshape(n, n)
rhs = e.dump("rhs")
F = e.dump("factor").reshape(n + 1, n + 1).T   # F[r][c] (column-major in memory)
L = np.tril(F[:n, :n])
Lref = np.linalg.cholesky(S)
z = F[n, :n]
zref = np.linalg.solve(Lref, rhs)
print("track", track, "n", n, "L err", np.abs(L - Lref).max() / np.abs(Lref).max(), "z err", np.abs(z - zref).max() / np.abs(zref).max(),
      "x err", np.abs(e.dump("x") - np.linalg.solve(S, rhs)).max() / np.abs(o.dump("x")).max())
bad = np.argwhere(np.abs(L - Lref) > 1e-9 * np.abs(Lref).max())
print("bad L entries", len(bad), bad[:10].tolist())
ratio = np.where(np.abs(Lref) > 1e-12 * np.abs(Lref).max(), L / np.where(Lref == 0, 1, Lref), np.nan)
