"""Reduced-system solve (K5) on a scene whose S is dense (loop closures): correctness against numpy and
in-situ timing of the multi-kernel blocked Cholesky (DMMA TRSM / SYRK).  usage: time_dense_solve.py n_poses n_points"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundle_adjustment_solver_b200 import capi, scenes
from bundle_adjustment_solver_b200 import solver as S

n_poses, n_points = int(sys.argv[1]), int(sys.argv[2])
check = len(sys.argv) > 3 and sys.argv[3] == "check"
sc = scenes.scene_trajectory(n_poses, n_points, 5, stereo=False, seed=3, heavy_tail=True, loop_fraction=0.05)
e = S.load_scene(S.FullBundleAdjustmentSolver(device=0), sc)
e._upload()
opt = capi.default_options()
e.set_debug(check)
t0 = time.perf_counter()
e.build_only(opt, 100.0, do_solve=True)
print("build+solve wall %.1f ms" % (1e3 * (time.perf_counter() - t0)))
sz = e.sizes()
n = 6 * sz["N"]
if check:
    Sm = e.dump("S").reshape(n, n)
    rhs = e.dump("rhs")
    x = e.dump("x")
    xr = np.linalg.solve(Sm, rhs)
    print("n", n, "x rel err", np.abs(x - xr).max() / np.abs(xr).max())
L = capi.lib()
for name, parts in (("all", 15), ("diag", 1), ("trsm", 2), ("syrk", 4), ("backward", 8)):
    ms = C.c_float(0)
    e.build_only(opt, 100.0, do_solve=False)
    rc = L.ba_debug_time_solve(e.h, parts, 3, C.byref(ms))
    extra = "  %.2f TFLOP/s (n^3/3)" % (n ** 3 / 3 / (ms.value * 1e-3) / 1e12) if parts in (15, 4) else ""
    print(f"n={n} {name:10s} {ms.value:9.3f} ms rc={rc}{extra}")
