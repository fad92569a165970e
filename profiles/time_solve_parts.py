"""In-situ (warm, graph-replayed) timing of the parts of the reduced solve K5 on config C3."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundle_adjustment_solver_b200 import capi, scenes
from bundle_adjustment_solver_b200 import solver as S

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
sc = scenes.scene_c3(seed=100, pose_noise_seed=7) if wl == "c3" else scenes.scene_test_ba(seed=0)
e = S.load_scene(S.FullBundleAdjustmentSolver(device=0), sc)
e._upload()
opt = capi.default_options()
e.build_only(opt, 100.0, do_solve=True)
L = capi.lib()
for name, parts in (("diag", 1), ("diag+trsm", 3), ("diag+trsm+syrk", 7), ("all", 15), ("backward", 8), ("trsm", 2), ("syrk", 4)):
    ms = C.c_float(0)
    e.build_only(opt, 100.0, do_solve=False)
    rc = L.ba_debug_time_solve(e.h, parts, 20, C.byref(ms))
    print(f"{wl} {name:16s} {ms.value*1000:9.1f} us  rc={rc}")
