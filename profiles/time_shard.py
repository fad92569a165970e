"""One rank's share of C4 sharded over 8 GPUs, alone on one GPU (no communicator): per-phase times and, under ncu,
the launch list -- what the per-rank work of the strong-scaling run costs without the exchange."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundle_adjustment_solver_b200 import capi, scenes, sharding
from bundle_adjustment_solver_b200 import solver as S

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sc = sharding.shard_scene(scenes.scene_c4(seed=100), 0, world)
e = S.load_scene(S.FullBundleAdjustmentSolver(device=0), sc)
e._upload()
L = capi.lib()
for profile in (False, True):
    e.set_profile(profile)
    for rep in range(2):
        opt = capi.default_options(max_num_iterations=20, threshold_cost_change=0.0, threshold_step_size=0.0, check_every=20)
        res = capi.Result()
        assert L.ba_solve(e.h, C.byref(opt), None, 0, C.byref(res)) == 0
    print(f"world {world} profile {profile}: {res.device_time_ms / 20:.4f} ms per iteration | lin {res.t_linearize_ms / 20:.4f} schur {res.t_schur_ms / 20:.4f} "
          f"solve {res.t_solve_ms / 20:.4f} backsub {res.t_backsub_ms / 20:.4f} update {res.t_update_cost_ms / 20:.4f}", flush=True)
