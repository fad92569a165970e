"""Where the end-to-end time of one solve from host buffers goes (C3): Python mirror vs C-ABI stages."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bundle_adjustment_solver_b200 import capi, scenes
from bundle_adjustment_solver_b200 import solver as S
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")
sc = scenes.scene_c3(seed=100, pose_noise_seed=7)
for rep in range(3):
    t0 = time.perf_counter()
    e = S.FullBundleAdjustmentSolver(device=0)
    t1 = time.perf_counter()
    S.load_scene(e, sc)
    t2 = time.perf_counter()
    e._upload()
    t3 = time.perf_counter()
    summ = S.Summary()
    e.solve(capi.default_options(max_num_iterations=50, threshold_cost_change=1e-6, threshold_step_size=1e-6), summ)
    t4 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f}  load_scene {1e3*(t2-t1):.1f}  upload+finalize {1e3*(t3-t2):.1f}  "
          f"solve+writeback {1e3*(t4-t3):.1f} ms  iters {len(summ.optimization_info_list)} device_ms {summ.result.device_time_ms:.1f} "
          f"total_ms {summ.result.total_time_ms:.1f}")
    del e
