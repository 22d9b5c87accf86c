"""Diagnostic (build with -DTTMPC_PROFILE -DTTMPC_PROFILE_STAGE): cycles a warp spends in stage_scene per scene,
with the parameter block resident in L2 and after an L2 flush."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trajtrack_mpcndqn_rlboost_b200 as t
from trajtrack_mpcndqn_rlboost_b200 import _lib
name = sys.argv[1] if len(sys.argv) > 1 else "static4096"
w = t.scenes.WORKLOADS[name]
cfg = t.Configurator().to_ttmpc(**w["solver"])
p = t.scenes.make_scenes(w["n"], cfg, seed=1000, n_static=w["n_static"], n_dynamic=w["n_dynamic"], blocking_fraction=w["blocking_fraction"])
s = t.BatchSolver(cfg); lib = _lib.load()
lib.ttmpc_read_stats24.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
o = (C.c_ulonglong * 24)()
dp = torch.from_numpy(p).cuda(); bufs = s.alloc_device(len(p))
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for mode in ("warm L2", "flushed L2", "flushed L2"):
    s.run_device(dp, bufs); torch.cuda.synchronize(); lib.ttmpc_read_stats24(o, 1)
    if mode.startswith("flushed"): flush.zero_(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); s.run_device(dp, bufs); e1.record(); torch.cuda.synchronize(); lib.ttmpc_read_stats24(o, 1)
    print(f"{name} {mode}: kernel {e0.elapsed_time(e1):.2f} ms; staging {o[8+9]/len(p):.0f} (part1 {o[8+8]/len(p):.0f}, part2 {o[8+7]/len(p):.0f}) cycles per scene = {o[8+9]/len(p)/1.965e3:.1f} us; solve {o[7]/len(p)/1.965e3:.0f} us per scene")
