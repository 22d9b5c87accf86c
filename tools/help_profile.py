"""Diagnostic (build with -DTTMPC_PROFILE -DTTMPC_PROFILE_HELP): owner-side timeline of a PANOC
iteration when a helper is attached.  One scene per launch, so the CTA's idle warps help."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trajtrack_mpcndqn_rlboost_b200 as t
from trajtrack_mpcndqn_rlboost_b200 import _lib
cfg = t.Configurator().to_ttmpc()
p = t.scenes.make_scenes(4096, cfg, seed=1000, n_static=4, n_dynamic=0, blocking_fraction=0.1)
idx = [int(a) for a in sys.argv[1:]] or [593, 1232, 1946]
s = t.BatchSolver(cfg); lib = _lib.load()
lib.ttmpc_read_stats24.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
o = (C.c_ulonglong * 24)()
for i in idx:
    dp = torch.from_numpy(p[i:i + 1]).cuda(); bufs = s.alloc_device(1)
    s.run_device(dp, bufs); torch.cuda.synchronize(); lib.ttmpc_read_stats24(o, 1)
    s.run_device(dp, bufs); torch.cuda.synchronize(); lib.ttmpc_read_stats24(o, 1)
    it = max(o[3], 1); ev = o[0] + o[1]
    print(f"scene {i}: iters {o[3]} evals {ev} ({ev/it:.2f}/iter) cycles/iter total {o[7]/it:.0f}: step {o[8+3]/it:.0f} = post {o[8+0]/it:.0f} + spec L-BFGS {o[8+1]/it:.0f} + wait helper {o[8+2]/it:.0f} + rest {(o[8+3]-o[8+0]-o[8+1]-o[8+2])/it:.0f}; local cost evals {o[4]/it:.0f} grad evals {o[5]/it:.0f} cycles/iter")
