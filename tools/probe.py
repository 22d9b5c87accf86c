import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trajtrack_mpcndqn_rlboost_b200 as t
from trajtrack_mpcndqn_rlboost_b200 import _lib
cfg = t.Configurator().to_ttmpc()
p = t.scenes.make_scenes(4, cfg, seed=1000, n_static=4, n_dynamic=0)
out = (C.c_longlong * 8)()
lib = _lib.load()
for rep in range(2):
    _lib.check(lib.ttmpc_probe_latency(C.byref(cfg), p[0].ctypes.data, out, 200), "probe")
print(dict(zip(["cost_eval", "grad_eval", "wsum", "lbfgs_apply10", "div", "sqrt", "dfma", "chk"], list(out))))
