"""In-situ cycle breakdown of the solve kernel (diagnostic build with -DTTMPC_PROFILE).
usage: TTMPC_LIB=build/libttmpc_prof.so python tools/phase_profile.py [workload] [n]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trajtrack_mpcndqn_rlboost_b200 as t
from trajtrack_mpcndqn_rlboost_b200 import _lib
name = sys.argv[1] if len(sys.argv) > 1 else "static4096"
w = t.scenes.WORKLOADS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else w["n"]
cfg = t.Configurator().to_ttmpc(**w["solver"])
p = t.scenes.make_scenes(max(n, w["n"]), cfg, seed=1000, n_static=w["n_static"], n_dynamic=w["n_dynamic"],
                         blocking_fraction=w["blocking_fraction"])[:n]
s = t.BatchSolver(cfg)
lib = _lib.load()
lib.ttmpc_read_stats8.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
dp = torch.from_numpy(p).cuda(); bufs = s.alloc_device(len(p))
out = (C.c_ulonglong * 8)()
for _ in range(2):
    lib.ttmpc_read_stats8(out, 1)
    s.run_device(dp, bufs); torch.cuda.synchronize()
lib.ttmpc_read_stats8(out, 1)
nc, ng, nb, it, cyc_c, cyc_g, cyc_l, cyc_t = list(out)
print(f"scenes {n}: cost evals {nc} grad evals {ng} panoc iters {it}")
print(f"cycles/cost eval {cyc_c/max(nc,1):.0f}  cycles/grad eval {cyc_g/max(ng,1):.0f}  lbfgs cycles/iter {cyc_l/max(it,1):.0f}")
print(f"share of solve cycles: cost evals {100*cyc_c/cyc_t:.1f}%  grad evals {100*cyc_g/cyc_t:.1f}%  lbfgs {100*cyc_l/cyc_t:.1f}%  rest {100*(cyc_t-cyc_c-cyc_g-cyc_l)/cyc_t:.1f}%")
try:
    lib.ttmpc_read_stats24.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    o24 = (C.c_ulonglong * 24)()
    s.run_device(dp, bufs); torch.cuda.synchronize()
    lib.ttmpc_read_stats24(o24, 1)
    ev = o24[0] + o24[1]
    names = ["rollout", "refpath", "speed+fleet", "dynamic", "terminal+static", "accel/ALM", "S+F2", "gradient+adjoint", "reduction"]
    print("lbfgs apply cycles per iteration:", round(o24[17] / max(o24[3], 1)), " update+apply:", round(o24[6] / max(o24[3], 1)))
    print("eval sections, cycles per evaluation:", {n: round(o24[8 + i] / max(ev, 1)) for i, n in enumerate(names)})
except Exception as ex:
    print("no eval section profile:", ex)
print(f"total cycles per panoc iteration {cyc_t/max(it,1):.0f}; evals per iteration {(nc+ng)/max(it,1):.2f}")
