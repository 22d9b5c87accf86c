"""Tiny driver for ncu: solve the static4096 workload a few times on device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import trajtrack_mpcndqn_rlboost_b200 as t
name = sys.argv[1] if len(sys.argv) > 1 else "static4096"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = dict(t.scenes.WORKLOADS[name])
if len(sys.argv) > 3: w["n"] = int(sys.argv[3])
cfg = t.Configurator().to_ttmpc(**w["solver"])
p = t.scenes.make_scenes(w["n"], cfg, seed=1000, n_static=w["n_static"], n_dynamic=w["n_dynamic"],
                         blocking_fraction=w["blocking_fraction"])
s = t.BatchSolver(cfg)
dp = torch.from_numpy(p).cuda()
bufs = s.alloc_device(len(p))
s.read_stats(reset=True)
for _ in range(reps):
    s.run_device(dp, bufs)
torch.cuda.synchronize()
st = s.read_stats()
print("done", st)
print("EVALS_PER_LAUNCH", (st["cost_evals"] + st["grad_evals"]) / reps, "PANOC_ITERS_PER_LAUNCH", st["panoc_iters"] / reps)
