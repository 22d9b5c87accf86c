"""Latency of ONE solve on the GPU (the drop-in Solver.run use): device-resident and host path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trajtrack_mpcndqn_rlboost_b200 as t
from tests import oracle_lib as O
cfg = t.Configurator().to_ttmpc()
p = t.scenes.make_scenes(64, cfg, seed=1000, n_static=4, n_dynamic=0, blocking_fraction=0.1)
s = t.BatchSolver(cfg)
dp = torch.from_numpy(p).cuda(); bufs = s.alloc_device(1)
dev_ms, host_ms, cpu_ms, iters = [], [], [], []
for i in range(64):
    for rep in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); s.run_device(dp[i:i + 1], bufs); e1.record(); torch.cuda.synchronize()
    dev_ms.append(e0.elapsed_time(e1)); iters.append(int(bufs["inner"][0]))
    t0 = time.perf_counter(); s.run(p[i:i + 1]); host_ms.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter(); O.solve_batch(cfg, p[i:i + 1], warp=False); cpu_ms.append((time.perf_counter() - t0) * 1e3)
q = lambda a: [round(float(x), 2) for x in np.quantile(a, [0.5, 0.9, 1.0])]
print("single-scene latency ms (p50, p90, max): device", q(dev_ms), "host API", q(host_ms), "CPU oracle (1 core)", q(cpu_ms))
print("PANOC iterations p50/p90/max", q(iters), "device us per iteration p50", round(float(np.median(np.array(dev_ms) * 1e3 / np.maximum(iters, 1))), 1))
