#!/usr/bin/env python
"""One-off stress of the bit-exactness claim: a large mixed batch (dispatch order, early helpers,
tail helpers and speculation all active) against the warp-order oracle on every scene."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trajtrack_mpcndqn_rlboost_b200 as t
from tests import oracle_lib as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg = t.Configurator().to_ttmpc()
p = np.concatenate([t.scenes.make_scenes(n // 2, cfg, seed=101, n_static=4, n_dynamic=0, blocking_fraction=0.15),
                    t.scenes.make_scenes(n - n // 2, cfg, seed=202, n_static=3, n_dynamic=4, blocking_fraction=0.15)])
rng = np.random.default_rng(0); p = p[rng.permutation(len(p))]
s = t.BatchSolver(cfg)
t0 = time.perf_counter(); a = s.run(p); t1 = time.perf_counter()
b = s.run(p)
ref = O.solve_batch(cfg, p, threads=os.cpu_count(), warp=True); t2 = time.perf_counter()
same = np.array([np.array_equal(a.solution[i], ref["u"][i]) and a.cost[i] == ref["cost"][i] for i in range(n)])
print(f"{n} scenes: GPU {t1 - t0:.2f} s, oracle {t2 - t1:.1f} s; bit-identical to the oracle {int(same.sum())}/{n}; "
      f"status equal {int((a.exit_status == ref['exit_status']).sum())}; inner equal {int((a.num_inner_iterations == ref['inner']).sum())}; "
      f"y equal {int(np.all(a.lagrange_multipliers == ref['y'], axis=1).sum())}; run-to-run identical {bool(np.array_equal(a.solution, b.solution))}")
