import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trajtrack_mpcndqn_rlboost_b200 as t
cfg = t.Configurator().to_ttmpc()
solver = t.BatchSolver(cfg)
ps = [t.scenes.make_scenes(2500, cfg, seed=70 + j, n_static=4, n_dynamic=j % 2, blocking_fraction=0.1) for j in range(3)]
alone = [solver.run(p) for p in ps]
many = solver.run_many(ps + ps, depth=3)
for j, sol in enumerate(many):
    a = alone[j % 3]
    bad = np.nonzero(~np.all(sol.solution == a.solution, axis=1))[0]
    zero = np.nonzero(np.all(sol.solution == 0, axis=1) & ~np.all(a.solution == 0, axis=1))[0]
    print(j, "mismatch rows", len(bad), "zero rows", len(zero), bad[:8], "status", np.bincount(sol.exit_status, minlength=4), "inner==0:", int((sol.num_inner_iterations == 0).sum()))
