#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer: solve (static + dynamic scenes, sweep shape, long limits),
fleet step with hint switch, DQN observe + tensor-core Q-net."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trajtrack_mpcndqn_rlboost_b200 as t
cfg = t.Configurator().to_ttmpc(max_inner_iterations=40, max_outer_iterations=3)
p = np.concatenate([t.scenes.make_scenes(24, cfg, seed=1, n_static=4, n_dynamic=0),
                    t.scenes.make_scenes(24, cfg, seed=2, n_static=3, n_dynamic=6, blocking_fraction=0.5)])
s = t.BatchSolver(cfg)
r = s.run(p); print("solve", np.bincount(r.exit_status, minlength=4))
big = t.scenes.make_scenes(1400, cfg, seed=3, n_static=4, n_dynamic=3)      # more scenes than resident warps: bulk kernel
r = s.run(big); print("bulk", np.bincount(r.exit_status, minlength=4))
mc = t.Configurator(N_hor=32, Nstcobs=10, Ndynobs=15); c2 = mc.to_ttmpc(max_inner_iterations=30, max_outer_iterations=2)
r = t.BatchSolver(c2).run(t.scenes.make_scenes(16, c2, seed=4, n_static=4, n_dynamic=3, mpc=mc)); print("N=32", np.bincount(r.exit_status, minlength=4))
mc = t.Configurator(N_hor=13, Nstcobs=3, Ndynobs=40); c3 = mc.to_ttmpc(max_inner_iterations=30, max_outer_iterations=2)
r = t.BatchSolver(c3).run(t.scenes.make_scenes(16, c3, seed=5, n_static=3, n_dynamic=9, mpc=mc)); print("runtime dims", np.bincount(r.exit_status, minlength=4))
fl = t.scenes.make_fleet(16, seed=5)
fp = t.FleetPlanner(t.Configurator(), fl["init"], fl["goal"], fl["paths"], mode="work", max_inner_iterations=30, max_outer_iterations=2)
fp.update_static_constraints(fl["static_polys"], per_robot=True); fp.set_moving_obstacles(fl["moving_pos"], fl["moving_disp"])
fp.enable_hint_switch(fl["static_polys"], per_robot=True)
for _ in range(2): fp.step()
torch.cuda.synchronize(); print("fleet ok")
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "qnet_ray.npz"))
w = t.dqn.QNetWeights(*[g[k] for k in ("w0", "b0", "w1", "b1", "w2", "b2")])
lay = t.dqn.default_layout()
rings = [[t.geometry.pad_polygon_round(np.array([(3., 3.), (3., 7.), (7., 7.), (7., 3.)]), 0.5), np.array([(0.5, 0.5), (9.5, 0.5), (9.5, 9.5), (0.5, 9.5)])]] * 37
xy, off, sol, cnt = t.dqn.pack_geometry(lay, rings, [[True, False]] * 37)
out = t.dqn.DqnCompanion(lay, w).observe_act(np.array([[1.0 + 0.1 * i, 2.0, 0.1 * i] for i in range(37)]), xy, off, sol, cnt, g["internal"][:37])
print("dqn", out["action"][:8])
