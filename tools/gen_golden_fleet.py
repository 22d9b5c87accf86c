#!/usr/bin/env python
"""Golden vectors for the fleet step (the caller side of the solve), produced by running the
reference's own, unmodified Python in this container:

  InterfaceMpc.initialization / get_local_ref_traj / get_action   src/interface_mpc.py:52-92
  TrajectoryGenerator.run_step / run_solver / check_termination_condition
      src/mpc_traj_tracker/trajectory_generator.py:158-164, 233-307
  unicycle_model                                                   src/pkg_motion_model/motion_model.py:153-176
  est_dyn_obs_positions (function source taken from src/main.py:80-89, the module itself
      needs gym / stable_baselines3 and cannot be imported here)

Only the OpEn-generated solver module is replaced: a stand-in whose run() records the packed
parameter vector it is given and returns a prescribed control sequence, so that everything
around the solver call executes as in the reference.  Writes tests/golden/fleet_step.npz.
"""
import ast
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cs = types.ModuleType("casadi.casadi"); cs.SX = type("SX", (), {})
pk = types.ModuleType("casadi"); pk.casadi = cs
sys.modules["casadi"] = pk; sys.modules["casadi.casadi"] = cs
ogm = types.ModuleType("opengen.opengen")
ogp = types.ModuleType("opengen"); ogp.opengen = ogm
ogp.tcp = types.SimpleNamespace(solver_status=types.SimpleNamespace(SolverStatus=object))
ogm.tcp = ogp.tcp
sys.modules["opengen"] = ogp; sys.modules["opengen.opengen"] = ogm
sys.path.insert(0, "/root/reference/src")

from interface_mpc import InterfaceMpc  # noqa: E402
from util.mpc_config import Configurator  # noqa: E402


class _Sol:
    def __init__(self, u):
        self.solution = list(u); self.cost = 1.25; self.exit_status = "Converged"; self.solve_time_ms = 0.5


class _FakeSolver:
    """Stands in for the PyO3 object built by opengen; records p, returns the next prescribed u."""
    def __init__(self):
        self.p_log, self.u_next = [], None

    def run(self, p, initial_guess=None, initial_lagrange_multipliers=None, initial_penalty=None):
        self.p_log.append(np.array(p, dtype=np.float64))
        return _Sol(self.u_next)


def reference_est_dyn_obs_positions():
    src = open("/root/reference/src/main.py").read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "est_dyn_obs_positions"][0]
    ns = {"DYN_OBS_SIZE": 0.8 + 0.8}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "main.py", "exec"), ns)
    return ns["est_dyn_obs_positions"]


def main():
    cfg = Configurator("/root/reference/config/mpc_default.yaml")
    fake_mod = types.ModuleType(cfg.optimizer_name)
    solvers = []

    def make():
        s = _FakeSolver(); solvers.append(s); return s
    fake_mod.solver = make
    sys.modules[cfg.optimizer_name] = fake_mod
    est = reference_est_dyn_obs_positions()
    rng = np.random.default_rng(11)
    N = cfg.N_hor
    robots = [
        dict(init=(0.6, 3.5, 0.0), goal=(15.4, 3.5, 0.0), path=[(0.6, 3.5), (15.4, 3.5)]),
        dict(init=(1.0, 1.2, 0.3), goal=(8.0, 8.0, 0.0), path=[(1.0, 1.0), (2.0, 5.0), (6.0, 6.0), (8.0, 8.0)]),
        dict(init=(2.4, 2.2, 0.8), goal=(3.0, 3.0, 0.0), path=[(2.4, 2.2), (3.0, 3.0)]),      # close to the goal
        dict(init=(5.0, 5.0, 0.0), goal=(5.02, 5.01, 0.0), path=[(5.0, 5.0), (5.02, 5.01)]),  # already there
    ]
    stc_polys = [np.array([(3.0, 3.0), (3.0, 7.0), (7.0, 7.0), (7.0, 3.0)]),
                 np.array([(9.0, 1.0), (9.0, 2.5), (11.0, 2.5), (11.0, 1.0)])]
    out = {"n_robots": np.array(len(robots)), "n_steps": np.array(6)}
    for r, rb in enumerate(robots):
        mpc = InterfaceMpc(cfg, motion_model=None)
        fs = solvers[-1]
        mpc.initialization(np.array(rb["init"]), np.array(rb["goal"]), rb["path"], mode="work")
        mpc.update_static_constraints([p.tolist() for p in stc_polys])
        out[f"r{r}_ref_traj"] = mpc.ref_traj.numpy()
        out[f"r{r}_goal"] = np.array(rb["goal"], dtype=np.float64)
        out[f"r{r}_stc"] = np.array(mpc.stc_constraints, dtype=np.float64)
        obs_cur = [[4.0 + r, 8.0], [12.0, 3.0 + 0.5 * r]]
        obs_disp = [[0.11, -0.07], [-0.13, 0.02]]
        obs_last = [list(c) for c in obs_cur]
        out[f"r{r}_obs0"] = np.array(obs_cur); out[f"r{r}_obs_disp"] = np.array(obs_disp)
        for t in range(6):
            pre = dict(state=np.array(mpc.state, dtype=np.float64), last_u=np.array(mpc.last_action, dtype=np.float64),
                       idx=mpc._traj_gen.idx_ref)
            pred = [est(l, c) for l, c in zip(obs_last, obs_cur)]
            mpc.update_dynamic_constraints(pred)
            ref_local, _ = mpc.get_local_ref_traj()
            u = np.concatenate([np.stack([rng.uniform(0.2, 1.4, N), rng.uniform(-0.4, 0.4, N)], 1).reshape(-1)])
            if r == 3:
                u[0] = 0.01
            fs.u_next = u
            n_before = len(fs.p_log)
            ret = mpc.get_action(ref_local, mode="work")
            k = f"r{r}_t{t}_"
            out[k + "state"] = pre["state"]; out[k + "last_u"] = pre["last_u"]; out[k + "idx"] = np.array(pre["idx"])
            out[k + "idx_next"] = np.array(mpc._traj_gen.idx_ref)
            out[k + "u"] = u
            out[k + "obs_last"] = np.array(obs_last); out[k + "obs_cur"] = np.array(obs_cur)
            out[k + "dyn_rows"] = np.array(mpc.dyn_constraints, dtype=np.float64)
            if ret is None:
                out[k + "reached"] = np.array(1)
            else:
                assert len(fs.p_log) == n_before + 1
                out[k + "reached"] = np.array(0)
                out[k + "p"] = fs.p_log[-1]
                out[k + "action"] = np.array(ret[0], dtype=np.float64)
                out[k + "pred_states"] = np.array(ret[1], dtype=np.float64)
            out[k + "state_next"] = np.array(mpc.state, dtype=np.float64)
            obs_last = [list(c) for c in obs_cur]
            obs_cur = [[c[0] + d[0], c[1] + d[1]] for c, d in zip(obs_cur, obs_disp)]
    path = os.path.join(ROOT, "tests", "golden", "fleet_step.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays; reached flags:",
          [[int(out[f"r{r}_t{t}_reached"]) for t in range(6)] for r in range(len(robots))])


if __name__ == "__main__":
    main()
