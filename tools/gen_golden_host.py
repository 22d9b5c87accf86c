#!/usr/bin/env python
"""Golden vectors for the host-side logic around the solver call, produced by
importing the reference's own Python (numpy/scipy parts) in this container:

  TrajectoryGenerator.get_global_ref_traj / get_local_ref_traj
      /root/reference/src/mpc_traj_tracker/trajectory_generator.py:160-230
  unicycle_model (numpy branch)   src/pkg_motion_model/motion_model.py:153-176
  polygon_halfspace_representation src/util/utils_geo.py:33-59

casadi / opengen are absent: empty stand-in modules satisfy the imports (none of
the functions above touches them).  Writes tests/golden/host_logic.npz.
"""
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

from mpc_traj_tracker.trajectory_generator import TrajectoryGenerator  # noqa: E402
from mpc_traj_tracker._path import PathNodeList  # noqa: E402
from pkg_motion_model.motion_model import unicycle_model  # noqa: E402
from util.utils_geo import polygon_halfspace_representation  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    out = {}
    paths = [
        [(0.6, 3.5), (15.4, 3.5)],
        [(1.0, 1.0), (2.0, 5.0), (6.0, 6.0), (8.0, 8.0)],
        [(18.9, 7.0), (24.0, 12.0), (24.5, 12.2), (30.0, 20.0), (44.7, 6.8)],
    ]
    starts = [(0.6, 3.5, 0.0), (1.0, 1.2, 0.3), (18.0, 6.5, 0.7)]
    for i, (path, st) in enumerate(zip(paths, starts)):
        for j, speed in enumerate([0.3, 1.2, 1.5]):
            traj = TrajectoryGenerator.get_global_ref_traj(0.2, PathNodeList.from_tuples(path), st, speed)
            g = traj.numpy()
            out[f"gref_{i}_{j}"] = g
            out[f"gref_{i}_{j}_path"] = np.array(path)
            out[f"gref_{i}_{j}_state"] = np.array(st)
            out[f"gref_{i}_{j}_speed"] = np.array(speed)
            # local windows at a few indices / robot positions
            loc = []
            for idx in [0, 3, max(0, len(g) - 25), max(0, len(g) - 5)]:
                pos = g[min(idx + 1, len(g) - 1), :2] + rng.normal(0, 0.1, 2)
                lt, nxt = TrajectoryGenerator.get_local_ref_traj(idx, traj, (pos[0], pos[1], 0.0), 1, 20)
                loc.append(np.concatenate([[idx, nxt], pos, lt.reshape(-1)]))
            out[f"lref_{i}_{j}"] = np.array(loc)
    S = rng.uniform(-3, 3, (32, 3)); A = rng.uniform(-1.5, 1.5, (32, 2))
    out["uni_state"], out["uni_action"] = S, A
    out["uni_next"] = np.array([unicycle_model(s, a, 0.2) for s, a in zip(S, A)])
    polys = [np.array([(3.0, 3.0), (3.0, 7.0), (7.0, 7.0), (7.0, 3.0)]),
             np.array([(35.7, 52.2), (48.2, 52.3), (48.7, 13.6), (36.1, 13.8)]),
             np.array([(0.0, 0.0), (2.0, 0.5), (2.5, 2.0), (0.5, 1.5)])]
    for i, poly in enumerate(polys):
        b, a0, a1 = polygon_halfspace_representation(poly)
        out[f"hs_{i}_poly"] = poly
        out[f"hs_{i}"] = np.array([b, a0, a1])
    path = os.path.join(ROOT, "tests", "golden", "host_logic.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
