#!/usr/bin/env python
"""Golden vectors for the DQN side of the hybrid loop, produced by the reference's own code:

  MobileRobot.step / step_with_ref_speed        src/pkg_dqn/environment/agent.py (module imported as is)
  rl_ref loop                                    src/main.py:184-193 (restated: 3 lines)
  SpeedObservation, AngularVelocityObservation, ReferencePathSampleObservation,
  ReferencePathCornerObservation .internal_obs() components/int_obsv_*.py (class bodies exec'd as is)
  normalize, normalize_distance                  components/utils.py (exec'd as is)

shapely / gym / matplotlib are absent here: `shapely.geometry.Point` is a stand-in that only
stores the coordinates (agent.py never computes with it) and the path is a small LineString
stand-in with GEOS' project / interpolate semantics -- those two functions are therefore NOT
pinned by this file.  Writes tests/golden/dqn_loop.npz.
"""
import ast
import copy
import importlib.util
import math
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/pkg_dqn/environment"


class Point:
    def __init__(self, xy):
        self.coords = [tuple(float(v) for v in xy)]


class LineString:
    """Length-indexed line: project = arc length of the closest point (first closest segment),
    interpolate = point at an arc length."""
    def __init__(self, coords):
        self.coords = [tuple(map(float, c)) for c in coords]

    def project(self, pt):
        px, py = pt.coords[0]
        best, best_s, cum = math.inf, 0.0, 0.0
        for (ax, ay), (bx, by) in zip(self.coords[:-1], self.coords[1:]):
            dx, dy = bx - ax, by - ay
            len2 = dx * dx + dy * dy
            t = 0.0
            if len2 > 0.0:
                t = min(1.0, max(0.0, ((px - ax) * dx + (py - ay) * dy) / len2))
            d = math.sqrt((ax + t * dx - px) ** 2 + (ay + t * dy - py) ** 2)
            if d < best:
                best, best_s = d, cum + t * math.sqrt(len2)
            cum += math.sqrt(len2)
        return best_s

    def interpolate(self, s):
        if s <= 0.0:
            return Point(self.coords[0])
        cum = 0.0
        for (ax, ay), (bx, by) in zip(self.coords[:-1], self.coords[1:]):
            ln = math.hypot(bx - ax, by - ay)
            if s <= cum + ln and ln > 0.0:
                t = (s - cum) / ln
                return Point((ax + t * (bx - ax), ay + t * (by - ay)))
            cum += ln
        return Point(self.coords[-1])


def load_agent_module():
    sh = types.ModuleType("shapely"); geo = types.ModuleType("shapely.geometry")
    geo.Point = Point; sh.geometry = geo
    sys.modules["shapely"] = sh; sys.modules["shapely.geometry"] = geo
    spec = importlib.util.spec_from_file_location("ref_agent", os.path.join(REF, "agent.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_components():
    ns = {}
    exec(compile(open(os.path.join(REF, "components", "utils.py")).read(), "utils.py", "exec"), ns)

    class Component:
        env = None
    out = {}
    for fname, cls in [("int_obsv_speed.py", "SpeedObservation"), ("int_obsv_angular_velocity.py", "AngularVelocityObservation"),
                       ("int_obsv_reference_path_sample.py", "ReferencePathSampleObservation"),
                       ("int_obsv_reference_path_corner.py", "ReferencePathCornerObservation")]:
        tree = ast.parse(open(os.path.join(REF, "components", fname)).read())
        node = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][0]
        env = dict(Component=Component, normalize=ns["normalize"], normalize_distance=ns["normalize_distance"],
                   np=np, npt=types.SimpleNamespace(ArrayLike=object), atan2=math.atan2, cos=math.cos, sin=math.sin)
        exec(compile(ast.Module(body=[node], type_ignores=[]), fname, "exec"), env)
        out[cls] = env[cls]
    return out


def main():
    agent_mod = load_agent_module()
    comps = load_components()
    rng = np.random.default_rng(3)
    paths = [[(1.0, 1.0), (2.0, 5.0), (6.0, 6.0), (8.0, 8.0)],
             [(0.6, 3.5), (15.4, 3.5)],
             [(18.9, 7.0), (24.0, 12.0), (24.5, 12.2), (30.0, 20.0), (44.7, 6.8)]]
    chain = [comps["SpeedObservation"](), comps["AngularVelocityObservation"](),
             comps["ReferencePathSampleObservation"](1, 0, 0), comps["ReferencePathCornerObservation"](3)]
    A, PI, OBS, PROG, ACT, RL = [], [], [], [], [], []
    for k in range(48):
        pi = k % len(paths)
        path = LineString(paths[pi])
        s = rng.uniform(0, 1)
        nodes = np.array(paths[pi])
        seg = rng.integers(0, len(nodes) - 1)
        pos = nodes[seg] + s * (nodes[seg + 1] - nodes[seg]) + rng.normal(0, 0.4, 2)
        state = np.array([pos[0], pos[1], rng.uniform(-3.1, 3.1), rng.uniform(-0.5, 1.5), rng.uniform(-0.5, 0.5)])
        robot = agent_mod.MobileRobot(state.copy())
        env = types.SimpleNamespace(agent=robot, path=path)
        env.path_progress = path.project(robot.point)
        for c in chain:
            c.env = env
        obs = np.hstack([np.asarray(c.internal_obs(), dtype=np.float32) for c in chain])   # environment.py:153
        action = int(rng.integers(0, 9))
        sim = copy.deepcopy(robot)                                                        # main.py:184-193
        rl = []
        for j in range(20):
            if j == 0:
                sim.step(action, 0.2)
            else:
                sim.step_with_ref_speed(0.2, 1.0)
            rl.append(list(sim.position))
        A.append(state); PI.append(pi); OBS.append(obs); PROG.append(env.path_progress); ACT.append(action); RL.append(rl)
    out = dict(agent=np.array(A), path_index=np.array(PI), internal=np.array(OBS, dtype=np.float32),
               progress=np.array(PROG), action=np.array(ACT), rl_ref=np.array(RL))
    for i, p in enumerate(paths):
        out[f"path_{i}"] = np.array(p)
    dst = os.path.join(ROOT, "tests", "golden", "dqn_loop.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
