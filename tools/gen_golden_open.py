#!/usr/bin/env python
"""Record golden solves of the reference's REAL solver (OpEn PANOC + ALM, Rust) -- the recipe.

This container has neither `opengen` nor a Rust toolchain, so the PANOC / ALM part of the oracle is
"parity unpinned" (oracle/ttmpc_oracle.h).  This script is what pins it: run it ONCE on a machine
that has the reference's requirements (`pip install opengen==0.7.1 casadi`, `rustup` / cargo):

    python tools/gen_golden_open.py --reference /path/to/TrajTrack-MPCnDQN-RLBoost

It (1) builds the solver exactly as the reference does -- `MpcModule(config).build(unicycle_model)`
(src/mpc_traj_tracker/mpc/mpc_generator.py:160-283, called at src/test_block_mpc.py:36) into
`<build_directory>/<optimizer_name>`, (2) loads it like `TrajectoryGenerator.__import_solver`
(trajectory_generator.py:62-76: `__import__(optimizer_name).solver()`), (3) runs
`solver.run(p, initial_guess, initial_lagrange_multipliers, initial_penalty)`
(trajectory_generator.py:284) on this repo's seeded synthetic scenes and (4) writes
`tests/golden/open_solve.npz`: per case the controls, cost, exit status, iteration counts,
infeasibilities, penalty and Lagrange multipliers OpEn returned.  The parameter vectors are NOT
stored (21 KB each): the file keeps (workload, seed, index) and a SHA-256 of the bytes of p, and
`tests/test_open_golden.py` regenerates them and checks the digest.

Cases (all from `scenes.make_scenes`, default `config/mpc_default.yaml` shapes):
  * `static`  : 4 static polygons, cold start, fresh solver object per scene     (BASELINE configs[1])
  * `mixed`   : 4 static + 3 moving ellipses, cold start, fresh solver object per scene
  * `warm`    : the `mixed` scenes again with initial_guess = the recorded solution shifted by
                one step and initial_lagrange_multipliers = the recorded multipliers
  * `sequence`: K consecutive `run()` calls on ONE solver object without multipliers, which pins
                the "Solver object keeps y between calls" behaviour the host mirror reproduces

`record(make_solver, ...)` takes the solver factory as an argument, so the whole recipe can be
exercised end to end against any object with OpEn's `run()` interface: tests/test_open_golden.py
does that with a stand-in backed by the CPU oracle (`--stub` is NOT offered here: tools/ never
loads oracle/).
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EXIT_CODES = {"Converged": 0, "NotConvergedIterations": 1, "NotConvergedOutOfTime": 2}
DEFAULT_OUT = os.path.join(ROOT, "tests", "golden", "open_solve.npz")
CASES = (("static", dict(n_static=4, n_dynamic=0)), ("mixed", dict(n_static=4, n_dynamic=3)))


def scenes_for(case: str, n: int, seed: int):
    """The parameter vectors of one case (this repo's seeded generator, default shapes)."""
    import trajtrack_mpcndqn_rlboost_b200 as t
    cfg = t.Configurator().to_ttmpc()
    kw = dict(CASES)[case]
    return t.scenes.make_scenes(n, cfg, seed=seed, blocking_fraction=0.1, **kw)


def digest(p: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(p, np.float64).tobytes()).hexdigest()


def _status_code(s) -> int:
    return EXIT_CODES.get(str(s), 3)


def _one(sol, nu, n1):
    """Fields of OpEn's Python `OptimizerSolution` (None = the binding reported a solver error)."""
    if sol is None:
        return dict(u=np.full(nu, np.nan), cost=np.nan, exit_status=3, outer=0, inner=0, fpr=np.nan,
                    f1=np.nan, f2=np.nan, pen=np.nan, y=np.full(n1, np.nan), ms=np.nan)
    return dict(u=np.asarray(sol.solution, np.float64), cost=float(sol.cost),
                exit_status=_status_code(sol.exit_status), outer=int(sol.num_outer_iterations),
                inner=int(sol.num_inner_iterations), fpr=float(sol.last_problem_norm_fpr),
                f1=float(sol.f1_infeasibility), f2=float(sol.f2_norm), pen=float(sol.penalty),
                y=np.asarray(sol.lagrange_multipliers, np.float64), ms=float(sol.solve_time_ms))


def _stack(rows):
    return {k: np.stack([np.asarray(r[k]) for r in rows]) for k in rows[0]}


def record(make_solver, n: int = 256, seed: int = 1000, seq_len: int = 8, out: str = DEFAULT_OUT,
           meta: dict | None = None):
    """Run the cases through `make_solver()` objects and write the fixture.  Returns the dict saved."""
    data = dict(meta_n=np.int64(n), meta_seed=np.int64(seed), meta_seq_len=np.int64(seq_len))
    for k, v in (meta or {}).items():
        data["meta_" + k] = np.asarray(str(v))
    recorded = {}
    for case, _ in CASES:
        p = scenes_for(case, n, seed)
        nu = n1 = None
        rows = []
        for i in range(n):
            s = make_solver()                       # fresh object: y = 0, like a first call
            sol = s.run(p=list(map(float, p[i])))   # cold start, as trajectory_generator.py:284
            if nu is None and sol is not None:
                nu, n1 = len(sol.solution), len(sol.lagrange_multipliers)
            rows.append(sol)
        nu = nu or 40; n1 = n1 or 40
        r = _stack([_one(s, nu, n1) for s in rows])
        recorded[case] = (p, r)
        data[case + "_sha256"] = np.asarray(digest(p))
        for k, v in r.items():
            data[f"{case}_{k}"] = v
    # warm starts: previous solution shifted by one step + the recorded multipliers
    p, r = recorded["mixed"]
    nw = min(n, 64)
    u0 = np.concatenate([r["u"][:nw, 2:], r["u"][:nw, -2:]], axis=1)
    y0 = r["y"][:nw]
    ok = np.isfinite(u0).all(axis=1) & np.isfinite(y0).all(axis=1)
    rows = []
    for i in range(nw):
        if not ok[i]:
            rows.append(None); continue
        s = make_solver()
        rows.append(s.run(p=list(map(float, p[i])), initial_guess=list(map(float, u0[i])),
                          initial_lagrange_multipliers=list(map(float, y0[i]))))
    w = _stack([_one(s, u0.shape[1], y0.shape[1]) for s in rows])
    data["warm_u0"] = u0; data["warm_y0"] = y0; data["warm_valid"] = ok
    for k, v in w.items():
        data[f"warm_{k}"] = v
    # sequence on ONE object: the multipliers of call k are the starting point of call k + 1
    p, _ = recorded["static"]
    s = make_solver()
    rows = [s.run(p=list(map(float, p[i]))) for i in range(min(seq_len, n))]
    q = _stack([_one(x, 40, 40) for x in rows])
    for k, v in q.items():
        data[f"sequence_{k}"] = v
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    np.savez_compressed(out, **data)
    return data


def build_reference_solver(reference: str, config_yaml: str, skip_build: bool):
    """`MpcModule(config).build(unicycle_model)` + the import TrajectoryGenerator does."""
    src = os.path.join(reference, "src")
    sys.path.insert(0, src)
    from util.mpc_config import Configurator                      # reference's own loader
    from mpc_traj_tracker.mpc.mpc_generator import MpcModule      # needs casadi + opengen
    from pkg_motion_model import motion_model
    cfg = Configurator(os.path.join(reference, "config", config_yaml), verbose=False)
    cwd = os.getcwd()
    os.chdir(reference)        # build_directory is relative, the reference runs from its root
    try:
        if not skip_build:
            MpcModule(cfg).build(motion_model.unicycle_model)     # cargo build of the generated crate
        solver_path = os.path.join(reference, cfg.build_directory, cfg.optimizer_name)
        sys.path.append(solver_path)
        mod = importlib.import_module(cfg.optimizer_name)
    finally:
        os.chdir(cwd)
    return mod.solver, cfg


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", required=True, help="checkout of Woodenonez/TrajTrack-MPCnDQN-RLBoost")
    ap.add_argument("--config", default="mpc_default.yaml")
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--skip-build", action="store_true", help="the generated solver is already built")
    ap.add_argument("--out", default=DEFAULT_OUT)
    a = ap.parse_args()
    make_solver, _ = build_reference_solver(os.path.abspath(a.reference), a.config, a.skip_build)
    try:
        import opengen
        ogv = getattr(opengen, "__version__", "?")
    except Exception:
        ogv = "?"
    d = record(make_solver, a.n, a.seed, out=a.out, meta=dict(source="OpEn", opengen=ogv, config=a.config))
    for case, _ in CASES:
        st = d[f"{case}_exit_status"]
        print(f"{case}: {len(st)} solves, exit status histogram {np.bincount(st, minlength=4).tolist()}")
    print("wrote", a.out, "-- commit it; tests/test_open_golden.py consumes it")


if __name__ == "__main__":
    main()
