"""CPU tests (-m "not gpu"): the oracle against the golden vectors produced from the
reference's own code (tools/gen_golden_*.py), the host-side mirror against the
reference's host logic, and the C-ABI export list."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import trajtrack_mpcndqn_rlboost_b200 as t
from trajtrack_mpcndqn_rlboost_b200 import _lib
from tests import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max())


# ------------------------------------------------------------------ problem functions
def test_oracle_matches_reference_problem_functions(cfg, golden_problem):
    """f, grad f, F1, F2 of the oracle vs MpcModule.build run on the casadi shim.
    Tolerance 1e-12 relative (observed 1e-15)."""
    g = golden_problem
    for i in range(len(g["f"])):
        f, F1, F2 = O.evaluate(cfg, g["u"][i], g["p"][i])
        assert abs(f - g["f"][i]) <= 1e-12 * abs(g["f"][i])
        assert np.abs(F1 - g["F1"][i]).max() <= 1e-12
        assert rel(F2, g["F2"][i]) <= 1e-12
        assert rel(O.psi_grad(cfg, g["u"][i], g["p"][i], 0.0), g["grad_f"][i]) <= 1e-12


def test_oracle_matches_reference_psi(cfg, golden_problem):
    g = golden_problem
    for i in range(len(g["f"])):
        ps = O.psi(cfg, g["u"][i], g["p"][i], g["c"][i], g["y"][i])
        gr = O.psi_grad(cfg, g["u"][i], g["p"][i], g["c"][i], g["y"][i])
        assert abs(ps - g["psi"][i]) <= 1e-12 * abs(g["psi"][i])
        assert rel(gr, g["grad_psi"][i]) <= 1e-12


def test_reference_problem_shape(cfg, golden_problem):
    """What build() hands to opengen: box on u, set C, n2 = Ndynobs (scalar+vector
    broadcast of penalty_constraints), initial penalty 10."""
    g = golden_problem
    N = cfg.N_hor
    assert g["F2"].shape[1] == cfg.Ndynobs == 15
    assert np.all(g["umin"][0::2] == cfg.lin_vel_min) and np.all(g["umax"][0::2] == cfg.lin_vel_max)
    assert np.all(g["umin"][1::2] == -cfg.ang_vel_max) and np.all(g["umax"][1::2] == cfg.ang_vel_max)
    assert np.all(g["c_min"][:N] == cfg.lin_acc_min) and np.all(g["c_max"][:N] == cfg.lin_acc_max)
    assert np.all(g["c_min"][N:] == -cfg.ang_acc_max) and np.all(g["c_max"][N:] == cfg.ang_acc_max)
    assert float(g["initial_penalty"]) == cfg.initial_penalty == 10.0
    assert set(g["solver_config_calls"].tolist()) == {"with_initial_penalty", "with_max_duration_micros"}


def test_oracle_gradient_finite_differences(cfg):
    p = t.scenes.make_scenes(6, cfg, seed=21, n_static=4, n_dynamic=5, blocking_fraction=0.5)
    rng = np.random.default_rng(1)
    for i in range(6):
        u = rng.uniform(-0.3, 1.2, 40); u[1::2] = rng.uniform(-0.4, 0.4, 20)
        y = rng.normal(0, 1, 40); c = 50.0
        g = O.psi_grad(cfg, u, p[i], c, y)
        gn = np.zeros(40)
        for k in range(40):
            e = np.zeros(40); e[k] = 1e-6
            gn[k] = (O.psi(cfg, u + e, p[i], c, y) - O.psi(cfg, u - e, p[i], c, y)) / 2e-6
        assert np.abs(g - gn).max() <= 2e-6 * max(1.0, np.abs(g).max())


def test_warp_ordered_eval_matches_reference_order(cfg, golden_problem):
    """Part 1b (GPU operation order) vs the goldens: 1e-12 relative."""
    g = golden_problem
    for i in range(len(g["f"])):
        f, F2, ps, gr = O.eval_warp(cfg, g["u"][i], g["p"][i], g["c"][i], g["y"][i])
        assert abs(f - g["f"][i]) <= 1e-12 * abs(g["f"][i])
        assert abs(ps - g["psi"][i]) <= 1e-12 * abs(g["psi"][i])
        assert rel(F2, g["F2"][i]) <= 1e-12
        assert rel(gr, g["grad_psi"][i]) <= 1e-12


def test_tt_sincos_accuracy():
    """The shared sincos (Cody-Waite + fdlibm kernels) vs libm: <= 2 ulp."""
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-10, 10, 4000), rng.uniform(-1e4, 1e4, 1000),
                         rng.uniform(-1e8, 1e8, 500), [0.0, math.pi / 2, math.pi, -math.pi / 4]])
    for x in xs:
        s, c = O.sincos(x)
        assert abs(s - math.sin(x)) <= 2 * np.spacing(abs(math.sin(x))) + 1e-300
        assert abs(c - math.cos(x)) <= 2 * np.spacing(abs(math.cos(x))) + 1e-300
    s, c = O.sincos(float("inf"))
    assert math.isnan(s) and math.isnan(c)


# ------------------------------------------------------------------ Q-network
def test_oracle_qnet_matches_reference_weights(golden_qnet):
    """Oracle MLP vs torch fp32 run of the reference's trained q_net: |dQ| <= 1e-5,
    identical argmax."""
    g = golden_qnet
    w = t.dqn.QNetWeights(*[g[k] for k in ("w0", "b0", "w1", "b1", "w2", "b2")])
    assert (w.n_in, w.n_h1, w.n_h2, w.n_out) == (46, 16, 16, 9)
    x = np.concatenate([g["ext"], g["internal"]], axis=1)
    h = np.maximum(x @ g["w0"].T + g["b0"], 0)
    h = np.maximum(h @ g["w1"].T + g["b1"], 0)
    q = h @ g["w2"].T + g["b2"]
    assert np.abs(q - g["q"]).max() <= 1e-5
    assert np.array_equal(q.argmax(1), g["action"])


# ------------------------------------------------------------------ host logic
def test_global_ref_traj_matches_reference(golden_host):
    g = golden_host
    for i in range(3):
        for j in range(3):
            key = f"gref_{i}_{j}"
            mine = t.TrajectoryGenerator.get_global_ref_traj(
                0.2, [tuple(x) for x in g[key + "_path"]], g[key + "_state"], float(g[key + "_speed"]))
            assert mine.shape == g[key].shape
            assert np.array_equal(mine, g[key])


def test_local_ref_traj_matches_reference(golden_host):
    g = golden_host
    for i in range(3):
        for j in range(3):
            glob = g[f"gref_{i}_{j}"]
            for row in g[f"lref_{i}_{j}"]:
                idx, nxt, pos, want = int(row[0]), int(row[1]), row[2:4], row[4:].reshape(20, 3)
                got, got_idx = t.TrajectoryGenerator.get_local_ref_traj(idx, glob, (pos[0], pos[1], 0.0), 1, 20)
                assert got_idx == nxt
                assert np.array_equal(got, want)


def test_unicycle_model_matches_reference(golden_host):
    g = golden_host
    for s, a, n in zip(g["uni_state"], g["uni_action"], g["uni_next"]):
        assert np.array_equal(t.unicycle_model(s, a, 0.2), n)


def test_halfspace_representation_matches_reference(golden_host):
    """Same half-spaces as utils_geo.polygon_halfspace_representation (edge order is the
    hull's, so compare as sets of rows)."""
    g = golden_host
    for i in range(3):
        b, a0, a1 = t.geometry.polygon_halfspace_representation(g[f"hs_{i}_poly"])
        mine = sorted(zip(np.round(b, 9), np.round(a0, 9), np.round(a1, 9)))
        ref = sorted(zip(*np.round(g[f"hs_{i}"], 9)))
        assert np.allclose(np.array(mine), np.array(ref), atol=1e-9)
        # inside test: the centroid satisfies every half-space strictly
        c = g[f"hs_{i}_poly"].mean(0)
        assert np.all(np.array(b) - np.array(a0) * c[0] - np.array(a1) * c[1] > 0)


def test_packed_parameter_layout(cfg):
    """assemble_parameters follows trajectory_generator.py:251-254 block for block."""
    off = t.param_offsets(cfg)
    assert off["np"] == t.num_params(cfg) == 2658
    assert (off["q"], off["r"], off["vref"], off["c"], off["os"], off["od"], off["qstc"], off["qdyn"]) == \
        (8, 18, 78, 98, 698, 818, 2618, 2638)


# ------------------------------------------------------------------ C-ABI
def test_cabi_exports_every_declared_symbol():
    """Every function include/ttmpc.h declares is exported by libttmpc.so (no compute call)."""
    hdr = open(os.path.join(ROOT, "include", "ttmpc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(tt(?:mpc|dqn)_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 16
    lib = C.CDLL(_lib.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in ttmpc.h but not exported"
    assert names == set(_lib.SYMBOLS), names ^ set(_lib.SYMBOLS)


def test_cabi_config_helpers(cfg):
    lib = _lib.load()
    d = _lib.default_config()
    for f, _ in _lib.TtmpcConfig._fields_:
        assert getattr(d, f) == getattr(cfg, f), f
    assert lib.ttmpc_num_params(C.byref(cfg)) == 2658
    assert lib.ttmpc_num_decision(C.byref(cfg)) == 40
    assert lib.ttmpc_num_alm(C.byref(cfg)) == 40
    assert lib.ttmpc_num_penalty(C.byref(cfg)) == 15
    assert lib.ttmpc_exit_status_name(0) == b"Converged"
    assert lib.ttmpc_exit_status_name(1) == b"NotConvergedIterations"
    assert lib.ttmpc_exit_status_name(2) == b"NotConvergedOutOfTime"


def test_cabi_rejects_bad_config():
    lib = _lib.load()
    bad = _lib.default_config()
    bad.N_hor = 40
    res = _lib.TtmpcResult()
    rc = lib.ttmpc_solve_batch_device(C.byref(bad), 1, None, 0, 0, None, C.byref(res), None)
    assert rc == -1 and b"N_hor" in lib.ttmpc_last_error()


def test_product_does_not_touch_the_oracle():
    """The product package must not import, load or name anything under oracle/."""
    pkg = os.path.join(ROOT, "trajtrack_mpcndqn_rlboost_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle_lib" not in src and "libttmpc_oracle" not in src, f
