"""Fleet step (caller side of the solve): the C oracle against golden vectors produced by the
reference's own Python (tools/gen_golden_fleet.py -> tests/golden/fleet_step.npz)."""
import os

import numpy as np
import pytest

import trajtrack_mpcndqn_rlboost_b200 as t
from trajtrack_mpcndqn_rlboost_b200.fleet import work_mode
from tests import oracle_lib as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fleet_step.npz"))
NR, NT = int(G["n_robots"]), int(G["n_steps"])


def _host(cfg, mc, r, tt):
    k = f"r{r}_t{tt}_"
    tuning, base = work_mode(mc, "work")
    ref = G[f"r{r}_ref_traj"]
    fh = O.FleetHost(cfg, G[k + "state"][None], G[f"r{r}_goal"][None], ref[None], [len(ref)], G[f"r{r}_stc"],
                     tuning, base, mc.low_speed, dyn_cur=G[k + "obs_cur"][None], dyn_disp=G[f"r{r}_obs_disp"][None])
    fh.dyn_last[:] = G[k + "obs_last"][None]
    fh.last_u[:] = G[k + "last_u"][None]
    fh.idx_ref[:] = int(G[k + "idx"])
    return fh, k


@pytest.mark.parametrize("use_libm", [True, False])
def test_pack_matches_reference_python(use_libm):
    mc = t.Configurator()
    cfg = mc.to_ttmpc()
    off = t.param_offsets(cfg)
    checked = 0
    for r in range(NR):
        for tt in range(NT):
            fh, k = _host(cfg, mc, r, tt)
            p = O.fleet_pack(fh, use_libm)[0]
            assert int(fh.idx_ref[0]) == int(G[k + "idx_next"])
            assert int(fh.status[0]) == int(G[k + "reached"])
            # the dynamic-obstacle rows are est_dyn_obs_positions(last, cur), bit for bit
            lo = off["od"]
            assert np.array_equal(p[lo:lo + len(G[k + "dyn_rows"])], G[k + "dyn_rows"])
            if int(G[k + "reached"]):
                continue
            gp = G[k + "p"]
            mask = np.ones(len(gp), bool)
            mask[off["vref"]:off["c"]] = False
            assert np.array_equal(p[mask], gp[mask])            # everything that is copied: exact
            # speed reference: hypot may differ from CPython's by an ulp
            np.testing.assert_allclose(p[~mask], gp[~mask], rtol=4e-16, atol=0)
            checked += 1
    assert checked >= 12


@pytest.mark.parametrize("use_libm", [True, False])
def test_advance_matches_reference_python(use_libm):
    mc = t.Configurator()
    cfg = mc.to_ttmpc()
    for r in range(NR):
        for tt in range(NT):
            fh, k = _host(cfg, mc, r, tt)
            O.fleet_pack(fh, use_libm)
            O.fleet_advance(fh, G[k + "u"][None], np.zeros(1, np.int32), use_libm)
            np.testing.assert_allclose(fh.state[0], G[k + "state_next"], rtol=0, atol=2e-15)
            if not int(G[k + "reached"]):
                assert np.array_equal(fh.last_u[0], G[k + "action"])
                # obstacles moved on by their displacement
                assert np.array_equal(fh.dyn_last[0], G[k + "obs_cur"])
                np.testing.assert_allclose(fh.dyn_cur[0], G[k + "obs_cur"] + G[f"r{r}_obs_disp"], rtol=0, atol=0)
            else:
                assert np.array_equal(fh.state[0], G[k + "state"])   # get_action returned None: nothing moves


def test_solver_failure_freezes_the_robot():
    mc = t.Configurator()
    cfg = mc.to_ttmpc()
    fh, k = _host(cfg, mc, 0, 0)
    O.fleet_pack(fh, True)
    before = fh.state.copy()
    O.fleet_advance(fh, G[k + "u"][None], np.array([3], np.int32), True)   # NotFiniteComputation
    assert int(fh.status[0]) == 2 and np.array_equal(fh.state, before)
    O.fleet_advance(fh, G[k + "u"][None], np.zeros(1, np.int32), True)     # stays frozen
    assert np.array_equal(fh.state, before)


def test_hint_matches_interface_mpc_semantics():
    """Hybrid mode: InterfaceMpc.get_local_ref_traj(rl_ref) appends the original headings to the
    hint positions (interface_mpc.py:76-78) and ref_traj_filter(decay=1) keeps the hint
    (main.py:35-42); run_step then packs that trajectory (finish_state = its last row)."""
    mc = t.Configurator()
    cfg = mc.to_ttmpc()
    off = t.param_offsets(cfg)
    N = cfg.N_hor
    fh, k = _host(cfg, mc, 1, 2)
    p0 = O.fleet_pack(fh, True)[0].copy()
    fh.idx_ref[:] = int(G[k + "idx"])
    rng = np.random.default_rng(0)
    hint = rng.normal(0, 3, (1, N, 2))
    fh.hint, fh.use_hint = hint, np.ones(1, np.int32)
    p1 = O.fleet_pack(fh, True)[0]
    original = p0[off["r"]:off["r"] + 3 * N].reshape(N, 3)
    local = np.concatenate((hint[0], original[:, [2]]), axis=1)          # interface_mpc.py:77
    filtered = original.copy()                                           # main.py:35-42 with decay = 1
    decay = 1
    for i in range(N):
        filtered[i, :] = (1 - decay) * filtered[i, :] + decay * local[i, :]
        decay *= decay
    expect = p0.copy()
    expect[off["r"]:off["r"] + 3 * N] = filtered.reshape(-1)
    expect[3:6] = filtered[-1]
    assert np.array_equal(p1, expect)
