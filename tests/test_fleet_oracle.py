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


# ------------------------------------------------------------------ HintSwitcher
class _Poly:
    """Minimal stand-in for shapely's Polygon (contains: strict interior; distance: 0 inside)."""
    def __init__(self, pts):
        self.p = [tuple(map(float, q)) for q in pts]

    def contains(self, pt):
        x, y = pt
        inside = False
        n = len(self.p)
        for i in range(n):
            (xi, yi), (xj, yj) = self.p[i], self.p[i - 1]
            if (yi > y) != (yj > y) and x < (xj - xi) * (y - yi) / (yj - yi) + xi:
                inside = not inside
        return inside

    def distance(self, pt):
        if self.contains(pt):
            return 0.0
        x, y = pt
        best = float("inf")
        n = len(self.p)
        for i in range(n):
            (ax, ay), (bx, by) = self.p[i - 1], self.p[i]
            dx, dy = bx - ax, by - ay
            l2 = dx * dx + dy * dy
            tt = 0.0 if l2 == 0 else min(1.0, max(0.0, ((x - ax) * dx + (y - ay) * dy) / l2))
            best = min(best, ((ax + tt * dx - x) ** 2 + (ay + tt * dy - y) ** 2) ** 0.5)
        return best


class _HintSwitcher:
    """main_pre.py:27-52, transcribed statement for statement (Polygon -> _Poly)."""
    def __init__(self, max_switch_distance, min_detach_distance, min_detach_steps=5):
        self.switch_distance = max_switch_distance
        self.detach_distance = min_detach_distance
        self.detach_steps = min_detach_steps
        self.detach_cnt = 0
        self.switch_on = False

    def switch(self, current_position, original_traj, new_traj, obstacle_list):
        cnt_flag = False
        for old_pos, new_pos in zip(original_traj, new_traj):
            for obstacle in obstacle_list:
                shapely_obstacle = _Poly(obstacle)
                dist = shapely_obstacle.distance(tuple(current_position))
                if shapely_obstacle.contains(tuple(old_pos[:2])):
                    if (dist < self.switch_distance) & (self.switch_on == False):  # noqa: E712
                        self.switch_on = True
                        return self.switch_on
                elif (dist > self.detach_distance) & (self.switch_on == True):  # noqa: E712
                    if self.detach_cnt > self.detach_steps:
                        self.switch_on = False
                        self.detach_cnt = 0
                    elif cnt_flag == False:  # noqa: E712
                        self.detach_cnt += 1
                        cnt_flag = True
        return self.switch_on


def test_hint_switch_state_machine():
    """The pack oracle's switch decision over a drive past an obstacle that sits on the reference:
    on when close, off again detach_steps + 2 steps after leaving it (main.py:129: HintSwitcher(10, 2, 10))."""
    mc = t.Configurator()
    cfg = mc.to_ttmpc()
    N = cfg.N_hor
    tuning, base = work_mode(mc, "work")
    ref = np.stack([np.arange(400) * 0.24, np.zeros(400), np.zeros(400)], axis=1)       # straight line along x
    block = [(20.0, -1.0), (23.0, -1.0), (23.0, 1.0), (20.0, 1.0)]                        # sits on the path
    far = [(60.0, 30.0), (62.0, 30.0), (62.0, 32.0), (60.0, 32.0)]
    polys = [block, far]
    stc = np.zeros(cfg.Nstcobs * cfg.nstcobs)
    fh = O.FleetHost(cfg, np.zeros((1, 3)), np.array([[90.0, 0.0, 0.0]]), ref[None], [len(ref)], stc, tuning, base,
                     mc.low_speed, dyn_cur=np.array([[[40.0, 0.5]]]), dyn_disp=np.zeros((1, 1, 2)))
    fh.hint, fh.use_hint = np.zeros((1, N, 2)), np.zeros(1, np.int32)
    fh.sw_state = np.zeros((1, 2), np.int32)
    fh.sw_poly_xy = np.array(polys, dtype=np.float64)[None]
    fh.sw_poly_nv = np.array([[4, 4]], np.int32)
    sw = _HintSwitcher(10, 2, 10)
    decisions = []
    for step in range(260):
        x = 0.24 * step
        fh.state[0, :2] = (x, 0.3)
        fh.idx_ref[0] = max(0, step - 1)
        O.fleet_pack(fh, True)
        idx = int(fh.idx_ref[0])
        original = [ref[min(idx + k, len(ref) - 1)] for k in range(N)]
        obstacles = polys + [[(40.0 - 1.6, 0.5 - 1.6), (40.0 + 1.6, 0.5 - 1.6), (40.0 + 1.6, 0.5 + 1.6), (40.0 - 1.6, 0.5 + 1.6)]]
        expect = sw.switch((x, 0.3), original, original, obstacles)
        assert bool(fh.use_hint[0]) == bool(expect), f"step {step}"
        assert int(fh.sw_state[0, 0]) == int(sw.switch_on) and int(fh.sw_state[0, 1]) == sw.detach_cnt
        decisions.append(int(expect))
    d = np.array(decisions)
    assert d[0] == 0 and d.max() == 1 and d[-1] == 0 and (np.diff(d) != 0).sum() >= 2   # off -> on -> ... -> off
