"""Fleet step on the GPU (through the C-ABI) against the CPU oracle: the packed parameter block
and K closed-loop steps (pack -> solve -> advance), bit for bit."""
import numpy as np
import pytest

import trajtrack_mpcndqn_rlboost_b200 as t
from trajtrack_mpcndqn_rlboost_b200.fleet import work_mode
from tests import oracle_lib as O

pytestmark = pytest.mark.gpu


def _make(n, seed, **solver):
    mc = t.Configurator()
    fl = t.scenes.make_fleet(n, seed=seed)
    fp = t.FleetPlanner(mc, fl["init"], fl["goal"], fl["paths"], mode="work", **solver)
    fp.update_static_constraints(fl["static_polys"], per_robot=True)
    fp.set_moving_obstacles(fl["moving_pos"], fl["moving_disp"])
    tuning, base = work_mode(mc, "work")
    fh = O.FleetHost(fp.cfg, fl["init"], fl["goal"], fp.ref_traj.cpu().numpy(), fp.ref_len.cpu().numpy(),
                     fp.stc.cpu().numpy(), tuning, base, mc.low_speed, dyn_cur=fl["moving_pos"], dyn_disp=fl["moving_disp"])
    return fp, fh


def test_pack_bit_exact():
    fp, fh = _make(96, seed=4)
    # put a few robots next to their goals (speed-reference branch, termination test)
    st = fp.state.cpu().numpy()
    g = fp.goal.cpu().numpy()
    st[:8, :2] = g[:8, :2] + np.linspace(0.0, 1.5, 8)[:, None] * 0.5
    st[8:12, :2] = g[8:12, :2] + 0.01
    fp.state.copy_(fp.torch.tensor(st)); fh.state[:] = st
    fp.pack(); fp.torch.cuda.synchronize()
    p_ref = O.fleet_pack(fh, use_libm=False)
    assert np.array_equal(fp.p.cpu().numpy(), p_ref)
    assert np.array_equal(fp.idx_ref.cpu().numpy(), fh.idx_ref)
    assert np.array_equal(fp.status.cpu().numpy(), fh.status)
    assert (fh.status[8:12] == 1).all() and (fh.status[12:] == 0).all()


def test_closed_loop_bit_exact():
    n, steps = 24, 4
    fp, fh = _make(n, seed=9, max_inner_iterations=60, max_outer_iterations=4)
    y = None  # the multipliers persist from step to step, like in the reference's Solver objects
    for k in range(steps):
        fp.step()
        fp.torch.cuda.synchronize()
        p = O.fleet_pack(fh, use_libm=False)
        assert np.array_equal(fp.p.cpu().numpy(), p), f"packed parameters differ at step {k}"
        ref = O.solve_batch(fp.cfg, p, y0=y, threads=8, warp=True)
        y = ref["y"]
        assert np.array_equal(fp.y.cpu().numpy(), y), f"multipliers differ at step {k}"
        assert np.array_equal(fp.u.cpu().numpy(), ref["u"]), f"controls differ at step {k}"
        assert np.array_equal(fp.exit_status.cpu().numpy(), ref["exit_status"])
        O.fleet_advance(fh, ref["u"], ref["exit_status"], use_libm=False)
        assert np.array_equal(fp.state.cpu().numpy(), fh.state), f"states differ at step {k}"
        assert np.array_equal(fp.last_u.cpu().numpy(), fh.last_u)
        assert np.array_equal(fp.idx_ref.cpu().numpy(), fh.idx_ref)
        assert np.array_equal(fp.status.cpu().numpy(), fh.status)
        assert np.array_equal(fp.dyn_cur.cpu().numpy(), fh.dyn_cur)
    moved = np.abs(fh.state[:, :2] - t.scenes.make_fleet(n, seed=9)["init"][:, :2]).max(axis=1)
    assert (moved > 0.05).mean() > 0.8          # the robots do drive


def test_fleet_matches_single_robot_mirror():
    """One robot through FleetPlanner and through the InterfaceMpc mirror (host numpy around the
    same solver): same packed vector on the first step, same state to rounding of libm vs device sincos."""
    mc = t.Configurator()
    fl = t.scenes.make_fleet(1, seed=21, n_moving=0)
    fp = t.FleetPlanner(mc, fl["init"], fl["goal"], fl["paths"], mode="work")
    fp.update_static_constraints(fl["static_polys"][0])
    one = t.InterfaceMpc(mc)
    one.initialization(fl["init"][0].copy(), fl["goal"][0].copy(), fl["paths"][0], mode="work")
    one.update_static_constraints(fl["static_polys"][0])
    ref_local, _ = one.get_local_ref_traj()
    params = one._traj_gen.assemble_parameters(one.stc_constraints, one.dyn_constraints, one.other_robot_states, ref_local)
    # set_work_mode('work') happens inside run_step; the fleet was built in that mode
    fp.pack(); fp.torch.cuda.synchronize()
    p_dev = fp.p.cpu().numpy()[0]
    one._traj_gen.set_work_mode("work")
    params = one._traj_gen.assemble_parameters(one.stc_constraints, one.dyn_constraints, one.other_robot_states, ref_local)
    np.testing.assert_allclose(p_dev, np.array(params, dtype=np.float64), rtol=4e-16, atol=0)
    action, pred, cost = one.get_action(ref_local, mode="work")
    fp.step(); fp.torch.cuda.synchronize()
    np.testing.assert_allclose(fp.state.cpu().numpy()[0], one.state, rtol=0, atol=1e-12)
    np.testing.assert_allclose(fp.last_u.cpu().numpy()[0], action, rtol=0, atol=1e-12)


def test_pack_with_dqn_hint_bit_exact():
    """Hybrid mode: hinted robots take rl_ref positions, headings from the original window --
    what InterfaceMpc.get_local_ref_traj(rl_ref) + ref_traj_filter(decay=1) produce."""
    import torch
    fp, fh = _make(40, seed=6)
    agent5 = torch.cat([fp.state, torch.rand(40, 2, dtype=torch.float64, device="cuda") - 0.3], dim=1).contiguous()
    act = torch.randint(0, 9, (40,), dtype=torch.int32, device="cuda")
    rl = t.dqn.rl_ref_device(agent5, act)
    use = (torch.arange(40, device="cuda") % 3 != 0).to(torch.int32)
    fp.set_hint(rl, use)
    fh.hint, fh.use_hint = rl.cpu().numpy(), use.cpu().numpy()
    fp.pack(); torch.cuda.synchronize()
    p_ref = O.fleet_pack(fh, use_libm=False)
    p_dev = fp.p.cpu().numpy()
    assert np.array_equal(p_dev, p_ref)
    # against the single-robot mirror for one hinted robot
    off = t.param_offsets(fp.cfg); N = fp.N
    e = 1
    refs = p_dev[e, off["r"]:off["r"] + 3 * N].reshape(N, 3)
    assert np.array_equal(refs[:, :2], rl.cpu().numpy()[e])
    fh.use_hint[:] = 0
    p_plain = O.fleet_pack(fh, use_libm=False)
    assert np.array_equal(refs[:, 2], p_plain[e, off["r"]:off["r"] + 3 * N].reshape(N, 3)[:, 2])
    assert np.array_equal(p_dev[e, 3:5], rl.cpu().numpy()[e, -1]) and p_dev[e, 5] == p_plain[e, 5]


def test_hint_switch_on_device_matches_oracle():
    """HintSwitcher evaluated inside the pack kernel: decision, its state and the packed vector
    equal the oracle's over several closed-loop steps."""
    import torch
    n = 48
    mc = t.Configurator()
    fl = t.scenes.make_fleet(n, seed=17)
    # move the first rectangle of every robot onto its path so that the switch has something to see
    for i in range(n):
        (ax, ay), (bx, by) = fl["paths"][i][0], fl["paths"][i][1]
        cx, cy = ax + 0.35 * (bx - ax), ay + 0.35 * (by - ay)
        fl["static_polys"][i][0] = [(cx - 0.6, cy - 0.6), (cx + 0.6, cy - 0.6), (cx + 0.6, cy + 0.6), (cx - 0.6, cy + 0.6)]
    fp = t.FleetPlanner(mc, fl["init"], fl["goal"], fl["paths"], mode="work", max_inner_iterations=40, max_outer_iterations=3)
    fp.update_static_constraints(fl["static_polys"], per_robot=True)
    fp.set_moving_obstacles(fl["moving_pos"], fl["moving_disp"])
    fp.enable_hint_switch(fl["static_polys"], per_robot=True)
    tuning, base = work_mode(mc, "work")
    fh = O.FleetHost(fp.cfg, fl["init"], fl["goal"], fp.ref_traj.cpu().numpy(), fp.ref_len.cpu().numpy(),
                     fp.stc.cpu().numpy(), tuning, base, mc.low_speed, dyn_cur=fl["moving_pos"], dyn_disp=fl["moving_disp"])
    fh.sw_state = np.zeros((n, 2), np.int32)
    fh.sw_poly_xy, fh.sw_poly_nv = fp.sw_poly_xy.cpu().numpy(), fp.sw_poly_nv.cpu().numpy()
    fh.use_hint = np.zeros(n, np.int32)
    g = torch.Generator(device="cpu").manual_seed(0)
    seen_on = 0
    y = None  # multipliers carry over from step to step (FleetPlanner.step default)
    for k in range(4):
        agent5 = torch.cat([fp.state, fp.last_u], 1).contiguous()
        act = torch.randint(0, 9, (n,), generator=g, dtype=torch.int32).cuda()
        rl = t.dqn.rl_ref_device(agent5, act)
        fp.set_hint(rl, fp.use_hint)
        fh.hint = rl.cpu().numpy()
        fp.step(); torch.cuda.synchronize()
        p = O.fleet_pack(fh, use_libm=False)
        assert np.array_equal(fp.use_hint.cpu().numpy(), fh.use_hint), f"switch decision differs at step {k}"
        assert np.array_equal(fp.sw_state.cpu().numpy(), fh.sw_state)
        assert np.array_equal(fp.p.cpu().numpy(), p)
        ref = O.solve_batch(fp.cfg, p, y0=y, threads=8, warp=True)
        y = ref["y"]
        O.fleet_advance(fh, ref["u"], ref["exit_status"], use_libm=False)
        assert np.array_equal(fp.state.cpu().numpy(), fh.state)
        seen_on += int(fh.use_hint.sum())
    assert seen_on > 0


def test_fleet_keeps_multipliers_like_per_robot_solver_objects():
    """FleetPlanner.step() must do what a loop of per-robot Solver objects does -- each
    robot's multipliers carry over from one control step to the next (the PyO3 Solver owns its
    AlmCache).  Six robots, four steps: the fleet's controls equal, bit for bit, those of six
    `Solver` mirrors fed the same packed vectors; a fleet stepped with keep_multipliers=False
    (fresh solver every step) differs from step 2 on for at least one robot."""
    n, steps = 6, 4
    fp, _ = _make(n, seed=31, max_inner_iterations=80, max_outer_iterations=5)
    cold, _ = _make(n, seed=31, max_inner_iterations=80, max_outer_iterations=5)
    solvers = [t.Solver(fp.cfg) for _ in range(n)]
    differs = False
    for k in range(steps):
        fp.pack(); fp.torch.cuda.synchronize()
        p = fp.p.cpu().numpy().copy()
        fp.step(); fp.torch.cuda.synchronize()
        u = fp.u.cpu().numpy()
        for i, sv in enumerate(solvers):
            sol = sv.run(p[i])
            assert sol is not None
            assert np.array_equal(np.asarray(sol.solution), u[i]), f"robot {i} differs at step {k}"
        cold.step(keep_multipliers=False); cold.torch.cuda.synchronize()
        if k >= 1 and not np.array_equal(cold.u.cpu().numpy(), u):
            differs = True
    assert differs, "keeping the multipliers made no difference: the test scenes never activate the ALM rows"
