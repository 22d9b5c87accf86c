"""Batched, device-resident counterpart of ``InterfaceMpc`` for a fleet of independent robots.

One ``FleetPlanner.step()`` does, for every robot at once and without leaving the GPU, what the
reference's control loop does per robot per step (src/main.py:155-172, decision_mode 1):

    original_ref_traj, _ = traj_gen.get_local_ref_traj()     # interface_mpc.py:72-79
    traj_gen.update_dynamic_constraints(pred)                 # main.py:80-89 est_dyn_obs_positions
    action, pred_states, cost = traj_gen.get_action(ref)      # interface_mpc.py:80-92

through ``ttmpc_fleet_step_device`` (pack -> solve -> advance, include/ttmpc.h).  The global
reference trajectories are sampled once on the host with the same
``TrajectoryGenerator.get_global_ref_traj`` the single-robot mirror uses.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .geometry import polygon_halfspace_representation
from .mpc_config import Configurator
from .planner import TrajectoryGenerator

RUNNING, REACHED, FAILED = 0, 1, 2
DYN_OBS_SIZE = 0.8 + 0.8  # main.py:32


def work_mode(config: Configurator, mode: str = "work"):
    """(tuning_params, base_speed) of TrajectoryGenerator.set_work_mode (trajectory_generator.py:117-139)."""
    c = config
    if mode == "aligning":
        tuning = [0.0] * 10
        tuning[2] = 100
        return tuning, c.lin_vel_max * c.medium_speed
    tuning = [c.qpos, c.qvel, c.qtheta, c.lin_vel_penalty, c.ang_vel_penalty,
              c.qpN, c.qthetaN, c.qrpd, c.lin_acc_penalty, c.ang_acc_penalty]
    speed = {"safe": c.low_speed, "work": c.high_speed, "super": c.full_speed}
    if mode not in speed:
        raise ModuleNotFoundError(f"There is no mode called {mode}.")
    return tuning, c.lin_vel_max * speed[mode]


class FleetPlanner:
    """n robots, each with its own start, goal, reference path and (optionally) obstacles."""

    def __init__(self, config: Configurator, init_states: np.ndarray, goal_states: np.ndarray,
                 ref_paths: Sequence[Sequence], mode: str = "work", device: str = "cuda",
                 stc_weights: float = 1e3, dyn_weights: float = 1e3, **solver_overrides):
        import torch
        self.torch = torch
        self.config = config
        self.cfg = config.to_ttmpc(**solver_overrides)
        self.lib = _lib.load()
        self.n = n = len(ref_paths)
        self.N = config.N_hor
        self.np = self.lib.ttmpc_num_params(C.byref(self.cfg))
        self.tuning, self.base_speed = work_mode(config, mode)
        init_states = np.asarray(init_states, dtype=np.float64).reshape(n, 3)
        goal_states = np.asarray(goal_states, dtype=np.float64).reshape(n, 3)
        trajs = [TrajectoryGenerator.get_global_ref_traj(config.ts, path, init_states[i], self.base_speed)
                 for i, path in enumerate(ref_paths)]
        self.ref_stride = max(1, max(len(t) for t in trajs))
        ref = np.zeros((n, self.ref_stride, 3))
        for i, t in enumerate(trajs):
            ref[i, :len(t)] = t
        f64 = dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        self.state = torch.tensor(init_states, **f64)
        self.goal = torch.tensor(goal_states, **f64)
        self.last_u = torch.zeros(n, 2, **f64)
        self.idx_ref = torch.zeros(n, **i32)
        self.status = torch.zeros(n, **i32)
        self.ref_traj = torch.tensor(ref, **f64)
        self.ref_len = torch.tensor([len(t) for t in trajs], **i32)
        self.stc = torch.zeros(1, config.Nstcobs * config.nstcobs, **f64)
        self.stc_shared = 1
        self.other = None
        self.dyn_cur = self.dyn_last = self.dyn_disp = None
        self.n_dyn_live = 0
        self.stc_weight, self.dyn_weight = float(stc_weights), float(dyn_weights)
        self.p = torch.empty(n, self.np, **f64)
        self.u = torch.zeros(n, 2 * self.N, **f64)
        self.y = torch.zeros(n, 2 * self.N, **f64)
        self.cost = torch.zeros(n, **f64)
        self.exit_status = torch.zeros(n, **i32)
        self.inner = torch.zeros(n, **i32)
        self.pred_states = torch.zeros(n, self.N, 3, **f64)
        self.hint = None       # [n][N][2] DQN hint positions (rl_ref) and the per-robot switch
        self.use_hint = None
        self.steps = 0

    # ------------------------------------------------------------------ obstacles
    def update_static_constraints(self, obstacle_list, per_robot: bool = False):
        """obstacle_list: polygons shared by every robot, or (per_robot=True) one list per robot.
        Same row layout as InterfaceMpc.update_static_constraints (interface_mpc.py:62-65)."""
        cfg = self.config

        def rows(polys):
            r = [0.0] * cfg.Nstcobs * cfg.nstcobs
            for i, poly in enumerate(polys):
                b, a0, a1 = polygon_halfspace_representation(np.array(poly))
                r[i * cfg.nstcobs:(i + 1) * cfg.nstcobs] = (b + a0 + a1)
            return r
        data = [rows(pl) for pl in obstacle_list] if per_robot else [rows(obstacle_list)]
        self.stc = self.torch.tensor(np.array(data, dtype=np.float64), device=self.state.device)
        self.stc_shared = 0 if per_robot else 1

    def set_moving_obstacles(self, positions: np.ndarray, displacement_per_step: np.ndarray):
        """positions, displacement_per_step: [n][k][2].  Every step the planner sees
        est_dyn_obs_positions(last, current) (main.py:80-89) and the obstacles then move on."""
        t = self.torch
        pos = np.asarray(positions, dtype=np.float64).reshape(self.n, -1, 2)
        self.n_dyn_live = pos.shape[1]
        dev = self.state.device
        self.dyn_cur = t.tensor(pos, device=dev)
        self.dyn_last = t.tensor(pos, device=dev)
        self.dyn_disp = t.tensor(np.asarray(displacement_per_step, dtype=np.float64).reshape(self.n, -1, 2), device=dev)

    def set_hint(self, rl_ref, use_hint):
        """Hybrid mode (main.py:194-201): rl_ref [n][N][2] float64 CUDA tensor (dqn.rl_ref_device),
        use_hint [n] int32 -- where non-zero the hint replaces the reference positions."""
        self.hint, self.use_hint = rl_ref, use_hint

    def enable_hint_switch(self, polygons, per_robot: bool = False, max_switch_distance: float = 10.0,
                           min_detach_distance: float = 2.0, min_detach_steps: int = 10):
        """Let the pack kernel decide per step whether the DQN hint is used, as
        ``HintSwitcher(10, 2, 10).switch(...)`` does in main.py:129,200 (main_pre.py:27-52).
        polygons: the processed (inflated) static obstacle polygons, shared or one list per robot;
        moving obstacles enter as circle_to_rect squares of radius DYN_OBS_SIZE."""
        t = self.torch
        lists = polygons if per_robot else [polygons]
        max_poly = max(1, max(len(pl) for pl in lists))
        max_pv = max(3, max((len(poly) for pl in lists for poly in pl), default=3))
        xy = np.zeros((len(lists), max_poly, max_pv, 2))
        nv = np.zeros((len(lists), max_poly), np.int32)
        for i, pl in enumerate(lists):
            for j, poly in enumerate(pl):
                a = np.asarray(poly, np.float64).reshape(-1, 2)
                xy[i, j, :len(a)] = a
                nv[i, j] = len(a)
        dev = self.state.device
        self.sw_poly_xy, self.sw_poly_nv = t.tensor(xy, device=dev), t.tensor(nv, device=dev)
        self.sw_shape = (max_poly, max_pv, 0 if per_robot else 1)
        self.sw_params = (float(max_switch_distance), float(min_detach_distance), int(min_detach_steps))
        self.sw_state = t.zeros(self.n, 2, dtype=t.int32, device=dev)
        if self.use_hint is None:
            self.use_hint = t.zeros(self.n, dtype=t.int32, device=dev)

    # ------------------------------------------------------------------ the step
    def _fleet_struct(self) -> _lib.TtmpcFleet:
        f = _lib.TtmpcFleet()
        f.n, f.ref_stride = self.n, self.ref_stride
        f.state, f.goal, f.last_u = self.state.data_ptr(), self.goal.data_ptr(), self.last_u.data_ptr()
        f.idx_ref, f.status = self.idx_ref.data_ptr(), self.status.data_ptr()
        f.ref_traj, f.ref_len = self.ref_traj.data_ptr(), self.ref_len.data_ptr()
        f.stc, f.stc_shared = self.stc.data_ptr(), self.stc_shared
        f.n_dyn_live, f.action_steps = self.n_dyn_live, 1
        f.other = self.other.data_ptr() if self.other is not None else None
        f.dyn = None
        if self.dyn_cur is not None:
            f.dyn_cur, f.dyn_last, f.dyn_disp = self.dyn_cur.data_ptr(), self.dyn_last.data_ptr(), self.dyn_disp.data_ptr()
        f.dyn_size = DYN_OBS_SIZE
        for i, v in enumerate(self.tuning):
            f.tuning[i] = float(v)
        f.base_speed, f.low_speed = float(self.base_speed), float(self.config.low_speed)
        f.stc_weight, f.dyn_weight = self.stc_weight, self.dyn_weight
        if self.hint is not None and self.use_hint is not None:
            f.hint, f.use_hint = self.hint.data_ptr(), self.use_hint.data_ptr()
        if getattr(self, "sw_state", None) is not None:
            f.sw_state = self.sw_state.data_ptr()
            f.sw_poly_xy, f.sw_poly_nv = self.sw_poly_xy.data_ptr(), self.sw_poly_nv.data_ptr()
            f.sw_max_poly, f.sw_max_pv, f.sw_poly_shared = self.sw_shape
            f.sw_switch_distance, f.sw_detach_distance, f.sw_detach_steps = self.sw_params
            f.sw_dyn_radius = DYN_OBS_SIZE
        return f

    def _result_struct(self) -> _lib.TtmpcResult:
        r = _lib.TtmpcResult()
        r.u, r.cost, r.exit_status = self.u.data_ptr(), self.cost.data_ptr(), self.exit_status.data_ptr()
        r.inner_iters, r.y, r.pred_states = self.inner.data_ptr(), self.y.data_ptr(), self.pred_states.data_ptr()
        return r

    def step(self, keep_multipliers: bool = True, stream=None):
        """One control step of every robot, asynchronous on the current torch stream.

        keep_multipliers=True (default) is what the reference's control loop does: the PyO3 ``Solver``
        object of a robot owns its AlmCache, so the Lagrange multipliers one ``run()`` ends with are
        the ones the next ``run()`` starts from (``self.y`` lives on the device between steps and is
        passed in / out of the solve).  False starts every step from y = 0 (a fresh solver object).

        ``pred_states`` is the rollout of u* from the state the step STARTED in (p.s), i.e. one step
        earlier than ``TrajectoryGenerator.run_step``'s list, which re-applies u from the post-step
        state (trajectory_generator.py:296-301; ``planner.TrajectoryGenerator`` reproduces that list
        on the host)."""
        st = stream if stream is not None else self.torch.cuda.current_stream().cuda_stream
        f, r = self._fleet_struct(), self._result_struct()
        _lib.check(self.lib.ttmpc_fleet_step_device(C.byref(self.cfg), C.byref(f), self.p.data_ptr(),
                                                    1 if keep_multipliers else 0, C.byref(r), st),
                   "ttmpc_fleet_step_device")
        self.steps += 1

    def pack(self, stream=None):
        st = stream if stream is not None else self.torch.cuda.current_stream().cuda_stream
        f = self._fleet_struct()
        _lib.check(self.lib.ttmpc_fleet_pack_device(C.byref(self.cfg), C.byref(f), self.p.data_ptr(), st), "fleet pack")

    def advance(self, stream=None):
        st = stream if stream is not None else self.torch.cuda.current_stream().cuda_stream
        f = self._fleet_struct()
        _lib.check(self.lib.ttmpc_fleet_advance_device(C.byref(self.cfg), C.byref(f), self.u.data_ptr(),
                                                       self.exit_status.data_ptr(), st), "fleet advance")
