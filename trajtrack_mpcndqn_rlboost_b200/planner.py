"""Host-side planner that drives the batched NMPC kernels through the reference's
own interface.

Mirrors, name for name, the two classes the rest of the reference talks to:

  TrajectoryGenerator  /root/reference/src/mpc_traj_tracker/trajectory_generator.py:32-333
  InterfaceMpc         /root/reference/src/interface_mpc.py:16-92

Only the solver behind ``run_solver`` is different: instead of importing the
OpEn-generated Rust extension (trajectory_generator.py:62-76) it owns a
``Solver`` that calls the CUDA library.  Everything else keeps the reference's
argument meaning, return values and error behaviour, so the reference's
``Simulator`` / ``main.py`` loops can hold one of these unchanged.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import motion_model as _mm
from .geometry import polygon_halfspace_representation
from .mpc_config import Configurator
from .solver import Solver

Number = Union[int, float]


def _as_xy_list(path) -> List[Tuple[float, float]]:
    """Accept PathNodeList-like objects, tuples or arrays; keep (x, y)."""
    out = []
    for node in path:
        if hasattr(node, "x") and hasattr(node, "y"):
            out.append((float(node.x), float(node.y)))
        else:
            out.append((float(node[0]), float(node[1])))
    return out


class TrajectoryGenerator:
    """Generate a smooth trajectory from the reference path and the obstacles (one robot)."""

    def __init__(self, config: Configurator, use_tcp: bool = False, verbose: bool = False,
                 **solver_overrides):
        if use_tcp:
            # trajectory_generator.py:72-75 starts OpEn's TCP server; there is no such
            # process here, the solver lives in this process on the GPU.
            raise NotImplementedError("use_tcp=True is not supported by the CUDA solver")
        self.__prtname = '[Traj]'
        self.vb = verbose
        self.config = config
        self.use_tcp = False
        self.ts, self.ns, self.nu, self.N_hor = config.ts, config.ns, config.nu, config.N_hor
        self.set_work_mode(mode='safe')
        self.set_obstacle_weights(stc_weights=1e3, dyn_weights=1e3)
        self.solver = Solver(config.to_ttmpc(**solver_overrides))
        self.motion_model: Callable = _mm.unicycle_model

    # ------------------------------------------------------------------ set-up
    def load_robot_dynamics(self, motion_model: Callable) -> None:
        """motion_model: s' = f(s, a, ts) used to roll the returned controls out on the host."""
        self.motion_model = motion_model

    def load_init_state(self, current_state: np.ndarray, goal_state: np.ndarray):
        if not isinstance(current_state, np.ndarray) or not isinstance(goal_state, np.ndarray):
            raise TypeError('State and action should be numpy.ndarry, '
                            f'got {type(current_state)}/{type(goal_state)}.')
        self.state = current_state
        self.final_goal = goal_state
        self.past_states: list = []
        self.past_actions: list = []
        self.cost_timelist: list = []
        self.solver_time_timelist: list = []
        self.idx_ref = 0

    def set_obstacle_weights(self, stc_weights: Union[list, Number], dyn_weights: Union[list, Number]):
        def expand(w):
            if isinstance(w, list):
                return w
            if isinstance(w, (float, int)):
                return [w] * self.N_hor
            raise TypeError(f'Unsupported datatype for obstacle weights, got {type(w)}.')
        self.stc_weights = expand(stc_weights)
        self.dyn_weights = expand(dyn_weights)

    def set_work_mode(self, mode: str = 'safe'):
        """base_speed and the 10 tuning parameters q of the cost (trajectory_generator.py:115-140)."""
        cfg = self.config
        if mode == 'aligning':
            self.base_speed = cfg.lin_vel_max * cfg.medium_speed
            self.tuning_params = [0.0] * 10
            self.tuning_params[2] = 100
            return
        speed_of = {'safe': cfg.low_speed, 'work': cfg.high_speed, 'super': cfg.full_speed}
        if mode not in speed_of:
            raise ModuleNotFoundError(f'There is no mode called {mode}.')
        self.tuning_params = [cfg.qpos, cfg.qvel, cfg.qtheta, cfg.lin_vel_penalty, cfg.ang_vel_penalty,
                              cfg.qpN, cfg.qthetaN, cfg.qrpd, cfg.lin_acc_penalty, cfg.ang_acc_penalty]
        self.base_speed = cfg.lin_vel_max * speed_of[mode]

    def set_current_state(self, current_state: np.ndarray):
        if not isinstance(current_state, np.ndarray):
            raise TypeError(f'State should be numpy.ndarry, got {type(current_state)}.')
        self.state = current_state

    def set_ref_trajectory(self, ref_path):
        self.idx_ref = 0
        self.ref_traj = self.get_global_ref_traj(self.ts, ref_path, self.state, self.base_speed)

    def check_termination_condition(self, state: np.ndarray, action: np.ndarray,
                                    final_goal: np.ndarray) -> bool:
        close = np.allclose(state[:2], final_goal[:2], atol=0.05, rtol=0)
        terminated = bool(close and abs(action[0]) < 0.05)
        if terminated:
            print(f"{self.__prtname} MPC solution found.")
        return terminated

    # ------------------------------------------------------------------ reference trajectory
    @staticmethod
    def get_global_ref_traj(ts: float, ref_path, state: Sequence[float], speed: float) -> np.ndarray:
        """Sample the polyline ``ref_path`` at constant ``speed`` every ``ts`` seconds,
        starting from ``state``; returns rows (x, y, heading).

        Follows trajectory_generator.py:160-201 step for step, including its
        corner behaviour (after landing on a node the walker immediately spends
        another full ``ts`` towards the next node within the same sample).
        """
        nodes = _as_xy_list(ref_path)
        px, py = float(state[0]), float(state[1])
        tx, ty = nodes[0]
        k = 0
        rows = []
        ux = uy = 0.0
        alive = True
        while alive:
            while True:
                gap = math.hypot(tx - px, ty - py)
                if gap < 1e-9:
                    k += 1
                    tx, ty = nodes[k]
                    break
                ux, uy = (tx - px) / gap, (ty - py) / gap
                eta = gap / speed
                if eta > ts:
                    px, py = px + ux * speed * ts, py + uy * speed * ts
                    break
                px, py = px + ux * speed * eta, py + uy * speed * eta
                k += 1
                if k > len(nodes) - 1:
                    alive = False
                    break
                tx, ty = nodes[k]
            if not gap < 1e-9:
                rows.append((px, py, math.atan2(uy, ux)))
        return np.array(rows, dtype=np.float64).reshape(-1, 3)

    @staticmethod
    def get_local_ref_traj(idx_ref: int, ref_traj_global, state: Sequence[float],
                           action_steps: int = 1, horizon: int = 20) -> Tuple[np.ndarray, int]:
        """The ``horizon`` reference states starting at the point of the global trajectory
        closest to ``state``, searched in [idx_ref - action_steps, idx_ref + 5*action_steps);
        short tails are padded with the last state (trajectory_generator.py:203-230)."""
        g = ref_traj_global.numpy() if hasattr(ref_traj_global, "numpy") else np.asarray(ref_traj_global)
        g = np.asarray(g, dtype=np.float64).reshape(-1, 3)
        lo = max(0, idx_ref - 1 * action_steps)
        hi = min(len(g), idx_ref + 5 * action_steps)
        d = [math.hypot(state[0] - g[i, 0], state[1] - g[i, 1]) for i in range(lo, hi)]
        idx_next = d.index(min(d)) + lo
        tail = g[idx_next:idx_next + horizon]
        if idx_next + horizon >= len(g):
            pad = horizon - (len(g) - idx_next)
            tail = np.concatenate([g[idx_next:], np.repeat(g[-1:], pad, axis=0)], axis=0)
        return np.array(tail, dtype=np.float64), idx_next

    # ------------------------------------------------------------------ one MPC step
    def assemble_parameters(self, stc_constraints: list, dyn_constraints: list,
                            other_robot_states: list, current_ref_traj: np.ndarray) -> list:
        """The packed parameter vector, block order of trajectory_generator.py:251-254."""
        finish_state = current_ref_traj[-1, :]
        current_refs = current_ref_traj.reshape(-1).tolist()
        dist_to_goal = math.hypot(self.state[0] - self.final_goal[0], self.state[1] - self.final_goal[1])
        if dist_to_goal >= self.base_speed * self.N_hor * self.ts:
            speed_ref_list = [self.base_speed] * self.N_hor
        else:
            speed_ref = max(dist_to_goal / self.N_hor / self.ts, self.config.low_speed)
            speed_ref_list = [speed_ref] * self.N_hor
        last_u = self.past_actions[-1] if len(self.past_actions) else np.zeros(self.nu)
        return (list(self.state) + list(finish_state) + list(last_u) + list(self.tuning_params)
                + current_refs + speed_ref_list + list(other_robot_states)
                + list(stc_constraints) + list(dyn_constraints)
                + list(self.stc_weights) + list(self.dyn_weights))

    def run_step(self, stc_constraints: list, dyn_constraints: list, other_robot_states: list,
                 current_ref_traj: np.ndarray, mode: str = 'safe',
                 initial_guess: Optional[np.ndarray] = None):
        """Returns (actions, pred_states, cost) like trajectory_generator.py:233-274."""
        self.set_work_mode(mode)
        params = self.assemble_parameters(stc_constraints, dyn_constraints, other_robot_states,
                                          current_ref_traj)
        try:
            taken_states, pred_states, actions, cost, solver_time, exit_status = \
                self.run_solver(params, self.state, self.config.action_steps, initial_guess)
        except RuntimeError as err:
            raise RuntimeError(f"Fatal: Cannot run solver. {err}.")
        self.past_states.append(self.state)
        self.past_states += taken_states[:-1]
        self.past_actions += actions
        self.state = taken_states[-1]
        self.cost_timelist.append(cost)
        self.solver_time_timelist.append(solver_time)
        if exit_status in self.config.bad_exit_codes and self.vb:
            print(f"{self.__prtname} Bad converge status: {exit_status}")
        return actions, pred_states, cost

    def run_solver(self, parameters: list, state: np.ndarray, take_steps: int = 1,
                   initial_guess: Optional[np.ndarray] = None):
        """Solve, then roll the controls out on the host with ``self.motion_model``
        (trajectory_generator.py:276-307)."""
        solution = self.solver.run(parameters, initial_guess)
        if solution is None:
            # the PyO3 binding returns None on a solver error; the reference then fails on
            # ``solution.solution`` -- surface it as the RuntimeError run_step expects
            raise RuntimeError("solver returned no solution (NotFiniteComputation)")
        u = solution.solution
        nu = self.nu
        taken_states: List[np.ndarray] = []
        for i in range(take_steps):
            taken_states.append(self.motion_model(state, np.array(u[i * nu:(i + 1) * nu]), self.ts))
        pred_states: List[np.ndarray] = []
        cursor = taken_states[-1]
        for i in range(len(u) // nu):
            cursor = self.motion_model(cursor, np.array(u[i * nu:i * nu + 2]), self.ts)
            pred_states.append(cursor)
        actions = [np.array(a) for a in np.array(u[:nu * take_steps]).reshape(take_steps, nu).tolist()]
        return taken_states, pred_states, actions, solution.cost, solution.solve_time_ms, \
            solution.exit_status


class InterfaceMpc:
    """The object ``main.py`` drives (interface_mpc.py:16-92)."""

    def __init__(self, config: Configurator, use_tcp: bool = False, verbose: bool = False,
                 motion_model: Optional[Callable] = None, **solver_overrides):
        self._traj_gen = TrajectoryGenerator(config, use_tcp, verbose=verbose, **solver_overrides)
        self._traj_gen.load_robot_dynamics(motion_model if motion_model is not None
                                           else _mm.unicycle_model)
        self._last_action = np.array([0.0, 0.0])
        cfg = self.config
        self.stc_constraints = [0.0] * cfg.Nstcobs * cfg.nstcobs
        self.dyn_constraints = [0.0] * cfg.Ndynobs * cfg.ndynobs * cfg.N_hor
        self.other_robot_states = [0] * cfg.ns * cfg.N_hor * cfg.Nother

    config = property(lambda self: self._traj_gen.config)
    state = property(lambda self: self._traj_gen.state)
    last_action = property(lambda self: self._last_action)
    goal = property(lambda self: self._traj_gen.final_goal)
    ref_path = property(lambda self: self._ref_path)
    ref_traj = property(lambda self: self._traj_gen.ref_traj)

    def set_current_state(self, state: np.ndarray):
        self._traj_gen.set_current_state(state)

    def initialization(self, init_state: np.ndarray, goal_state: np.ndarray,
                       ref_path_list: List[tuple], mode: str = 'work'):
        self._ref_path = [tuple(x) for x in ref_path_list]
        self._traj_gen.load_init_state(init_state, goal_state)
        self._traj_gen.set_work_mode(mode)
        self._traj_gen.set_ref_trajectory(self._ref_path)

    def update_static_constraints(self, obstacle_list):
        n = self.config.nstcobs
        for i, map_obstacle in enumerate(obstacle_list):
            b, a0, a1 = polygon_halfspace_representation(np.array(map_obstacle))
            self.stc_constraints[i * n:(i + 1) * n] = (b + a0 + a1)

    def update_dynamic_constraints(self, full_dyn_obstacle_list):
        per_obs = self.config.ndynobs * self.config.N_hor
        for i, dyn_obstacle in enumerate(full_dyn_obstacle_list):
            self.dyn_constraints[i * per_obs:(i + 1) * per_obs] = list(itertools.chain(*dyn_obstacle))

    def update_other_robot_states(self, other_robot_states):
        self.other_robot_states = other_robot_states

    def get_local_ref_traj(self, local_ref_traj: Optional[np.ndarray] = None):
        tg = self._traj_gen
        original, idx = tg.get_local_ref_traj(tg.idx_ref, self.ref_traj, self.state,
                                              action_steps=self.config.action_steps,
                                              horizon=self.config.N_hor)
        tg.idx_ref = idx
        if local_ref_traj is not None and local_ref_traj.shape[1] == 2:
            local_ref_traj = np.concatenate((local_ref_traj, original[:, [2]]), axis=1)
        return original, local_ref_traj

    def get_action(self, current_ref_traj: np.ndarray, mode='work',
                   initial_guess: Optional[np.ndarray] = None):
        tg = self._traj_gen
        if tg.check_termination_condition(self.state, self._last_action, self.goal):
            return None
        actions, pred_states, cost = tg.run_step(self.stc_constraints, self.dyn_constraints,
                                                 self.other_robot_states, current_ref_traj, mode,
                                                 initial_guess)
        self._last_action = actions[0]
        return actions[0], pred_states, cost
