"""Configuration loader mirroring the reference's ``util.mpc_config.Configurator``
(/root/reference/src/util/mpc_config.py:8-20): every key of the YAML file becomes
an attribute.  ``to_ttmpc`` maps it onto the C-ABI ``ttmpc_config``; the solver
fields are opengen 0.7.1 ``SolverConfiguration`` defaults with
``initial_penalty = 10`` as the reference sets at mpc_generator.py:268-276.
"""
from __future__ import annotations

import yaml

from ._lib import TtmpcConfig

# opengen.config.SolverConfiguration defaults + mpc_generator.py:269 (.with_initial_penalty(10))
SOLVER_DEFAULTS = dict(
    tolerance=1e-4, initial_tolerance=1e-4, delta_tolerance=1e-4, initial_penalty=10.0,
    penalty_update_factor=5.0, inner_tolerance_update_factor=0.1,
    sufficient_decrease_coeff=0.1, lbfgs_memory=10, max_inner_iterations=500,
    max_outer_iterations=10,
    max_duration_ms=5000,  # MAX_SOVLER_TIME = 5_000_000 us (mpc_generator.py:22, 270); 0 = no limit
)

# config/mpc_default.yaml of the reference, so that synthetic benches do not need the file
MPC_DEFAULT = dict(
    vehicle_width=0.5, vehicle_margin=0.1, social_margin=0.2, lin_vel_min=-0.5, lin_vel_max=1.5,
    lin_acc_min=-1, lin_acc_max=1, ang_vel_max=0.5, ang_acc_max=3,
    full_speed=1.0, high_speed=0.8, medium_speed=0.5, low_speed=0.2,
    ts=0.2, N_hor=20, action_steps=1,
    lin_vel_penalty=0, lin_acc_penalty=10.0, ang_vel_penalty=0, ang_acc_penalty=20.0,
    qrpd=100.0, qpos=0.0, qvel=10.0, qtheta=0.0, qpN=0.0, qthetaN=0.0,
    nu=2, ns=3, nq=10, Nother=10, Nstcobs=10, nstcobs=12, Ndynobs=15, ndynobs=6,
    build_type='release', build_directory='mpc_solver',
    bad_exit_codes=["NotConvergedIterations", "NotConvergedOutOfTime"],
    optimizer_name='navi_default',
)


class Configurator:
    def __init__(self, yaml_fp=None, verbose: bool = False, **overrides):
        self.__prtname = '[MPC-CFG]'
        if yaml_fp is None:
            loaded = dict(MPC_DEFAULT)
        else:
            if verbose:
                print(f'{self.__prtname} Loading configuration from "{yaml_fp}".')
            with open(yaml_fp, 'r') as stream:
                loaded = yaml.safe_load(stream)
        loaded.update(overrides)
        for key in loaded:
            setattr(self, key, loaded[key])
        if verbose:
            print(f'{self.__prtname} Configuration done.')

    def to_ttmpc(self, **solver_overrides) -> TtmpcConfig:
        c = TtmpcConfig()
        for f in ("N_hor", "nu", "ns", "nq", "Nother", "Nstcobs", "nstcobs", "Ndynobs", "ndynobs"):
            setattr(c, f, int(getattr(self, f)))
        for f in ("ts", "vehicle_width", "social_margin", "lin_vel_min", "lin_vel_max",
                  "ang_vel_max", "lin_acc_min", "lin_acc_max", "ang_acc_max"):
            setattr(c, f, float(getattr(self, f)))
        s = dict(SOLVER_DEFAULTS)
        s.update({k: getattr(self, k) for k in SOLVER_DEFAULTS if hasattr(self, k)})
        s.update(solver_overrides)
        for k, v in s.items():
            setattr(c, k, type(getattr(c, k))(v))
        return c


def num_params(cfg: TtmpcConfig) -> int:
    N = cfg.N_hor
    return (2 * cfg.ns + cfg.nu + cfg.nq + cfg.ns * N + N + cfg.ns * N * cfg.Nother
            + cfg.Nstcobs * cfg.nstcobs + cfg.Ndynobs * cfg.ndynobs * N + 2 * N)


def param_offsets(cfg: TtmpcConfig) -> dict:
    """Offsets of the blocks of the packed parameter vector (mpc_generator.py:175-184)."""
    N = cfg.N_hor
    o = {"s": 0}
    o["q"] = 2 * cfg.ns + cfg.nu
    o["r"] = o["q"] + cfg.nq
    o["vref"] = o["r"] + cfg.ns * N
    o["c"] = o["vref"] + N
    o["os"] = o["c"] + cfg.ns * N * cfg.Nother
    o["od"] = o["os"] + cfg.Nstcobs * cfg.nstcobs
    o["qstc"] = o["od"] + cfg.Ndynobs * cfg.ndynobs * N
    o["qdyn"] = o["qstc"] + N
    o["np"] = o["qdyn"] + N
    return o
