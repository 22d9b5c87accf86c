"""trajtrack_mpcndqn_rlboost_b200 -- B200-native batched NMPC planner + DQN companion.

Holds only what the hot path needs: csrc/ (sm_100a CUDA kernels + the C-ABI of
include/ttmpc.h) and the host-side mirror of the reference's solver interface.
"""
from .mpc_config import Configurator, num_params, param_offsets  # noqa: F401
from .solver import BatchSolver, Solver, OptimizerSolution, BatchSolution, EXIT_STATUS_NAMES, pinned_empty  # noqa: F401
from .planner import TrajectoryGenerator, InterfaceMpc  # noqa: F401
from .motion_model import unicycle_model  # noqa: F401
from .fleet import FleetPlanner  # noqa: F401
from . import scenes, dqn, geometry, fleet  # noqa: F401

__all__ = ["Configurator", "BatchSolver", "Solver", "TrajectoryGenerator", "InterfaceMpc",
           "FleetPlanner", "unicycle_model", "pinned_empty", "scenes", "dqn", "geometry", "fleet"]
