"""Host mirror of the solver object the reference loads from its OpEn build.

Reference interface (/root/reference/src/mpc_traj_tracker/trajectory_generator.py):
    l.27-29   class Solver: run(p, initial_guess, initial_lagrange_multipliers, initial_penalty)
    l.69-71   built_solver = __import__(optimizer_name); self.solver = built_solver.solver()
    l.284-289 solution = self.solver.run(parameters, initial_guess)
              solution.solution / .cost / .exit_status / .solve_time_ms

``Solver`` is that object for ONE scene (a batch of one on the GPU).
``BatchSolver`` is the same call over many independent scenes, either with host
(numpy) buffers or with device-resident torch tensors.  Both go through the
C-ABI in include/ttmpc.h; neither has a CPU path.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import TtmpcConfig, TtmpcResult
from .mpc_config import num_params

EXIT_STATUS_NAMES = ["Converged", "NotConvergedIterations", "NotConvergedOutOfTime",
                     "NotFiniteComputation"]


@dataclass
class OptimizerSolution:
    """Same attribute names as OpEn's python ``OptimizerSolution``."""
    exit_status: str
    num_outer_iterations: int
    num_inner_iterations: int
    last_problem_norm_fpr: float
    f1_infeasibility: float
    f2_norm: float
    solve_time_ms: float
    penalty: float
    solution: List[float]
    lagrange_multipliers: List[float]
    cost: float


@dataclass
class BatchSolution:
    """Struct-of-arrays result of a batched solve (numpy on host, torch on device)."""
    solution: object          # [n, nu*N]
    cost: object              # [n]
    exit_status: object       # [n] int32 codes (see EXIT_STATUS_NAMES)
    num_outer_iterations: object
    num_inner_iterations: object
    last_problem_norm_fpr: object
    f1_infeasibility: object
    f2_norm: object
    penalty: object
    lagrange_multipliers: object  # [n, 2N], as left in the solver cache
    pred_states: object           # [n, N, 3]
    evals: object                 # [n, 4]: cost evals, grad evals, solve ns, start ns
    solve_time_ms: float = 0.0

    def exit_status_names(self):
        codes = self.exit_status.tolist() if hasattr(self.exit_status, "tolist") else list(self.exit_status)
        return [EXIT_STATUS_NAMES[c] for c in codes]


def pinned_empty(shape, dtype=np.float64):
    """A numpy array in page-locked host memory (a view of a pinned torch tensor).  Parameter blocks
    built in such an array are copied to the GPU straight from it; ordinary (pageable) arrays
    are staged through the library's own pinned buffer first, which costs a host memcpy."""
    import torch
    tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
           np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}[np.dtype(dtype)]
    return torch.empty(tuple(np.atleast_1d(shape)), dtype=tdt).pin_memory().numpy()


def _ptr(a) -> Optional[int]:
    return None if a is None else a.ctypes.data


class BatchSolver:
    """Batched ``Solver.run`` over independent scenes."""

    def __init__(self, cfg: TtmpcConfig):
        self.lib = _lib.load()
        self.cfg = cfg
        self.np = num_params(cfg)
        assert self.np == self.lib.ttmpc_num_params(C.byref(cfg))
        self.nu = cfg.nu * cfg.N_hor
        self.n1 = 2 * cfg.N_hor
        # the CUDA device of the constructing thread; host calls from other threads (run_many) switch
        # their thread to it first -- a new thread starts on device 0
        dev = C.c_int(-1)
        self.device = dev.value if self.lib.ttmpc_get_device(C.byref(dev)) == 0 else -1

    # -------------------------------------------------------------- host buffers
    def run(self, p, initial_guess=None, initial_lagrange_multipliers=None,
            initial_penalty=None, want_pred_states: bool = True) -> BatchSolution:
        p = np.ascontiguousarray(p, dtype=np.float64)
        if p.ndim != 2 or p.shape[1] != self.np:
            raise ValueError(f"p must be [n, {self.np}], got {p.shape}")
        n = p.shape[0]
        N = self.cfg.N_hor
        u = np.zeros((n, self.nu)) if initial_guess is None else \
            np.array(initial_guess, dtype=np.float64, order="C").reshape(n, self.nu)
        y = np.zeros((n, self.n1)) if initial_lagrange_multipliers is None else \
            np.array(initial_lagrange_multipliers, dtype=np.float64, order="C").reshape(n, self.n1)
        c0 = None if initial_penalty is None else \
            np.ascontiguousarray(np.broadcast_to(np.asarray(initial_penalty, dtype=np.float64), (n,)))
        out = dict(
            cost=np.zeros(n), exit_status=np.zeros(n, np.int32), outer=np.zeros(n, np.int32),
            inner=np.zeros(n, np.int32), fpr=np.zeros(n), f1=np.zeros(n), f2=np.zeros(n),
            pen=np.zeros(n), pred=np.zeros((n, N, 3)) if want_pred_states else None,
            evals=np.zeros((n, 4), np.int64))
        res = TtmpcResult(u=_ptr(u), cost=_ptr(out["cost"]), exit_status=_ptr(out["exit_status"]),
                          outer_iters=_ptr(out["outer"]), inner_iters=_ptr(out["inner"]),
                          last_fpr=_ptr(out["fpr"]), f1_infeas=_ptr(out["f1"]), f2_norm=_ptr(out["f2"]),
                          penalty=_ptr(out["pen"]), y=_ptr(y), pred_states=_ptr(out["pred"]),
                          evals=_ptr(out["evals"]))
        t0 = time.perf_counter()
        if self.device >= 0:
            _lib.check(self.lib.ttmpc_set_device(self.device), "ttmpc_set_device")
        rc = self.lib.ttmpc_solve_batch_host(C.byref(self.cfg), n, _ptr(p),
                                             int(initial_guess is not None),
                                             int(initial_lagrange_multipliers is not None),
                                             _ptr(c0), C.byref(res))
        _lib.check(rc, "ttmpc_solve_batch_host")
        ms = (time.perf_counter() - t0) * 1e3
        return BatchSolution(u, out["cost"], out["exit_status"], out["outer"], out["inner"],
                             out["fpr"], out["f1"], out["f2"], out["pen"], y, out["pred"],
                             out["evals"], ms)

    def run_many(self, batches, depth: int = 6, **kw) -> List[BatchSolution]:
        """Solve several host batches with up to ``depth`` of them in flight.

        Each batch is one ``run`` call (same arguments, same results) issued from its own host
        thread; the library gives every in-flight call its own staging buffers, streams and
        scene queue, so the next batch's copies and bulk overlap the tail of the previous one
        (a 4096-scene batch ends with a handful of long scenes that keep 1 % of the GPU busy).
        Results are returned in the order of ``batches``.
        """
        from concurrent.futures import ThreadPoolExecutor
        batches = list(batches)
        if depth <= 1 or len(batches) <= 1:
            return [self.run(p, **kw) for p in batches]
        with ThreadPoolExecutor(max_workers=min(depth, 8)) as ex:
            return list(ex.map(lambda p: self.run(p, **kw), batches))

    # -------------------------------------------------------------- device tensors
    def alloc_device(self, n: int, device="cuda", want_pred_states: bool = True):
        """Allocate the output tensors of ``run_device`` once (reused across steps)."""
        import torch
        N = self.cfg.N_hor
        f64 = dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        return dict(
            u=torch.zeros(n, self.nu, **f64), y=torch.zeros(n, self.n1, **f64),
            cost=torch.zeros(n, **f64), exit_status=torch.zeros(n, **i32),
            outer=torch.zeros(n, **i32), inner=torch.zeros(n, **i32), fpr=torch.zeros(n, **f64),
            f1=torch.zeros(n, **f64), f2=torch.zeros(n, **f64), pen=torch.zeros(n, **f64),
            pred=torch.zeros(n, N, 3, **f64) if want_pred_states else None,
            evals=torch.zeros(n, 4, dtype=torch.int64, device=device))

    def run_device(self, p, bufs: dict, use_u0: bool = False, use_y0: bool = False, c0=None,
                   stream=None) -> BatchSolution:
        """p: torch.float64 CUDA tensor [n, np].  Asynchronous on the current torch stream.

        Solves issued on different streams (with different ``bufs``) overlap on the GPU: each
        stream has its own scene queue and scratch tables inside the library.

        With use_u0 / use_y0 the initial guess / multipliers are read from
        bufs['u'] / bufs['y'], which are overwritten with the results.
        """
        import torch
        if not (p.is_cuda and p.dtype == torch.float64 and p.is_contiguous()):
            raise ValueError("p must be a contiguous float64 CUDA tensor")
        n = p.shape[0]
        if p.shape[1] != self.np:
            raise ValueError(f"p must be [n, {self.np}]")
        dp = lambda t: None if t is None else t.data_ptr()
        res = TtmpcResult(u=dp(bufs["u"]), cost=dp(bufs["cost"]), exit_status=dp(bufs["exit_status"]),
                          outer_iters=dp(bufs["outer"]), inner_iters=dp(bufs["inner"]),
                          last_fpr=dp(bufs["fpr"]), f1_infeas=dp(bufs["f1"]), f2_norm=dp(bufs["f2"]),
                          penalty=dp(bufs["pen"]), y=dp(bufs["y"]), pred_states=dp(bufs["pred"]),
                          evals=dp(bufs["evals"]))
        st = torch.cuda.current_stream(p.device).cuda_stream if stream is None else stream
        # the library works on the calling thread's current device: make that p's device for the call
        prev = C.c_int(-1)
        self.lib.ttmpc_get_device(C.byref(prev))
        if prev.value != p.device.index:
            _lib.check(self.lib.ttmpc_set_device(p.device.index), "ttmpc_set_device")
        rc = self.lib.ttmpc_solve_batch_device(C.byref(self.cfg), n, p.data_ptr(), int(use_u0),
                                               int(use_y0), dp(c0), C.byref(res), C.c_void_p(st))
        if prev.value >= 0 and prev.value != p.device.index:
            self.lib.ttmpc_set_device(prev.value)
        _lib.check(rc, "ttmpc_solve_batch_device")
        return BatchSolution(bufs["u"], bufs["cost"], bufs["exit_status"], bufs["outer"],
                             bufs["inner"], bufs["fpr"], bufs["f1"], bufs["f2"], bufs["pen"],
                             bufs["y"], bufs["pred"], bufs["evals"])

    # -------------------------------------------------------------- problem functions
    def evaluate(self, p, u, c=None, y=None):
        """f, F1, F2, psi, grad psi for a host batch (parity testing)."""
        p = np.ascontiguousarray(p, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        n = p.shape[0]
        c = None if c is None else np.ascontiguousarray(np.broadcast_to(np.asarray(c, np.float64), (n,)))
        y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        f, psi = np.zeros(n), np.zeros(n)
        F1, grad = np.zeros((n, self.n1)), np.zeros((n, self.nu))
        F2 = np.zeros((n, max(self.cfg.Ndynobs, 1)))
        rc = self.lib.ttmpc_eval_batch_host(C.byref(self.cfg), n, _ptr(p), _ptr(u), _ptr(c), _ptr(y),
                                            _ptr(f), _ptr(F1), _ptr(F2), _ptr(psi), _ptr(grad))
        _lib.check(rc, "ttmpc_eval_batch_host")
        return dict(f=f, F1=F1, F2=F2[:, :self.cfg.Ndynobs], psi=psi, grad=grad)

    def read_stats(self, reset: bool = True):
        out = (C.c_ulonglong * 4)()
        _lib.check(self.lib.ttmpc_read_stats(out, int(reset)), "ttmpc_read_stats")
        return dict(cost_evals=out[0], grad_evals=out[1], dyn_bodies=out[2], panoc_iters=out[3])

    def launch_info(self, n: int):
        v = [C.c_int() for _ in range(5)]
        _lib.check(self.lib.ttmpc_launch_info(C.byref(self.cfg), n, *[C.byref(x) for x in v]),
                   "ttmpc_launch_info")
        return dict(grid=v[0].value, block=v[1].value, smem_bytes=v[2].value,
                    blocks_per_sm=v[3].value, sm_count=v[4].value)


class Solver:
    """Drop-in for the OpEn-generated ``<optimizer_name>.solver()`` object.

    Like the PyO3 object it keeps its ALM cache between calls: when
    ``initial_lagrange_multipliers`` is not given the multipliers left by the
    previous ``run`` are used (they start at zero).
    """

    def __init__(self, cfg: TtmpcConfig):
        self._batch = BatchSolver(cfg)
        self._y = np.zeros((1, self._batch.n1))

    def run(self, p, initial_guess=None, initial_lagrange_multipliers=None,
            initial_penalty=None) -> Optional[OptimizerSolution]:
        b = self._batch
        p = np.asarray(p, dtype=np.float64).reshape(1, -1)
        if p.shape[1] != b.np:
            print(f"1600 -> wrong number of parameters (p): expected {b.np}, got {p.shape[1]}")
            return None
        if initial_guess is not None and len(initial_guess) != b.nu:
            print("1600 -> initial guess has incompatible dimensions")
            return None
        if initial_lagrange_multipliers is not None:
            if len(initial_lagrange_multipliers) != b.n1:
                print("1700 -> wrong dimension of Lagrange multipliers")
                return None
            y0 = np.asarray(initial_lagrange_multipliers, dtype=np.float64).reshape(1, -1)
        else:
            y0 = self._y
        u0 = None if initial_guess is None else np.asarray(initial_guess, np.float64).reshape(1, -1)
        sol = b.run(p, u0, y0, initial_penalty, want_pred_states=False)
        code = int(sol.exit_status[0])
        if code == 3:  # SolverError::NotFiniteComputation -> the binding returns None
            return None
        self._y = sol.lagrange_multipliers.copy()
        return OptimizerSolution(
            exit_status=EXIT_STATUS_NAMES[code],
            num_outer_iterations=int(sol.num_outer_iterations[0]),
            num_inner_iterations=int(sol.num_inner_iterations[0]),
            last_problem_norm_fpr=float(sol.last_problem_norm_fpr[0]),
            f1_infeasibility=float(sol.f1_infeasibility[0]),
            f2_norm=float(sol.f2_norm[0]),
            solve_time_ms=sol.solve_time_ms,
            penalty=float(sol.penalty[0]),
            solution=sol.solution[0].tolist(),
            lagrange_multipliers=sol.lagrange_multipliers[0].tolist(),
            cost=float(sol.cost[0]))
