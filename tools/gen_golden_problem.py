#!/usr/bin/env python
"""Generate golden vectors for the NMPC problem functions FROM THE REFERENCE'S OWN CODE.

The reference defines its cost / constraints symbolically with casadi inside
``MpcModule.build`` (/root/reference/src/mpc_traj_tracker/mpc/mpc_generator.py:155-283)
and hands them to opengen.  Neither casadi nor opengen is installed here, so this
script installs two tiny stand-ins into ``sys.modules`` before importing the
reference module, and then RUNS THE REFERENCE'S build() UNMODIFIED:

  * ``casadi.casadi``  -> a numeric shim over torch.float64 tensors with casadi's
                          shape rules (column vectors, scalar broadcast, horizontal
                          repmat, linear indexing) and casadi SX's on-the-fly
                          simplification sq(sqrt(x)) -> x.  Because the symbols carry
                          values, build() computes f, F1, F2 numerically and
                          torch.autograd gives the exact gradient of the reference's
                          expression graph.
  * ``opengen``        -> records what build() passes to og.builder.Problem /
                          og.constraints.Rectangle / og.config.SolverConfiguration.

Outputs tests/golden/problem_default.npz:
  u [K,40], p [K,2658], f [K], grad_f [K,40], F1 [K,40], F2 [K,15],
  umin/umax, c_min/c_max (set C), and the solver settings build() chose.
Run here (needs /root/reference); the .npz is committed and travels to the GPU box.
"""
import os
import sys
import types

import numpy as np
import torch

REF_SRC = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

torch.set_default_dtype(torch.float64)


# --------------------------------------------------------------------------- casadi shim
class SX:
    """A dense casadi-like matrix over a 2-D torch tensor."""
    __array_priority__ = 1000

    def __init__(self, t, sqrt_of=None):
        if isinstance(t, SX):
            t = t.t
        t = torch.as_tensor(t, dtype=torch.float64)
        if t.dim() == 0:
            t = t.reshape(1, 1)
        elif t.dim() == 1:
            t = t.reshape(-1, 1)
        self.t = t
        self._sqrt_of = sqrt_of  # remembers x for sq(sqrt(x)) -> x

    # -- construction
    values = {}

    @staticmethod
    def sym(name, n, m=1):
        v = SX.values[name]
        assert v.shape[0] == n * m, (name, v.shape, n, m)
        return SX(v.reshape(n, m) if m > 1 else v.reshape(-1, 1))

    @staticmethod
    def ones(n, m=1):
        return SX(torch.ones(n, m))

    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def T(self):
        return SX(self.t.transpose(0, 1))

    # -- indexing: casadi linear (column-major) indexing for a single index
    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            r, c = idx
            out = self.t[r, c]
            if out.dim() == 1:
                out = out.reshape(-1, 1) if isinstance(r, slice) else out.reshape(1, -1)
            return SX(out)
        flat = self.t.transpose(0, 1).reshape(-1)  # column-major order
        out = flat[idx]
        if out.dim() == 0:
            return SX(out.reshape(1, 1))
        if self.t.shape[0] == 1 and self.t.shape[1] > 1:
            return SX(out.reshape(1, -1))  # row vectors stay rows
        return SX(out.reshape(-1, 1))

    # -- arithmetic with casadi's broadcasting
    @staticmethod
    def _bc(a, b):
        a = a if isinstance(a, SX) else SX(a)
        b = b if isinstance(b, SX) else SX(b)
        x, y = a.t, b.t
        if x.shape == y.shape or x.numel() == 1 or y.numel() == 1:
            return x, y
        if x.shape[0] == y.shape[0]:  # horizontal repmat rule
            if x.shape[1] % y.shape[1] == 0:
                return x, y.repeat(1, x.shape[1] // y.shape[1])
            if y.shape[1] % x.shape[1] == 0:
                return x.repeat(1, y.shape[1] // x.shape[1]), y
        raise ValueError(f"Dimension mismatch {tuple(x.shape)} vs {tuple(y.shape)}")

    def __add__(self, o): x, y = SX._bc(self, o); return SX(x + y)
    def __radd__(self, o): x, y = SX._bc(o, self); return SX(x + y)
    def __sub__(self, o): x, y = SX._bc(self, o); return SX(x - y)
    def __rsub__(self, o): x, y = SX._bc(o, self); return SX(x - y)
    def __mul__(self, o): x, y = SX._bc(self, o); return SX(x * y)
    def __rmul__(self, o): x, y = SX._bc(o, self); return SX(x * y)
    def __truediv__(self, o): x, y = SX._bc(self, o); return SX(x / y)
    def __rtruediv__(self, o): x, y = SX._bc(o, self); return SX(x / y)
    def __neg__(self): return SX(-self.t)

    def __pow__(self, e):
        if isinstance(e, (int, float)) and e == 2:
            if self._sqrt_of is not None:  # casadi: sq(sqrt(x)) simplifies to x
                return SX(self._sqrt_of)
            return SX(self.t * self.t)
        e = e.t if isinstance(e, SX) else e
        return SX(self.t ** e)


def _t(x):
    return x.t if isinstance(x, SX) else torch.as_tensor(x, dtype=torch.float64).reshape(1, 1) \
        if not torch.is_tensor(x) else x


cs = types.ModuleType("casadi.casadi")
cs.SX = SX
cs.DM = lambda x: SX(torch.as_tensor(np.asarray(x, dtype=np.float64)))
cs.vertcat = lambda *a: SX(torch.cat([_t(x).reshape(-1, _t(x).shape[1] if _t(x).dim() == 2 else 1) for x in a], dim=0))
cs.horzcat = lambda *a: SX(torch.cat([_t(x) for x in a], dim=1))
cs.vcat = lambda lst: cs.vertcat(*lst)
cs.hcat = lambda lst: cs.horzcat(*lst)
cs.transpose = lambda x: x.T
cs.sum1 = lambda x: SX(_t(x).sum(dim=0, keepdim=True))
cs.sum2 = lambda x: SX(_t(x).sum(dim=1, keepdim=True))
cs.dot = lambda a, b: SX((_t(a) * _t(b)).sum().reshape(1, 1))
cs.mtimes = lambda a, b: SX(_t(a) @ _t(b))
cs.cos = lambda x: SX(torch.cos(_t(x)))
cs.sin = lambda x: SX(torch.sin(_t(x)))
cs.acos = lambda x: SX(torch.acos(_t(x)))
cs.sign = lambda x: SX(torch.sign(_t(x)))
cs.norm_2 = lambda x: SX(torch.linalg.norm(_t(x)).reshape(1, 1))


def _sqrt(x):
    return SX(torch.sqrt(_t(x)), sqrt_of=_t(x))


def _fmax(a, b):
    x, y = SX._bc(a, b)
    return SX(torch.where(x >= y, x, y))  # casadi: d/dx fmax = (x >= y)


def _fmin(a, b):
    x, y = SX._bc(a, b)
    return SX(torch.where(x <= y, x, y))  # casadi: d/dx fmin = (x <= y)


def _mmin(x):
    flat = _t(x).reshape(-1)
    out = flat[0]
    for i in range(1, flat.numel()):
        out = torch.where(out <= flat[i], out, flat[i])
    return SX(out.reshape(1, 1))


cs.sqrt, cs.fmax, cs.fmin, cs.mmin = _sqrt, _fmax, _fmin, _mmin
casadi_pkg = types.ModuleType("casadi")
casadi_pkg.casadi = cs
sys.modules["casadi"] = casadi_pkg
sys.modules["casadi.casadi"] = cs

# --------------------------------------------------------------------------- opengen shim
CAPTURE = {}


class _Rect:
    def __init__(self, xmin, xmax):
        self.xmin, self.xmax = list(xmin), list(xmax)


class _Problem:
    def __init__(self, u, z, cost):
        CAPTURE.update(u=u, z=z, cost=cost)

    def with_constraints(self, b): CAPTURE["bounds"] = b; return self
    def with_aug_lagrangian_constraints(self, f1, c, y=None): CAPTURE.update(F1=f1, set_c=c); return self
    def with_penalty_constraints(self, f2): CAPTURE["F2"] = f2; return self


class _Chain:
    def __init__(self, tag): self._tag = tag; CAPTURE.setdefault(tag, {})

    def __getattr__(self, name):
        def rec(*a, **k):
            CAPTURE[self._tag][name] = a[0] if len(a) == 1 else a
            return self
        return rec


og = types.ModuleType("opengen.opengen")
og.constraints = types.SimpleNamespace(Rectangle=_Rect)
og.builder = types.SimpleNamespace(Problem=_Problem,
                                   OpEnOptimizerBuilder=lambda *a: _Chain("builder"))
og.config = types.SimpleNamespace(BuildConfiguration=lambda: _Chain("build_config"),
                                  OptimizerMeta=lambda: _Chain("meta"),
                                  SolverConfiguration=lambda: _Chain("solver_config"))
og_pkg = types.ModuleType("opengen")
og_pkg.opengen = og
for k in ("constraints", "builder", "config"):
    setattr(og_pkg, k, getattr(og, k))
sys.modules["opengen"] = og_pkg
sys.modules["opengen.opengen"] = og

# --------------------------------------------------------------------------- run the reference
sys.path.insert(0, REF_SRC)
from util.mpc_config import Configurator  # noqa: E402  (reference)
from pkg_motion_model import motion_model  # noqa: E402  (reference)
from mpc_traj_tracker.mpc.mpc_generator import MpcModule  # noqa: E402  (reference)

import contextlib  # noqa: E402
import io  # noqa: E402


def reference_functions(config, u_np, p_np, c=None, y_np=None):
    """Evaluate the reference's f, grad f, F1, F2 at (u, p) by running MpcModule.build.
    With (c, y) also psi = f + c/2 dist^2_C(F1 + y/max(c,1)) + c/2 |F2|^2 and its gradient,
    psi assembled as opengen's builder does (__construct_function_psi) from the
    reference's own f / F1 / F2 / set C."""
    N, ns, nu = config.N_hor, config.ns, config.nu
    u = torch.tensor(u_np, dtype=torch.float64, requires_grad=True)
    p = torch.tensor(p_np, dtype=torch.float64)
    sizes = [("s", 2 * ns + nu), ("q", config.nq), ("r", ns * N + N), ("c", ns * N * config.Nother),
             ("os", config.Nstcobs * config.nstcobs), ("od", config.Ndynobs * config.ndynobs * N),
             ("qstc", N), ("qdyn", N)]
    SX.values = {"u": u}
    o = 0
    for name, n in sizes:
        SX.values[name] = p[o:o + n]
        o += n
    assert o == p.numel()
    CAPTURE.clear()
    with contextlib.redirect_stdout(io.StringIO()):
        MpcModule(config).build(motion_model.unicycle_model)
    cost = CAPTURE["cost"].t.reshape(())
    (grad,) = torch.autograd.grad(cost, u, retain_graph=True)
    F1t, F2t = CAPTURE["F1"].t.reshape(-1), CAPTURE["F2"].t.reshape(-1)
    F1 = F1t.detach().numpy()
    F2 = F2t.detach().numpy()
    if c is None:
        return float(cost.detach()), grad.numpy(), F1, F2
    lo = torch.tensor(CAPTURE["set_c"].xmin); hi = torch.tensor(CAPTURE["set_c"].xmax)
    z = F1t + torch.tensor(y_np) / max(c, 1.0)
    e = z - torch.minimum(torch.maximum(z, lo), hi)
    psi = cost + c * (e * e).sum() / 2 + c * (F2t * F2t).sum() / 2
    (gpsi,) = torch.autograd.grad(psi, u)
    return float(cost.detach()), grad.numpy(), F1, F2, float(psi.detach()), gpsi.numpy()


def main():
    from trajtrack_mpcndqn_rlboost_b200 import Configurator as MyCfg, scenes
    config = Configurator(os.path.join("/root/reference/config", "mpc_default.yaml"), verbose=False)
    my = MyCfg(os.path.join("/root/reference/config", "mpc_default.yaml")).to_ttmpc()
    rng = np.random.default_rng(2024)
    K = 48
    # scenes with static + dynamic obstacles, some other robots near the path, and
    # controls that drive the rollout into obstacles so every term is exercised
    p = scenes.make_scenes(K, my, seed=7, n_static=5, n_dynamic=6, blocking_fraction=0.6)
    from trajtrack_mpcndqn_rlboost_b200.mpc_config import param_offsets
    off = param_offsets(my)
    N = my.N_hor
    for i in range(K):
        if i % 3 == 0:  # other robots: predicted states close to the reference path
            for j in range(rng.integers(1, 4)):
                ref = p[i, off["r"]:off["r"] + 3 * N].reshape(N, 3)
                blk = ref + rng.normal(0, 0.3, (N, 3))
                p[i, off["c"] + j * 3 * N: off["c"] + (j + 1) * 3 * N] = blk.reshape(-1)
        if i % 5 == 0:  # non-trivial terminal / control weights
            p[i, off["q"] + 3] = 0.5; p[i, off["q"] + 4] = 0.3
            p[i, off["q"] + 5] = 2.0; p[i, off["q"] + 6] = 1.5
        if i % 7 == 0:  # alpha != 1 on the dynamic obstacles
            od = p[i, off["od"]:off["od"] + my.Ndynobs * 6 * N].reshape(-1, 6)
            od[:, 5] *= rng.uniform(0.2, 2.0)
    u = np.zeros((K, 2 * N))
    u[:, 0::2] = rng.uniform(-0.5, 1.5, (K, N))
    u[:, 1::2] = rng.uniform(-0.5, 0.5, (K, N))
    u[0] = 0.0                       # the all-zero initial guess
    u[1, 0::2] = 1.2; u[1, 1::2] = 0  # straight at the reference speed
    u[2] *= 3.0                      # outside the box (PANOC evaluates such points)
    f = np.zeros(K); g = np.zeros((K, 2 * N)); F1 = np.zeros((K, 2 * N)); F2 = np.zeros((K, my.Ndynobs))
    psi = np.zeros(K); gpsi = np.zeros((K, 2 * N))
    c = rng.choice([10.0, 50.0, 250.0, 0.5], K)
    y = rng.normal(0.0, 2.0, (K, 2 * N))
    for i in range(K):
        f[i], g[i], F1[i], F2[i], psi[i], gpsi[i] = reference_functions(config, u[i], p[i], float(c[i]), y[i])
        print(i, f"f={f[i]:.6g} |g|={np.abs(g[i]).max():.4g} F2max={F2[i].max():.4g}")
    sc = CAPTURE["solver_config"]
    out = os.path.join(ROOT, "tests", "golden", "problem_default.npz")
    np.savez_compressed(
        out, u=u, p=p, f=f, grad_f=g, F1=F1, F2=F2, c=c, y=y, psi=psi, grad_psi=gpsi,
        umin=np.array(CAPTURE["bounds"].xmin), umax=np.array(CAPTURE["bounds"].xmax),
        c_min=np.array(CAPTURE["set_c"].xmin), c_max=np.array(CAPTURE["set_c"].xmax),
        initial_penalty=np.array(float(sc.get("with_initial_penalty", np.nan))),
        max_duration_micros=np.array(float(sc.get("with_max_duration_micros", np.nan))),
        solver_config_calls=np.array(sorted(sc.keys())))
    print("wrote", out, "F2>0 cases:", int((F2.max(1) > 0).sum()))


if __name__ == "__main__":
    main()
