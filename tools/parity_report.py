#!/usr/bin/env python
"""Quantify solver parity.

For each BASELINE workload (static4096 / mixed4096 / dynamic8192, seeded samples) compare

  A. the kernel's operation order (the WARP-order oracle on the CPU here, or the GPU itself with
     --gpu; the two are bit-identical, tests/test_gpu_parity.py) against the REFERENCE-order oracle
     (sequential loops, libm, literal two-loop L-BFGS), and
  B. the self-sensitivity baseline: the reference-order oracle against ITSELF with every entry of
     the parameter vector p moved by one ulp (random direction).  This is what ANY other
     implementation of the same algorithm must expect -- a different compiler, libm, FMA
     contraction or summation order perturbs the arithmetic at least this much.

and print / save the distribution north_star's tolerance is stated on: exit status agreement,
|du|_inf per scene (quantiles and the fraction within 1e-4), relative cost difference (fraction
within 1e-6).  Output: a markdown table on stdout and profiles/r2_parity_distribution.json.

usage: python tools/parity_report.py [--n 512] [--gpu] [--threads 8] [--out profiles/r2_parity_distribution.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trajtrack_mpcndqn_rlboost_b200 as t  # noqa: E402
from tests import oracle_lib as O  # noqa: E402


def perturb_one_ulp(p, seed):
    """Every finite non-zero entry of p moved to a neighbouring double (random direction)."""
    rng = np.random.default_rng(seed)
    up = rng.random(p.shape) < 0.5
    q = np.where(up, np.nextafter(p, np.inf), np.nextafter(p, -np.inf))
    return np.where(p == 0.0, p, q)      # zero padding stays zero (slot inactive in both runs)


def compare(a, b):
    """Distribution of the differences between two solve_batch results (dicts)."""
    du = np.abs(a["u"] - b["u"]).max(axis=1)
    ca, cb = a["cost"], b["cost"]
    rel = np.abs(ca - cb) / np.maximum(np.abs(cb), 1e-300)
    same = a["exit_status"] == b["exit_status"]
    both = (a["exit_status"] == 0) & (b["exit_status"] == 0)
    out = dict(
        n=int(len(du)),
        status_agree=float(same.mean()),
        du_le_1e4=float((du <= 1e-4).mean()),
        du_p50=float(np.quantile(du, 0.5)), du_p90=float(np.quantile(du, 0.9)),
        du_p99=float(np.quantile(du, 0.99)), du_max=float(du.max()),
        cost_rel_le_1e6=float((rel <= 1e-6).mean()),
        cost_rel_p50=float(np.quantile(rel, 0.5)), cost_rel_p99=float(np.quantile(rel, 0.99)),
        cost_rel_max=float(rel.max()),
        both_converged=int(both.sum()),
    )
    if both.any():
        out.update(conv_du_le_1e4=float((du[both] <= 1e-4).mean()),
                   conv_du_p50=float(np.quantile(du[both], 0.5)),
                   conv_du_p99=float(np.quantile(du[both], 0.99)),
                   conv_du_max=float(du[both].max()),
                   conv_cost_rel_le_1e6=float((rel[both] <= 1e-6).mean()))
    return out


def workload_sample(name, n, seed=1000):
    w = dict(t.scenes.WORKLOADS[name])
    cfg = t.Configurator().to_ttmpc(**w["solver"])
    p = t.scenes.make_scenes(n, cfg, seed=seed, n_static=w["n_static"], n_dynamic=w["n_dynamic"],
                             blocking_fraction=w["blocking_fraction"])
    return cfg, p


def gpu_solve(cfg, p):
    s = t.BatchSolver(cfg)
    r = s.run(p)
    return dict(u=np.asarray(r.solution), cost=np.asarray(r.cost), exit_status=np.asarray(r.exit_status))


def run(names, n, threads, use_gpu, n_dyn=None):
    rows = {}
    for name in names:
        nn = n_dyn if (name.startswith("dynamic") and n_dyn) else n
        cfg, p = workload_sample(name, nn)
        ref = O.solve_batch(cfg, p, threads=threads, warp=False)
        krn = gpu_solve(cfg, p) if use_gpu else O.solve_batch(cfg, p, threads=threads, warp=True)
        ref1 = O.solve_batch(cfg, perturb_one_ulp(p, 7), threads=threads, warp=False)
        rows[name] = dict(kernel_vs_reference_order=compare(krn, ref),
                          reference_order_vs_itself_1ulp=compare(ref1, ref),
                          kernel_arm="gpu" if use_gpu else "warp-order oracle (bit-identical to the GPU)",
                          converged_reference_order=int((ref["exit_status"] == 0).sum()))
    return rows


def markdown(rows):
    out = ["| workload | comparison | exit status agrees | \\|Δu\\| ≤ 1e-4 | \\|Δu\\| p50 / p99 / max | cost within 1e-6 rel. | both converged: \\|Δu\\| ≤ 1e-4, p50 / p99 |",
           "|---|---|---:|---:|---|---:|---|"]
    for name, r in rows.items():
        for key, label in (("kernel_vs_reference_order", "kernel order vs reference order"),
                           ("reference_order_vs_itself_1ulp", "reference order vs itself, p ± 1 ulp")):
            c = r[key]
            conv = "—"
            if "conv_du_p50" in c:
                conv = f"{100 * c['conv_du_le_1e4']:.1f} %, {c['conv_du_p50']:.1e} / {c['conv_du_p99']:.1e} (n={c['both_converged']})"
            out.append(f"| `{name}` (n={c['n']}) | {label} | {100 * c['status_agree']:.1f} % | {100 * c['du_le_1e4']:.1f} % | "
                       f"{c['du_p50']:.1e} / {c['du_p99']:.1e} / {c['du_max']:.2f} | {100 * c['cost_rel_le_1e6']:.1f} % | {conv} |")
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--n-dynamic", type=int, default=128, help="sample size of the long-limit workload")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--gpu", action="store_true", help="kernel arm = the GPU (default: WARP-order oracle)")
    ap.add_argument("--workloads", default="static4096,mixed4096,dynamic8192")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_parity_distribution.json"))
    a = ap.parse_args()
    rows = run(a.workloads.split(","), a.n, a.threads, a.gpu, a.n_dynamic)
    print(markdown(rows))
    with open(a.out, "w") as f:
        json.dump(rows, f, indent=1)
    print("\nwrote", a.out)


if __name__ == "__main__":
    main()
