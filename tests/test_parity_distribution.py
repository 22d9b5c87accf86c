"""Solver parity, quantified -- CPU part.

north_star asks for "same exit status, controls within 1e-4, cost within 1e-6 relative" against the
reference solve.  The reference solver (OpEn, Rust) cannot be built here (oracle/ttmpc_oracle.h:
PARITY UNPINNED for PANOC / ALM), so what can be measured is

  A. the kernel's operation order (WARP-order oracle == the GPU bit for bit, tests/test_gpu_parity.py)
     against the reference-order oracle (sequential sums, libm, unfused products like Rust), and
  B. the SELF-SENSITIVITY of the reference-order oracle: the same code on parameter vectors whose
     entries were moved by ONE ulp.

B is the floor ANY independent implementation of the algorithm sits on -- a different libm, FMA
contraction or summation order perturbs the arithmetic at least that much.  The tests below pin
both distributions (so a regression of either shows) and document the finding: PANOC's stopping
rule |gamma * fpr| < 1e-4 with gamma ~ 1e-3 leaves the iterate free within ~1e-2, and the 1e-4 /
1e-6 bars are not met even by the reference-order code against ITSELF under a 1-ulp perturbation.
The same table for the GPU is asserted in tests/test_gpu_parity.py and printed in BENCH.md.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import parity_report as PR  # noqa: E402


@pytest.fixture(scope="module")
def rows():
    return PR.run(["static4096", "mixed4096"], n=192, threads=os.cpu_count() or 1, use_gpu=False)


@pytest.mark.parametrize("name", ["static4096", "mixed4096"])
def test_kernel_order_against_reference_order(rows, name):
    k = rows[name]["kernel_vs_reference_order"]
    base = rows[name]["reference_order_vs_itself_1ulp"]
    # exit status: the two orders agree on almost every scene
    assert k["status_agree"] >= 0.90
    # where both converge the controls agree to the slack of the stopping rule (a few 1e-3 typical)
    assert k["both_converged"] >= 40
    assert k["conv_du_p50"] <= 5e-3
    # the kernel order is inside the envelope the reference order has against itself: its median
    # deviation is within a small factor of the 1-ulp self-sensitivity, not orders of magnitude off
    assert k["du_p50"] <= 10.0 * base["du_p50"] + 1e-3
    assert k["status_agree"] >= base["status_agree"] - 0.08


@pytest.mark.parametrize("name", ["static4096", "mixed4096"])
def test_one_ulp_self_sensitivity_already_breaks_the_stated_tolerance(rows, name):
    """The claim of DESIGN.md section 4, demonstrated instead of asserted: the reference-order code
    does not meet north_star's tolerance against itself when p moves by one ulp."""
    b = rows[name]["reference_order_vs_itself_1ulp"]
    assert b["status_agree"] >= 0.90            # the algorithm is stable in its exit status ...
    assert b["du_le_1e4"] <= 0.75               # ... but far from all scenes stay within 1e-4
    assert b["du_p99"] >= 1e-3                  # with a tail of scenes that move by much more
    assert b["cost_rel_le_1e6"] <= 0.75


def test_perturbation_is_one_ulp():
    p = np.array([[1.0, -2.5, 0.0, 1e-300, 3e8]])
    q = PR.perturb_one_ulp(p, 0)
    assert q[0, 2] == 0.0
    nz = p != 0
    assert np.all(q[nz] != p[nz])
    assert np.all(np.abs(q[nz] - p[nz]) <= np.spacing(np.abs(p[nz])) * 1.0000001)
