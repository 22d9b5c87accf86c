"""DQN side of the hybrid loop: the C oracle against goldens produced by the reference's own
MobileRobot / observation components (tools/gen_golden_dqn_loop.py -> tests/golden/dqn_loop.npz)."""
import os

import numpy as np

from tests import oracle_lib as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "dqn_loop.npz"))


def test_internal_observation_matches_reference_components():
    worst = 0.0
    for k in range(len(G["agent"])):
        path = G[f"path_{int(G['path_index'][k])}"]
        obs, prog = O.internal_obs(G["agent"][k], path)
        assert abs(prog - G["progress"][k]) <= 1e-12 * max(1.0, abs(prog))
        worst = max(worst, float(np.abs(obs - G["internal"][k]).max()))
    assert worst <= 6e-8          # fp32 observation: at most the last bit (libm vs CPython rounding)


def test_rl_ref_matches_reference_robot():
    for use_libm, tol in ((True, 1e-14), (False, 1e-13)):
        for k in range(len(G["agent"])):
            rl = O.rl_ref(G["agent"][k], int(G["action"][k]), use_libm=use_libm)
            np.testing.assert_allclose(rl, G["rl_ref"][k], rtol=0, atol=tol)


def test_projection_edge_cases():
    sq = np.array([(0.0, 0.0), (4.0, 0.0), (4.0, 3.0)])
    P = lambda x, y: O.load().ttdqn_oracle_project(sq.ctypes.data, len(sq), x, y)
    assert P(-2.0, 1.0) == 0.0                       # before the start: clamped
    assert P(9.0, 9.0) == 7.0                        # beyond the end: total length
    assert P(2.0, -1.0) == 2.0                       # interior of the first segment
    assert P(2.0, 5.0) == 7.0                        # closer to the far end of the second segment
    assert P(5.0, -1.0) == 4.0                       # the corner, reached from the first segment
