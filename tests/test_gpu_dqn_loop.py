"""DQN side of the hybrid loop on the GPU against the CPU oracle and the reference goldens."""
import os

import numpy as np
import pytest

import trajtrack_mpcndqn_rlboost_b200 as t
from tests import oracle_lib as O

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "dqn_loop.npz"))


def _device_inputs():
    import torch
    paths = [G[f"path_{int(i)}"] for i in G["path_index"]]
    xy, cnt = t.dqn.pack_paths(paths, max_nodes=6)
    return (torch.tensor(G["agent"]).cuda(), torch.tensor(xy).cuda(), torch.tensor(cnt).cuda(), paths)


def test_internal_obs_parity():
    agent, xy, cnt, paths = _device_inputs()
    obs, prog = t.dqn.internal_obs_device(agent, xy, cnt)
    obs, prog = obs.cpu().numpy(), prog.cpu().numpy()
    ref = [O.internal_obs(G["agent"][k], paths[k]) for k in range(len(paths))]
    assert np.array_equal(prog, np.array([r[1] for r in ref]))            # fp64 polyline walk: exact
    assert np.abs(obs - np.array([r[0] for r in ref])).max() <= 6e-8      # fp32, CUDA vs glibc libm
    assert np.abs(obs - G["internal"]).max() <= 6e-8                      # the reference's components


def test_rl_ref_parity():
    import torch
    agent, _, _, _ = _device_inputs()
    act = torch.tensor(G["action"].astype(np.int32)).cuda()
    rl = t.dqn.rl_ref_device(agent, act).cpu().numpy()
    ref = np.array([O.rl_ref(G["agent"][k], int(G["action"][k]), use_libm=False) for k in range(len(rl))])
    assert np.array_equal(rl, ref)                                        # same tt_sincos: bit for bit
    np.testing.assert_allclose(rl, G["rl_ref"], rtol=0, atol=1e-13)       # the reference's MobileRobot


def test_hint_feeds_the_planner():
    """observe -> act -> rl_ref -> local reference with the DQN hint -> solve, all on the device side of the API."""
    import torch
    agent, xy, cnt, paths = _device_inputs()
    internal, _ = t.dqn.internal_obs_device(agent, xy, cnt)
    assert internal.shape == (len(paths), 14) and bool(torch.isfinite(internal).all())
    act = torch.tensor(G["action"].astype(np.int32)).cuda()
    rl = t.dqn.rl_ref_device(agent, act)
    assert bool(torch.isfinite(rl).all())
    step = torch.linalg.norm(rl[:, 1:] - rl[:, :-1], dim=-1)
    assert float(step.max()) <= 0.2 * 1.0 + 1e-9                           # ts * ref_speed per step
