#!/usr/bin/env python
"""Golden vectors for the DQN Q-network from the reference's trained model.

Loads /root/reference/Model/ray/best_model.zip (SB3 1.6.2 DQN, MultiInputPolicy,
net_arch [16, 16]; test_block_rl.py:52-56, main.py:69) and evaluates the policy's
q_net with plain torch fp32: SB3's QNetwork is
    q_values = Sequential(Linear(46,16), ReLU, Linear(16,16), ReLU, Linear(16,9))(
                   cat([flatten(obs['external']), flatten(obs['internal'])], dim=1))
(CombinedExtractor iterates the Dict space's keys in sorted order) and
``predict(deterministic=True)`` returns ``q_values.argmax(dim=1)``.
stable_baselines3 itself is not installed here; the state_dict keys and shapes in
policy.pth (q_net.q_net.{0,2,4}) fix the architecture.
Writes tests/golden/qnet_ray.npz (weights + 512 observations + Q + actions).
"""
import io
import os
import zipfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ZIP = "/root/reference/Model/ray/best_model.zip"


def main():
    with zipfile.ZipFile(ZIP) as z:
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    net = torch.nn.Sequential(torch.nn.Linear(46, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16),
                              torch.nn.ReLU(), torch.nn.Linear(16, 9))
    net.load_state_dict({k.replace("q_net.q_net.", ""): v for k, v in sd.items()
                         if k.startswith("q_net.q_net.")})
    rng = np.random.default_rng(11)
    K = 512
    ext = rng.uniform(0, 1, (K, 32)).astype(np.float32)          # Box(0,1) external
    ext[:64, :] = 1.0                                             # nothing in sight (inf -> 1)
    ext[64:128, 16:] = 0.0                                        # first step: empty memory
    internal = rng.uniform(-1, 1, (K, 14)).astype(np.float32)
    internal[:, [4, 7, 10, 13]] = rng.uniform(0, 1, (K, 4))       # normalised distances
    obs = np.concatenate([ext, internal], axis=1)
    with torch.no_grad():
        q = net(torch.from_numpy(obs)).numpy()
    act = q.argmax(axis=1).astype(np.int32)
    g = lambda k: sd[k].numpy()
    out = os.path.join(ROOT, "tests", "golden", "qnet_ray.npz")
    np.savez_compressed(out, w0=g("q_net.q_net.0.weight"), b0=g("q_net.q_net.0.bias"),
                        w1=g("q_net.q_net.2.weight"), b1=g("q_net.q_net.2.bias"),
                        w2=g("q_net.q_net.4.weight"), b2=g("q_net.q_net.4.bias"),
                        ext=ext, internal=internal, q=q, action=act)
    gap = np.sort(q, axis=1)
    print("wrote", out, "action histogram", np.bincount(act, minlength=9),
          "min top-2 gap", (gap[:, -1] - gap[:, -2]).min())


if __name__ == "__main__":
    main()
