import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import trajtrack_mpcndqn_rlboost_b200 as t
g = np.load('/root/repo/tests/golden/qnet_ray.npz')
w = t.dqn.QNetWeights(*[g[k] for k in ("w0", "b0", "w1", "b1", "w2", "b2")])
lay = t.dqn.default_layout()
n = 512
xy, off, sol, cnt = t.dqn.pack_geometry(lay, [[] for _ in range(n)], [[] for _ in range(n)])
agent = np.zeros((n, 3))
old = g["ext"][:n, 16:].copy()
for mode in ("mma", "fma"):
    os.environ["TTDQN_QNET"] = mode
    out = t.dqn.DqnCompanion(lay, w).observe_act(agent, xy, off, sol, cnt, g["internal"][:n], old.copy())
    ok = np.all(out["ext"][:, :16] == 1, axis=1) & np.all(g["ext"][:n, :16] == 1, axis=1)
    d = np.abs(out["q"] - g["q"][:n])[ok]
    print(mode, "rows", int(ok.sum()), "max|dQ| vs torch", float(d.max()), "p99", float(np.quantile(d, 0.99)), "max|Q|", float(np.abs(g["q"][:n]).max()),
          "actions equal", float((out["action"] == g["action"][:n])[ok].mean()))
