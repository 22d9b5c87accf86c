#!/usr/bin/env python
"""observe + Q-network on 16384 environments: tensor-core path (default) against the FMA-pipe path."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trajtrack_mpcndqn_rlboost_b200 as t
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(ROOT, "tests", "golden", "qnet_ray.npz"))
wq = t.dqn.QNetWeights(*[g[k] for k in ("w0", "b0", "w1", "b1", "w2", "b2")])
lay = t.dqn.default_layout(max_poly=8, max_vert=160)
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
base = [t.geometry.pad_polygon_round(np.array([(3., 3.), (3., 7.), (7., 7.), (7., 3.)]), 0.5),
        t.geometry.pad_polygon_round(np.array([(12., 2.), (12., 9.), (15., 9.), (15., 2.)]), 0.5),
        t.geometry.pad_polygon_round(np.array([(5., 12.), (5., 15.), (16., 15.), (16., 12.)]), 0.5),
        np.array([(0.5, 0.5), (19.5, 0.5), (19.5, 19.5), (0.5, 19.5)])]
xy1, off1, sol1, cnt1 = t.dqn.pack_geometry(lay, [base], [[True, True, True, False]])
xy = torch.from_numpy(np.repeat(xy1, n, 0)).cuda(); off = torch.from_numpy(np.repeat(off1, n, 0)).cuda()
sol = torch.from_numpy(np.repeat(sol1, n, 0)).cuda(); cnt = torch.from_numpy(np.repeat(cnt1, n, 0)).cuda()
agent = torch.from_numpy(np.c_[rng.uniform(1, 19, (n, 2)), rng.uniform(-3, 3, n)]).cuda()
internal = torch.from_numpy(np.resize(g["internal"], (n, 14)).astype(np.float32)).cuda()
comp = t.dqn.DqnCompanion(lay, wq)
qs = wq.device_struct()
out = {}
res = {}
for mode in ("mma", "fma"):
    os.environ["TTDQN_QNET"] = mode
    old = torch.zeros(n, 16, dtype=torch.float32, device="cuda")
    bufs = dict(ext=torch.zeros(n, 32, dtype=torch.float32, device="cuda"), q=torch.zeros(n, 9, dtype=torch.float32, device="cuda"),
                action=torch.zeros(n, dtype=torch.int32, device="cuda"), seg=torch.zeros(n, 8, dtype=torch.float64, device="cuda"),
                ray=torch.zeros(n, 8, dtype=torch.float64, device="cuda"))
    for _ in range(3): comp.observe_act_device(agent, xy, off, sol, cnt, internal, old, bufs, qs)
    torch.cuda.synchronize()
    old.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): comp.observe_act_device(agent, xy, off, sol, cnt, internal, old, bufs, qs)
    e1.record(); torch.cuda.synchronize()
    out[mode] = e0.elapsed_time(e1) / 20
    res[mode] = (bufs["q"].cpu().numpy().copy(), bufs["action"].cpu().numpy().copy())
dq = float(np.abs(res["mma"][0] - res["fma"][0]).max())
print(json.dumps(dict(n_envs=n, ms_observe_plus_qnet_mma=out["mma"], ms_observe_plus_qnet_fma=out["fma"],
                      max_abs_dq=dq, actions_equal=float((res["mma"][1] == res["fma"][1]).mean()))))
