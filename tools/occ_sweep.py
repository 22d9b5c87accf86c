#!/usr/bin/env python
"""Diagnostic: batch time of a workload (argv[1], scenes argv[2]) for 1..3 resident blocks per SM."""
import os, sys, subprocess, json
if len(sys.argv) > 3:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np, torch
    import trajtrack_mpcndqn_rlboost_b200 as t
    name, n = sys.argv[1], int(sys.argv[2])
    w = t.scenes.WORKLOADS[name]
    cfg = t.Configurator().to_ttmpc(**w["solver"])
    p = t.scenes.make_scenes(n, cfg, seed=1000, n_static=w["n_static"], n_dynamic=w["n_dynamic"], blocking_fraction=w["blocking_fraction"])
    s = t.BatchSolver(cfg); dp = torch.from_numpy(p).cuda(); bufs = s.alloc_device(n)
    ms = []
    for it in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); s.run_device(dp, bufs); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    st = s.read_stats(reset=True)
    it = bufs["inner"].sum().item()
    print(json.dumps(dict(bps=os.environ.get("TTMPC_MAX_BLOCKS_PER_SM"), helpers=os.environ.get("TTMPC_NO_HELPERS"), ms=min(ms[2:]), iters=it, info=s.launch_info(n))))
else:
    for bps in os.environ.get("OCC_LIST", "1,2,3,4").split(","):
        env = dict(os.environ, TTMPC_MAX_BLOCKS_PER_SM=bps, TTMPC_NO_HELPERS="1")
        subprocess.run([sys.executable, __file__, sys.argv[1], sys.argv[2], "child"], env=env)
