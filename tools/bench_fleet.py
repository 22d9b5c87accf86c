#!/usr/bin/env python
"""Closed-loop fleet step on one GPU: n robots, K control steps without touching the host.

  python tools/bench_fleet.py [n=4096] [steps=30]

Per step = pack + solve + advance (ttmpc_fleet_step_device).  Prints one JSON line with the
p50 / p90 step time, robot-steps/s, the share of pack and advance, the HBM rate of the pack
kernel against MEASURED_PEAKS.json, and the CPU oracle (reference order, all host cores) doing the
same step on a bounded sample.  Appended to gpurun_out/bench_fleet.json.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import trajtrack_mpcndqn_rlboost_b200 as t
from trajtrack_mpcndqn_rlboost_b200.fleet import work_mode
from tests import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    mc = t.Configurator()
    fl = t.scenes.make_fleet(n, seed=1000)
    fp = t.FleetPlanner(mc, fl["init"], fl["goal"], fl["paths"], mode="work")
    fp.update_static_constraints(fl["static_polys"], per_robot=True)
    fp.set_moving_obstacles(fl["moving_pos"], fl["moving_disp"])
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(3):          # warm-up steps (they do advance the fleet)
        fp.step()
    torch.cuda.synchronize()
    t_step, t_pack, t_adv, iters, running = [], [], [], [], []
    for k in range(steps):
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record(); fp.pack(); e1.record()
        r = fp._result_struct()
        import ctypes as C
        from trajtrack_mpcndqn_rlboost_b200 import _lib
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(fp.lib.ttmpc_solve_batch_device(C.byref(fp.cfg), n, fp.p.data_ptr(), 0, 0, None, C.byref(r), st), "solve")
        e2.record(); fp.advance(); e3.record()
        torch.cuda.synchronize()
        t_step.append(e0.elapsed_time(e3)); t_pack.append(e0.elapsed_time(e1)); t_adv.append(e2.elapsed_time(e3))
        iters.append(float(fp.inner.float().mean())); running.append(int((fp.status == 0).sum()))
    q = lambda a, p: float(np.quantile(a, p))
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    pack_bytes = fp.p.numel() * 8 + n * (fp.N * 3 * 8 + 12 * 8 + fp.stc.shape[1] * 8)   # row written + ref window, state, own static rows read
    out = dict(workload=f"fleet{n}", robots=n, steps=steps, step_ms_p50=q(t_step, 0.5), step_ms_p90=q(t_step, 0.9),
               robot_steps_per_s=n / (q(t_step, 0.5) * 1e-3), pack_ms_p50=q(t_pack, 0.5), advance_ms_p50=q(t_adv, 0.5),
               pack_GBps=pack_bytes / (q(t_pack, 0.5) * 1e-3) / 1e9, hbm_peak_GBps=peaks.get("hbm_gbs"),
               mean_inner_iters=float(np.mean(iters)), running_last=running[-1],
               step_ms=[round(x, 1) for x in t_step], inner_iters=[round(x, 1) for x in iters])
    # CPU: the same step with the oracle (reference operation order), bounded sample
    ns = min(n, 256)
    tuning, base = work_mode(mc, "work")
    fh = O.FleetHost(fp.cfg, fl["init"][:ns], fl["goal"][:ns], fp.ref_traj.cpu().numpy()[:ns], fp.ref_len.cpu().numpy()[:ns],
                     fp.stc.cpu().numpy()[:ns], tuning, base, mc.low_speed, dyn_cur=fl["moving_pos"][:ns], dyn_disp=fl["moving_disp"][:ns])
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    for _ in range(2):
        p = O.fleet_pack(fh, use_libm=True)
        ref = O.solve_batch(fp.cfg, p, threads=cores, warp=False)
        O.fleet_advance(fh, ref["u"], ref["exit_status"], use_libm=True)
    dt = (time.perf_counter() - t0) / 2
    out["cpu_baseline"] = dict(robot_steps_per_s=ns / dt, cores=cores, kind="port", sample=f"{ns} robots x 2 steps")
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_fleet.json"), "a") as f:
        f.write(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
