#!/usr/bin/env python
"""Measure every named configuration of BASELINE.json that fits one GPU and write BENCH.md.

  python tools/bench_all.py            (on the GPU box; ~2 minutes)

Rows: the NMPC solve on the workloads of scenes.WORKLOADS plus the horizon / obstacle sweep
(run-time-dimension kernel), each with the CPU oracle (reference order, all host cores) on a
bounded sample, and the DQN observe+act kernel on 16384 environments.
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import trajtrack_mpcndqn_rlboost_b200 as t
from trajtrack_mpcndqn_rlboost_b200 import _lib
from tests import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def time_solve(cfg, p, steps=5, warmup=3):
    s = t.BatchSolver(cfg)
    dp = torch.from_numpy(p).cuda()
    bufs = s.alloc_device(len(p))
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    ms = []
    for it in range(warmup + steps):
        if it == warmup:
            s.read_stats(reset=True)
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); s.run_device(dp, bufs); e1.record(); torch.cuda.synchronize()
        if it >= warmup:
            ms.append(e0.elapsed_time(e1))
    st = s.read_stats(reset=True)
    # the same batch several times with up to 4 of them in flight
    D = 4
    streams = [torch.cuda.Stream() for _ in range(D)]
    ring = [bufs] + [s.alloc_device(len(p)) for _ in range(D - 1)]
    n_fl = 8 if med_guess(ms) < 200 else 4
    for k in range(D):
        with torch.cuda.stream(streams[k]):
            s.run_device(dp, ring[k])
    torch.cuda.synchronize()
    ea = torch.cuda.Event(enable_timing=True); eb = torch.cuda.Event(enable_timing=True)
    ea.record()
    for st_ in streams: st_.wait_stream(torch.cuda.current_stream())
    for k in range(n_fl):
        with torch.cuda.stream(streams[k % D]):
            s.run_device(dp, ring[k % D])
    for st_ in streams: torch.cuda.current_stream().wait_stream(st_)
    eb.record(); torch.cuda.synchronize()
    in_flight = len(p) * n_fl / (ea.elapsed_time(eb) * 1e-3)
    s.read_stats(reset=True)
    t0 = time.perf_counter(); host = s.run(p); e2e = time.perf_counter() - t0
    t0 = time.perf_counter(); host = s.run(p); e2e = min(e2e, time.perf_counter() - t0)
    status = bufs["exit_status"].cpu().numpy()
    bodies = st["dyn_bodies"] / max(1, st["cost_evals"] + st["grad_evals"])
    flops = (st["cost_evals"] * bench.eval_flops(cfg, False, bodies) + st["grad_evals"] * bench.eval_flops(cfg, True, bodies)) / steps
    med = float(np.median(ms))
    return dict(ms_p50=med, ms_min=float(min(ms)), solves_per_s=len(p) / (med * 1e-3), in_flight_solves_per_s=in_flight,
                e2e_solves_per_s=len(p) / e2e, evals_per_solve=(st["cost_evals"] + st["grad_evals"]) / steps / len(p),
                tflops=flops / (med * 1e-3) / 1e12, converged=int((status == 0).sum()), n=len(p))


def med_guess(ms):
    return float(np.median(ms))


def time_cpu(cfg, p, sample):
    cores = os.cpu_count() or 1
    sample = min(sample, len(p))
    t0 = time.perf_counter(); O.solve_batch(cfg, p[:sample], threads=cores, warp=False); dt = time.perf_counter() - t0
    return dict(cpu_solves_per_s=sample / dt, cores=cores, sample=sample)


def main():
    peak = C.c_double()
    _lib.check(_lib.load().ttmpc_measure_fp64_peak(C.byref(peak), None), "peak")
    rows = []
    # named workloads
    for name in ("static4096", "mixed4096", "dynamic8192"):
        w = t.scenes.WORKLOADS[name]
        cfg = t.Configurator().to_ttmpc(**w["solver"])
        p = t.scenes.make_scenes(w["n"], cfg, seed=1000, n_static=w["n_static"], n_dynamic=w["n_dynamic"],
                                 blocking_fraction=w["blocking_fraction"])
        r = time_solve(cfg, p); r.update(time_cpu(cfg, p, 256)); r["name"] = name
        r["shape"] = f"N=20, {w['n_static']} static + {w['n_dynamic']} dynamic per scene, max {cfg.max_inner_iterations}x{cfg.max_outer_iterations}"
        rows.append(r); print(json.dumps(r), flush=True)
    # large batch on one GPU (throughput regime)
    w = t.scenes.WORKLOADS["static4096"]
    cfg = t.Configurator().to_ttmpc()
    p = t.scenes.make_scenes(32768, cfg, seed=1000, n_static=4, n_dynamic=0, blocking_fraction=0.1)
    r = time_solve(cfg, p, steps=3); r.update(time_cpu(cfg, p, 256)); r["name"] = "static32768"
    r["shape"] = "N=20, 4 static per scene, 32768 scenes on one GPU"
    rows.append(r); print(json.dumps(r), flush=True)
    # horizon / obstacle-count sweep (run-time-dimension kernel)
    for N, nst, ndy in [(10, 10, 15), (32, 10, 15), (20, 4, 4), (20, 20, 30)]:
        mc = t.Configurator(N_hor=N, Nstcobs=nst, Ndynobs=ndy)
        cfg = mc.to_ttmpc()
        p = t.scenes.make_scenes(4096, cfg, seed=1000, n_static=min(4, nst), n_dynamic=min(3, ndy), blocking_fraction=0.1)
        r = time_solve(cfg, p, steps=3); r.update(time_cpu(cfg, p, 128)); r["name"] = f"sweep N={N} Nstc={nst} Ndyn={ndy}"
        r["shape"] = f"4096 scenes, {min(4, nst)} static + {min(3, ndy)} dynamic active"
        rows.append(r); print(json.dumps(r), flush=True)
    # DQN companion: 16384 envs
    g = np.load(os.path.join(ROOT, "tests", "golden", "qnet_ray.npz"))
    wq = t.dqn.QNetWeights(*[g[k] for k in ("w0", "b0", "w1", "b1", "w2", "b2")])
    lay = t.dqn.default_layout(max_poly=8, max_vert=160)
    rng = np.random.default_rng(0)
    n = 16384
    base = [t.geometry.pad_polygon_round(np.array([(3., 3.), (3., 7.), (7., 7.), (7., 3.)]), 0.5),
            t.geometry.pad_polygon_round(np.array([(12., 2.), (12., 9.), (15., 9.), (15., 2.)]), 0.5),
            t.geometry.pad_polygon_round(np.array([(5., 12.), (5., 15.), (16., 15.), (16., 12.)]), 0.5),
            np.array([(0.5, 0.5), (19.5, 0.5), (19.5, 19.5), (0.5, 19.5)])]
    xy1, off1, sol1, cnt1 = t.dqn.pack_geometry(lay, [base], [[True, True, True, False]])
    xy = torch.from_numpy(np.repeat(xy1, n, 0)).cuda(); off = torch.from_numpy(np.repeat(off1, n, 0)).cuda()
    sol = torch.from_numpy(np.repeat(sol1, n, 0)).cuda(); cnt = torch.from_numpy(np.repeat(cnt1, n, 0)).cuda()
    agent = torch.from_numpy(np.c_[rng.uniform(1, 19, (n, 2)), rng.uniform(-3, 3, n)]).cuda()
    internal = torch.from_numpy(rng.uniform(-1, 1, (n, 14)).astype(np.float32)).cuda()
    old = torch.zeros(n, 16, dtype=torch.float32, device="cuda")
    out = dict(ext=torch.zeros(n, 32, device="cuda"), q=torch.zeros(n, 9, device="cuda"),
               action=torch.zeros(n, dtype=torch.int32, device="cuda"),
               seg=torch.zeros(n, 8, dtype=torch.float64, device="cuda"), ray=torch.zeros(n, 8, dtype=torch.float64, device="cuda"))
    comp = t.dqn.DqnCompanion(lay, wq); qs = wq.device_struct()
    ms = []
    for it in range(8):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); comp.observe_act_device(agent, xy, off, sol, cnt, internal, old, out, qs); e1.record()
        torch.cuda.synchronize()
        if it >= 3: ms.append(e0.elapsed_time(e1))
    nedges = int(off1[0, cnt1[0]])
    dq = dict(name="DQN observe+act", n=n, ms_p50=float(np.median(ms)), envs_per_s=n / (np.median(ms) * 1e-3),
              edges_per_env=nedges, hbm_GBps=(xy.numel() * 8 + off.numel() * 4 + n * (24 + 56 + 64 + 128 + 36 + 4 + 128)) / (np.median(ms) * 1e-3) / 1e9)
    print(json.dumps(dq), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(rows=rows, dqn=dq, fp64_peak_tflops=peak.value), open(os.path.join(ROOT, "gpurun_out", "bench_all.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
