"""Per-scene solve-time distribution of one batched solve (diagnostics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trajtrack_mpcndqn_rlboost_b200 as t
name = sys.argv[1] if len(sys.argv) > 1 else "static4096"
w = t.scenes.WORKLOADS[name]
cfg = t.Configurator().to_ttmpc(**w["solver"])
n_over = int(sys.argv[2]) if len(sys.argv) > 2 else w["n"]
p = t.scenes.make_scenes(max(w["n"], n_over), cfg, seed=1000, n_static=w["n_static"], n_dynamic=w["n_dynamic"],
                         blocking_fraction=w["blocking_fraction"])[:n_over]
s = t.BatchSolver(cfg)
dp = torch.from_numpy(p).cuda(); bufs = s.alloc_device(len(p))
for _ in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); s.run_device(dp, bufs); e1.record(); torch.cuda.synchronize()
ev = bufs["evals"].cpu().numpy()
dur = ev[:, 2] / 1e6; start = (ev[:, 3] - ev[:, 3].min()) / 1e6; n_ev = ev[:, 0] + ev[:, 1]
print("kernel ms", e0.elapsed_time(e1), "info", s.launch_info(len(p)))
print("per-scene ms: mean %.3f median %.3f p90 %.3f p99 %.3f max %.3f sum %.1f" % (dur.mean(), np.median(dur), *np.quantile(dur, [0.9, 0.99]), dur.max(), dur.sum()))
print("last finish ms %.3f; start of the slowest %.3f; evals of slowest %d; us/eval (slowest) %.2f; us/eval overall %.2f" % ((start + dur).max(), start[dur.argmax()], n_ev[dur.argmax()], 1e3 * dur.max() / n_ev[dur.argmax()], 1e3 * dur.sum() / n_ev.sum()))
order = np.argsort(-dur)[:8]
print("slowest:", [(int(i), round(float(dur[i]), 2), int(n_ev[i]), round(float(start[i]), 2)) for i in order])
li = s.launch_info(len(p)); slots = li["grid"] * li["block"] // 32
print("warp slots", slots, "ideal balanced ms", dur.sum() / slots)
fin = start + dur
late = np.argsort(-fin)[:8]
print("last to finish:", [(int(i), "start %.1f" % start[i], "dur %.1f" % dur[i], int(n_ev[i])) for i in late])
print("scenes still running at 70/80/90%% of the kernel: %s" % [int(((start < f * fin.max()) & (fin > f * fin.max())).sum()) for f in (0.7, 0.8, 0.9)])
