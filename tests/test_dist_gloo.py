"""World-size-2 test of the N>1 logic on CPU (gloo): scenes shard across ranks with no
collective in the solve; only the timing/result gather uses the process group.  The
solve itself is replaced by the oracle here (no GPU in this container)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import trajtrack_mpcndqn_rlboost_b200 as t
    from tests import oracle_lib as O
    # what bench.py builds: the same R seeded batches on every rank, rank r starting its rotation at
    # batch r (equal work on every rank over K = multiple of R steps, different scenes at any one step)
    t.scenes.WORKLOADS["tiny"] = dict(n=6, n_static=2, n_dynamic=1, blocking_fraction=0.0, solver={})
    R = 2
    batches = [bench.build_workload("tiny", rank, world, j)[1] for j in range(R)]
    cfg = bench.build_workload("tiny", rank, world, 0)[0]
    rot = rank % R
    p = batches[(0 + rot) % R]            # the batch of step 0 on this rank
    np.save(os.path.join(out_dir, f"all{rank}.npy"), np.stack(batches))
    sol = O.solve_batch(cfg, p, warp=True)
    # max-over-ranks timing + gathered metrics, like bench.py
    tns = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tns, op=dist.ReduceOp.MAX)
    hist = torch.from_numpy(np.bincount(sol["exit_status"], minlength=4).astype(np.int64))
    gathered = [torch.zeros_like(hist) for _ in range(world)]
    dist.all_gather(gathered, hist)
    np.save(os.path.join(out_dir, f"p{rank}.npy"), p)
    if rank == 0:
        np.save(os.path.join(out_dir, "max.npy"), tns.numpy())
        np.save(os.path.join(out_dir, "hist.npy"), torch.stack(gathered).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    p0, p1 = np.load(tmp_path / "p0.npy"), np.load(tmp_path / "p1.npy")
    assert p0.shape == p1.shape == (6, 2658)
    assert not np.array_equal(p0, p1)                    # at one step the ranks solve different scenes
    a0, a1 = np.load(tmp_path / "all0.npy"), np.load(tmp_path / "all1.npy")
    assert np.array_equal(a0, a1)                         # ... drawn from the same seeded batches
    assert np.array_equal(p0, a0[0]) and np.array_equal(p1, a0[1])   # rotation offset = rank
    assert float(np.load(tmp_path / "max.npy")[0]) == 2.0  # max over ranks
    hist = np.load(tmp_path / "hist.npy")
    assert hist.shape == (2, 4) and hist.sum() == 12      # all 12 scenes accounted for
