"""world_size-2 gloo test (CPU) of bench.py's multi-rank host logic: per-rank seeds differ (weak scaling), contiguous
slices of one global batch and the one all_gather of the metric block (strong scaling), timings reduced with MAX over
ranks, only rank 0 of the reference arm prints a line."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = [10.0 + rank, 5.0 - rank]
    red = bench.aggregate_max(local)
    from interdiff_b200 import synthetic as S
    b = S.make_smpl_batch(B=2, T=30, seed=bench.rank_seed(rank))
    # strong scaling: contiguous slices of ONE global batch + the path's only collective (one all_gather of the (6, B/G) block)
    from interdiff_b200.sampling import gather_metrics
    gb = S.make_smpl_batch(B=4, T=30, seed=233)
    sl = bench.rank_slice(4, rank, world)
    mine = torch.from_numpy(gb["gt"][sl]).flatten(1).sum(1)                    # a per-sample quantity of this rank's slice
    block = torch.stack([mine * (k + 1) for k in range(6)])                    # (6, B/G)
    allm = gather_metrics(block)
    out[rank] = (red, float(b["gt"].sum()), allm.tolist(), [sl.start, sl.stop])
    dist.barrier()
    dist.destroy_process_group()


def test_max_over_ranks_and_rank_seeds():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, 29511, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert r0[0] == r1[0] == [11.0, 5.0]          # MAX over ranks, identical on every rank
    assert r0[1] != r1[1]                          # different sequences per rank
    # all_gather: both ranks hold the (6, 4) block in GLOBAL sample order = what a single rank computes on the whole batch
    sys.path.insert(0, ROOT)
    from interdiff_b200 import synthetic as S
    gb = S.make_smpl_batch(B=4, T=30, seed=233)
    want = torch.stack([torch.from_numpy(gb["gt"]).flatten(1).sum(1) * (k + 1) for k in range(6)]).tolist()
    assert r0[2] == r1[2] == want
    assert r0[3] == [0, 2] and r1[3] == [2, 4]


def test_reference_arm_prints_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_aggregate_without_process_group():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.aggregate_max([1.5, 2.5]) == [1.5, 2.5]
