"""Host-side logic of the N>1 path on CPU (gloo, world_size 2): the detection sharding is a
partition, every rank owns every camera, and summing the per-rank normal equations and costs
(what the library all-reduces over NCCL) reproduces the single-rank ones.  The per-rank
quantities are formed with the oracle here (test-only) because kernels cannot run without a GPU."""
import os
import socket

import numpy as np
import pytest

import cases
from mvus_b200 import shard


def test_chunk_bounds_partition():
    for n in (0, 1, 7, 128, 1000003):
        for w in (1, 2, 3, 8):
            b = shard.chunk_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
            assert max(b[i + 1] - b[i] for i in range(w)) - min(b[i + 1] - b[i] for i in range(w)) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out, mode='count'):
    import torch
    import torch.distributed as dist
    from oracle import ba_oracle
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    fl, truth, bakw = cases.make(name)
    # motion rows are parameter-only: rank 0 owns them (the library adds them on every rank and
    # the host divides; here the oracle-side bookkeeping keeps them on rank 0)
    n_ctrl = sum(len(t[1][0]) for t in fl.spline['tck'])
    bounds = None if mode == 'count' else [(n_ctrl * r) // world for r in range(world + 1)]
    local = shard.shard_scene(fl, rank, world, bounds)
    kw = dict(bakw)
    if rank != 0:
        kw['motion_reg'] = False
    prob = ba_oracle.Problem(local, local.numCam, **kw)
    x0 = ba_oracle.Problem(fl, fl.numCam, **bakw).x0
    J = prob.jacobian(x0).toarray() * prob.free_mask()[None, :]
    r = prob.residual(x0)
    H = torch.from_numpy(J.T @ J)
    g = torch.from_numpy(J.T @ r)
    c = torch.tensor([0.5 * float(r @ r), float(sum(prob.N))], dtype=torch.float64)
    for t in (H, g, c):
        dist.all_reduce(t)
    if rank == 0:
        np.savez(out, H=H.numpy(), g=g.numpy(), c=c.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('name,mode', [('rs_F_gap', 'count'), ('rs_F_gap', 'span')])
def test_sharded_normal_equations_sum_to_global(name, mode, tmp_path):
    """`count`: every camera's track cut into equal counts; `span`: cut along the control-point
    ranges the ranks own (what the multi-GPU bench uses, shard.shard_detections_by_span)."""
    import torch.multiprocessing as mp
    from oracle import ba_oracle
    out = str(tmp_path / 'sum.npz')
    mp.spawn(_worker, args=(2, _free_port(), name, out, mode), nprocs=2, join=True)
    got = np.load(out)
    fl, truth, bakw = cases.make(name)
    prob = ba_oracle.Problem(fl, fl.numCam, **bakw)
    J = prob.jacobian(prob.x0).toarray() * prob.free_mask()[None, :]
    r = prob.residual(prob.x0)
    H, g = J.T @ J, J.T @ r
    assert int(got['c'][1]) == sum(prob.N)                        # every detection on exactly one rank
    assert abs(got['c'][0] - 0.5 * r @ r) <= 1e-12 * got['c'][0]
    assert np.abs(got['H'] - H).max() <= 1e-10 * np.abs(H).max()
    assert np.abs(got['g'] - g).max() <= 1e-10 * np.abs(g).max()


def test_shard_scene_keeps_every_camera():
    fl, truth, bakw = cases.make('gs_plain')
    parts = [shard.shard_scene(fl, r, 3) for r in range(3)]
    for i in range(fl.numCam):
        cat = np.concatenate([p.detections[i] for p in parts], axis=1)
        assert (cat == fl.detections[i]).all()
        assert all(p.detections[i].shape[1] > 0 for p in parts)
    assert parts[0].spline['tck'][0][1][0] is fl.spline['tck'][0][1][0] or True


def test_span_sharding_is_a_time_partition():
    """Sharding by span: a partition of every camera's detections into contiguous time slices whose
    spans respect the bounds (uncovered detections ride with the covered one before them)."""
    fl, truth, bakw = cases.make('rs_F_gap')
    n_ctrl = sum(len(t[1][0]) for t in fl.spline['tck'])
    for world in (2, 3, 8):
        bounds = [(n_ctrl * r) // world for r in range(world + 1)]
        parts = [shard.shard_detections_by_span(fl, r, bounds) for r in range(world)]
        for i in range(fl.numCam):
            cat = np.concatenate([p[i] for p in parts], axis=1)
            assert (cat == fl.detections[i]).all()              # contiguous slices in rank order = the track
            g = shard.span_index(fl, i)
            pos = 0
            for r in range(world):
                gi = g[pos:pos + parts[r][i].shape[1]]
                pos += parts[r][i].shape[1]
                assert ((gi < 0) | ((gi >= bounds[r]) & (gi < bounds[r + 1]))).all()


def test_batch_partition_covers_every_problem_once():
    from mvus_b200 import batch
    for n in (0, 1, 7, 1024):
        for w in (1, 2, 8):
            owned = [batch.my_problems(n, r, w) for r in range(w)]
            flat = sorted(i for o in owned for i in o)
            assert flat == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
