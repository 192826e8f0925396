"""Multi-GPU sharding of one BA: one process per GPU, detections split by
camera and time chunk, NCCL all-reduce of the normal equations inside the CUDA library.

Host-side logic only (what rank owns which detections, communicator bootstrap through
torch.distributed); it is exercised on CPU with the gloo backend in tests/test_shard_gloo.py.
"""
import copy

import numpy as np


def chunk_bounds(n, world):
    """Contiguous, balanced split of n items into `world` chunks: bounds[r]..bounds[r+1]."""
    return [(n * r) // world for r in range(world + 1)]


def shard_detections(detections, rank, world):
    """Every camera's (time-sorted, common.py:1190) detection track is cut into `world`
    contiguous time chunks; rank r keeps chunk r of every camera.  Each detection lands on
    exactly one rank, each rank sees every camera (so every camera block gets contributions
    everywhere and the all-reduce sums them)."""
    out = []
    for d in detections:
        b = chunk_bounds(d.shape[1], world)
        out.append(np.ascontiguousarray(d[:, b[rank]:b[rank + 1]]))
    return out


def shard_scene(scene, rank, world):
    """Shallow copy of `scene` whose detections are this rank's shard (parameters and splines
    are replicated)."""
    s = copy.copy(scene)
    s.detections = shard_detections(scene.detections, rank, world)
    s.detections_global = []
    s.cameras = [copy.copy(c) for c in scene.cameras]
    s.spline = {'tck': [[t[0], list(t[1]), t[2]] for t in scene.spline['tck']],
                'int': scene.spline['int']}
    s.alpha = np.array(scene.alpha, dtype=np.float64)
    s.beta = np.array(scene.beta, dtype=np.float64)
    s.rs = np.array(scene.rs, dtype=np.float64)
    return s


def init_comm():
    """Bootstrap the library's NCCL communicator from an initialised torch.distributed process
    group: rank 0 creates the NCCL unique id, it is broadcast as a byte tensor, and every
    later Handle joins it (mvus_b200.ba._COMM)."""
    import ctypes
    import torch
    import torch.distributed as dist
    from . import _cabi, ba
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        ba._COMM = None
        return
    uid = ctypes.create_string_buffer(128)
    if rank == 0:
        rc = _cabi.load().mvus_ba_nccl_unique_id(uid)
        if rc != 0:
            raise _cabi.MvusError('mvus_ba_nccl_unique_id failed (%d)' % rc)
    t = torch.tensor(list(uid.raw), dtype=torch.uint8)
    if dist.get_backend() == 'nccl':
        t = t.cuda()
    dist.broadcast(t, 0)
    ba._COMM = (world, rank, bytes(t.cpu().tolist()))
