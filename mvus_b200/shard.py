"""Multi-GPU sharding of one BA: one process per GPU, detections split along the
GLOBAL TIME AXIS so that rank r holds the detections whose knot spans fall into the control-point
range rank r owns in the sharded solve (mvus_ba_shard_bounds); the exchange of the normal
equations inside the CUDA library then only moves the halo rows at the range boundaries.

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


def span_index(scene, cam):
    """Global index of the last active control point of every detection of camera `cam` at the
    scene's current parameters (what K1 calls the span), -1 where no spline interval covers the time
    stamp: t = alpha (f + rho y / H) + beta (common.py:125), membership (t >= a) xor (t >= b)
    (util.py:103-106), FITPACK span lookup.  Host-side, only used to CUT the shards."""
    d = np.asarray(scene.detections[cam], dtype=np.float64)
    t = scene.alpha[cam] * (d[0] + scene.rs[cam] * d[2] / scene.cameras[cam].resolution[1]) + scene.beta[cam]
    interval = np.asarray(scene.spline['int'], dtype=np.float64).reshape(2, -1)
    g = np.full(d.shape[1], -1, dtype=np.int64)
    off = 0
    for s, tck in enumerate(scene.spline['tck']):
        knots, k, nco = np.asarray(tck[0], dtype=np.float64), int(tck[2]), len(tck[1][0])
        m = np.logical_xor(t - interval[0, s] >= 0, t - interval[1, s] >= 0)
        if m.any():
            g[m] = off + np.clip(np.searchsorted(knots, t[m], side='right') - 1, k, nco - 1)
        off += nco
    return g


def shard_bounds(scene, world, motion_reg=False):
    """Control-point bounds of the ranks' block ranges, from the library (needs a GPU)."""
    from . import _cabi, ba
    from .problem import FlatProblem

    class _S:
        pass
    s = _S()
    s.__dict__.update({k: getattr(scene, k) for k in ('settings', 'alpha', 'beta', 'rs', 'spline', 'cameras')})
    s.detections = list(scene.detections)
    s.detections[0] = np.zeros((3, 0))
    s.sequence = [0]
    fp = FlatProblem(s, 1, motion_reg=motion_reg)
    hd = _cabi.Handle(fp, device=ba.DEVICE)
    try:
        return hd.shard_bounds(world)
    finally:
        hd.close()


def shard_detections_by_span(scene, rank, bounds):
    """Rank r keeps the detections whose span lies in [bounds[r], bounds[r+1]); a detection that no
    interval covers goes where the last covered detection before it went (it only has to be
    somewhere, exactly once).  Time-sorted tracks give one contiguous slice per camera."""
    out = []
    for cam in range(len(scene.detections)):
        g = span_index(scene, cam)
        idx = np.where(g >= 0, np.arange(len(g)), 0)
        np.maximum.accumulate(idx, out=idx)                  # forward fill of the last covered detection
        gf = np.where(g[idx] >= 0, g[idx], 0)
        owner = np.searchsorted(np.asarray(bounds)[1:-1], gf, side='right')
        out.append(np.ascontiguousarray(np.asarray(scene.detections[cam])[:, owner == rank]))
    return out


def shard_scene(scene, rank, world, bounds=None):
    """Shallow copy of `scene` whose detections are this rank's shard (parameters and splines
    are replicated).  With `bounds` (shard_bounds) the cut follows the solver's owner ranges;
    without, every camera's track is cut into equal counts (correct, but the library then has to
    move whole block ranges between ranks instead of halos)."""
    s = copy.copy(scene)
    s.detections = shard_detections(scene.detections, rank, world) if bounds is None else \
        shard_detections_by_span(scene, rank, bounds)
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
