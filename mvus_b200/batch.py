"""Config 5 of BASELINE.json: a batch of INDEPENDENT reconstructions (multi-sequence throughput).

Problems are partitioned across ranks (problem p -> rank p mod world); there is no collective on
the data path, only the final gather of the per-problem results ("scaling": "weak").  Every
problem goes through the same ``Scene.BA`` drop-in (one handle per problem; device and pinned
memory come from the process-level pools, so creating a handle per problem costs microseconds
of allocation).
"""
import numpy as np


def my_problems(n_problems, rank, world):
    """Indices of the problems rank `rank` owns (round robin: balanced for any n)."""
    return list(range(rank, n_problems, world))


def solve_many(scenes, numCam=None, rank=0, world=1, **ba_kw):
    """Run Scene.BA on every scene this rank owns.  Returns {index: OptimizeResult}."""
    out = {}
    for p in my_problems(len(scenes), rank, world):
        sc = scenes[p]
        out[p] = sc.BA(numCam or sc.numCam, **ba_kw)
    return out


def gather_costs(results, n_problems):
    """Final gather of the per-problem costs over torch.distributed (no-op without a process
    group): returns an array of length n_problems on every rank."""
    costs = np.zeros(n_problems)
    for p, r in results.items():
        costs[p] = r.cost
    try:
        import torch
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            t = torch.from_numpy(costs)
            if dist.get_backend() == 'nccl':
                t = t.cuda()
            dist.all_reduce(t)
            costs = t.cpu().numpy()
    except ImportError:
        pass
    return costs
