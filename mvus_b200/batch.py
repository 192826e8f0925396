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


def solve_many(scenes, numCam=None, rank=0, world=1, threads=8, **ba_kw):
    """Run Scene.BA on every scene this rank owns.  Returns {index: OptimizeResult}.

    Every problem has its own handle = its own CUDA stream, and the ctypes calls release the GIL, so
    `threads` host threads keep that many small problems in flight on the GPU at once: a 7 x 5000
    problem occupies a few per cent of a B200 and its LM loop is latency-bound (host round trips for the
    accept / reject scalars), which is what the concurrency hides."""
    mine = my_problems(len(scenes), rank, world)
    return solve_scenes({p: scenes[p] for p in mine}, numCam=numCam, threads=threads, **ba_kw)


def solve_scenes(scenes, numCam=None, threads=8, **ba_kw):
    """{index: scene} -> {index: OptimizeResult}, `threads` problems in flight."""
    if threads <= 1 or len(scenes) <= 1:
        return {p: sc.BA(numCam or sc.numCam, **ba_kw) for p, sc in scenes.items()}
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(threads, len(scenes))) as pool:
        futs = {p: pool.submit(sc.BA, numCam or sc.numCam, **ba_kw) for p, sc in scenes.items()}
        return {p: f.result() for p, f in futs.items()}


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
