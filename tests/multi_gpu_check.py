"""Run under torchrun on N GPUs (gpurun --gpus N): the sharded BA must reproduce the 1-GPU BA.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
Two shardings: `span` = along the control-point ranges the ranks own (mvus_ba_shard_bounds; only halo
rows move), `count` = equal counts per camera (whole block ranges move; must stay correct).
Prints PASS/FAIL lines on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.distributed as dist

import cases
from mvus_b200 import _cabi, ba, shard
from mvus_b200.problem import FlatProblem

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
ba.DEVICE = local
shard.init_comm()
ok_all = True
CHUNK = int(os.environ.get('MVUS_TEST_CHUNK', '0'))     # force the solver's pre-reduction chunk length
for name in ('rs_F_gap', 'calib_KE', 'gs_plain'):
    fl, truth, bakw = cases.make(name, det_per_cam=int(os.environ.get('MVUS_TEST_DET', '3000')))
    # single-GPU reference on every rank (no communicator)
    fp = FlatProblem(fl, fl.numCam, **bakw)
    h1 = _cabi.Handle(fp, device=local, max_nfev=12, solver_chunk=1)
    x1, r1, s1 = h1.solve(fp.x0)
    A1, g1, _, _, c1 = h1.normal_equations(fp.x0, want_dense=False)
    h1.close()
    for mode in ('span', 'count'):
        bounds = shard.shard_bounds(fl, world, motion_reg=bakw.get('motion_reg', False)) if mode == 'span' else None
        loc = shard.shard_scene(fl, rank, world, bounds)
        fpl = FlatProblem(loc, loc.numCam, **bakw)
        hN = _cabi.Handle(fpl, device=local, max_nfev=12, solver_chunk=CHUNK)
        hN.comm_init(*ba._COMM)
        AN, gN, _, _, cN = hN.normal_equations(fpl.x0, want_dense=False)
        xN, rN, sN = hN.solve(fpl.x0)
        hN.close()
        e_g = np.abs(gN - g1).max() / np.abs(g1).max()
        e_A = np.abs(AN - A1).max() / np.abs(A1).max()
        e_c = abs(cN - c1) / c1
        e_cost = abs(sN.cost - s1.cost) / s1.cost
        xs = torch.from_numpy(xN.copy()).cuda()
        x0r = xs.clone()
        dist.broadcast(x0r, 0)
        same = bool((xs == x0r).all().item())
        nloc = torch.tensor([float(fpl.N)], device='cuda')
        dist.all_reduce(nloc)
        good = (e_g < 1e-10 and e_A < 1e-10 and e_c < 1e-12 and e_cost < 1e-6 and same and sN.nfev == s1.nfev
                and int(nloc.item()) == fp.N)
        ok_all &= good
        if rank == 0:
            print('%s %s [%s]: world %d  grad %.2e  A %.2e  cost0 %.2e  final cost rel %.2e (%.6g vs %.6g)  nfev %d/%d  '
                  'x identical across ranks: %s  reduce %.2f ms of accumulate %.2f ms'
                  % ('PASS' if good else 'FAIL', name, mode, world, e_g, e_A, e_c, e_cost, sN.cost, s1.cost, sN.nfev,
                     s1.nfev, same, sN.ms_reduce, sN.ms_accum))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
