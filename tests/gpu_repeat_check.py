"""Race detector for the solver kernels (run on a GPU box): the same LM solve repeated in one process must
reproduce itself to rounding (FP64 RED order is the only non-determinism).  Prints the spread."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mvus_b200 import _cabi
from mvus_b200.problem import FlatProblem

cams, det, coef = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 300000, 200000)))
fl = bench.make_workload(cams, det, coef)
fp = FlatProblem(fl, fl.numCam, **bench.BA_KW)
xs, costs = [], []
for rep in range(4):
    hd = _cabi.Handle(fp, ftol=0.0, xtol=0.0, gtol=0.0, max_nfev=5)
    x, r, st = hd.solve(fp.x0, want_r=False)
    hd.close()
    xs.append(x)
    costs.append(st.cost)
    print('rep', rep, 'cost %.12e' % st.cost, 'solves', st.lm_iterations, flush=True)
step = np.abs(xs[0] - fp.x0).max()
for k in range(1, len(xs)):
    print('rep %d vs 0: max |dx| / max step = %.3e   cost rel %.3e' % (
        k, np.abs(xs[k] - xs[0]).max() / step, abs(costs[k] - costs[0]) / costs[0]))
