import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import make_golden_points as mg
case = sys.argv[1] if len(sys.argv) > 1 else 'points_KE_calib'
it = int(sys.argv[2]) if len(sys.argv) > 2 else 10
fl, kw = mg.make(case)
res = fl.BA(fl.numCam, max_iter=it, motion_prior=True, **kw)
print('cost', res.cost, 'nfev', res.nfev, 'status', res.status, res.stats['lm_iterations'])
