"""Diagnostic (not a test): host-phase breakdown of Scene.BA on config-5-sized problems
(7 cameras x 5000 detections).  Run on a GPU box: python tests/cfg5_probe.py"""
import sys, time, io, contextlib
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mvus_b200 import synth
BA_KW = dict(rs=True, motion_reg=True, motion_weights=1e4)
scenes = [synth.make_flight(nc=7, det_per_cam=5000, seed=7 * p, rolling_shutter=True, distortion=True, motion_type='F', motion_weights=1e4, uncovered=0.0)[0] for p in range(40)]
with contextlib.redirect_stdout(io.StringIO()):
    for s in scenes[:4]:
        s.BA(7, max_iter=9, **BA_KW)
    t0 = time.perf_counter()
    res = [s.BA(7, max_iter=9, **BA_KW) for s in scenes[4:]]
    dt = time.perf_counter() - t0
agg = {}
for r in res:
    for k, v in r.stats['host'].items():
        agg[k] = agg.get(k, 0) + v
n = len(res)
print('per problem ms', dt / n * 1e3, {k: round(v / n, 3) for k, v in agg.items()}, 'device ms', sum(r.stats['ms_total'] for r in res) / n, 'launches', sum(r.stats['launches'] for r in res) / n)
