"""Golden fixtures for Scene.BA(motion_prior=True) (the discrete-trajectory mode, common.py:466-467, 527-550,
587-605): the UNMODIFIED reference's error_BA closure, start vector and one shipped solve on seeded flights.
    python tests/golden/make_golden_points.py
Per case: x0 (reference layout, points interleaved), xs[k] / rs[k] = error_BA at x0 and at perturbed points,
global_traj as the reference built it (7 x G), the reference's own BA(max_iter=10) cost / nfev and the spline it
leaves behind (knots and coefficients of every interval)."""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import cases                                  # noqa: E402
from oracle import ref_shim                   # noqa: E402

CASES = {'points_F': ('rs_F_gap', dict(rs=True, motion_weights=1e2)),
         'points_KE_calib': ('calib_KE', dict(rs=True, motion_weights=1e1)),
         'points_F_nors': ('rs_F_gap', dict(rs=False, motion_weights=1e3))}
DET = 150


def make(case):
    name, kw = CASES[case]
    fl, truth, _ = cases.make(name, det_per_cam=DET)
    fl.settings['smooth_factor'] = [10, 20]
    return fl, dict(kw)


def main():
    for case in CASES:
        fl, kw = make(case)
        ref = ref_shim.to_reference_scene(fl)
        with contextlib.redirect_stdout(io.StringIO()):
            fn, x0, A, _ = ref_shim.capture_ba(ref, fl.numCam, motion_prior=True, **kw)
        gtraj = ref.global_traj.copy()
        rng = np.random.default_rng(5)
        n_other = fl.numCam * (3 + (15 if fl.settings['opt_calib'] else 6))
        xs = [x0.copy()]
        for scale in (1e-5, 1e-4):
            x = x0 + rng.normal(size=x0.shape) * scale * np.maximum(1.0, np.abs(x0))
            xs.append(x)
        x = x0.copy()
        x[n_other:] += rng.normal(size=len(x0) - n_other) * 0.05          # points only: 5 cm
        xs.append(x)
        rs = [np.asarray(fn(x)).copy() for x in xs]
        ref2 = ref_shim.to_reference_scene(fl)
        with contextlib.redirect_stdout(io.StringIO()):
            res = ref2.BA(fl.numCam, max_iter=10, motion_prior=True, **kw)
        out = dict(x0=x0, xs=np.array(xs), rs=np.array(rs), global_traj=gtraj, shipped_cost=res.cost,
                   shipped_nfev=res.nfev, shipped_x=res.x, spline_int=np.asarray(ref2.spline['int']),
                   n_tck=len(ref2.spline['tck']), traj_after=np.asarray(ref2.traj))
        for s, t in enumerate(ref2.spline['tck']):
            out['knots_%d' % s] = np.asarray(t[0])
            out['coefs_%d' % s] = np.asarray(t[1])
            out['deg_%d' % s] = int(t[2])
        np.savez_compressed(os.path.join(HERE, case + '.npz'), **out)
        print(case, 'n', len(x0), 'm', len(rs[0]), 'G', gtraj.shape[1], 'cost0', 0.5 * rs[0] @ rs[0], 'shipped', res.cost, res.nfev)


if __name__ == '__main__':
    main()
