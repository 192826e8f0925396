"""Generate the golden fixtures that pin the oracle to the REFERENCE ITSELF.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
For every seeded flight in tests/cases.py this script copies the flight into the
reference's own ``Scene``/``Camera`` classes (oracle/ref_shim.py, no source edits), captures
the reference's ``error_BA`` closure, start vector and ``jac_BA`` pattern from inside
``Scene.BA`` (common.py:665-670) and records, per flight:
    x0                      the reference's packed parameter vector (common.py:616-650)
    xs[k], rs[k]            error_BA evaluated by the reference at x0 and perturbed points
                            (one of them moves detections across interval edges)
    pattern (rows, cols)    the reference's jac_sparsity
    shipped_cost/nfev       the reference's own Scene.BA(max_iter=10) result (SciPy 1.18.1)
    book_*                  the pickled bookkeeping arrays (visible, global_traj,
                            global_detections, frame_id_all, global_time_stamps_all, traj) as the
                            reference leaves them after error_BA(x0) (ref_shim.reference_bookkeeping)
The flights themselves are rebuilt from their seeds (tests/cases.py), so nothing else is stored.
(main.py end to end is not a fixture: it runs live against oracle/_ref in
tests/test_gpu_reference.py::test_main_py_through_the_dropin.)
"""
import io
import os
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import cases                                  # noqa: E402
from oracle import ref_shim                   # noqa: E402


def points(x0, nc, seed=3):
    rng = np.random.default_rng(seed)
    out = [x0.copy()]
    for scale in (1e-4, 1e-3):
        out.append(x0 + rng.normal(size=x0.shape) * scale * np.maximum(1.0, np.abs(x0)))
    x = x0.copy()
    x[nc:2 * nc] += 0.5
    out.append(x)
    return out


def main():
    import cv2
    import scipy
    common = ref_shim.load()
    for name in cases.CASES:
        fl, truth, bakw = cases.make(name)
        ref = ref_shim.to_reference_scene(fl)
        with contextlib.redirect_stdout(io.StringIO()):
            fn, x0, A, kw = ref_shim.capture_ba(ref, fl.numCam, **bakw)
        xs = points(x0, fl.numCam)
        rs = [np.asarray(fn(x)).copy() for x in xs]
        rows, cols = np.nonzero(np.asarray(A))
        # the reference's own solve (oracle A), fresh scene
        ref2 = ref_shim.to_reference_scene(fl)
        with contextlib.redirect_stdout(io.StringIO()):
            res = ref2.BA(fl.numCam, **bakw)
        book = ref_shim.reference_bookkeeping(fl, fl.numCam, x=x0, **bakw)
        book = {'book_' + k: v for k, v in book.items()}
        out = os.path.join(HERE, 'ba_%s.npz' % name)
        np.savez_compressed(out, **book, x0=x0, xs=np.array(xs), rs=np.array(rs),
                            pat_rows=rows.astype(np.int32), pat_cols=rows.astype(np.int32) * 0 + cols.astype(np.int32),
                            pat_shape=np.array(np.asarray(A).shape), shipped_cost=res.cost,
                            shipped_nfev=res.nfev, shipped_x=res.x,
                            versions=np.array([np.__version__, scipy.__version__, cv2.__version__]))
        print(name, 'n', len(x0), 'm', len(rs[0]), 'shipped cost', res.cost, '->', os.path.basename(out))


if __name__ == '__main__':
    main()
