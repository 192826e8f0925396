"""CPU check of the arithmetic the CUDA kernels run: mvus_b200/csrc/ba_math.cuh is compiled
for the host (tests/emul) and compared with the oracle.  This is not the product path (the
product only runs the CUDA build); it lets the math be verified in the GPU-less container."""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
import helpers
from mvus_b200.problem import FlatProblem
from oracle import ba_oracle


@pytest.mark.parametrize('name', list(cases.CASES))
def test_emulated_kernel_math(name):
    fl, truth, bakw = cases.make(name)
    fp = FlatProblem(fl, fl.numCam, **bakw)
    prob = ba_oracle.Problem(fl, fl.numCam, **bakw)
    assert np.abs(fp.x0 - prob.x0).max() <= 1e-12 * max(1.0, np.abs(prob.x0).max())
    x = prob.x0
    r, span, J, mbase, mJ = helpers.emul_resjac(fp, x)
    ro = prob.residual(x)
    assert np.abs(r - ro).max() <= 1e-9 * max(1.0, np.abs(ro).max())
    Jg = helpers.expand_jacobian(fp, span, J, mbase, mJ)
    Jo = prob.jacobian(x).tocsc() @ sp.diags(prob.free_mask().astype(float))
    colmax = np.maximum(abs(Jo).max(axis=0).toarray().ravel(), 1e-300)
    err = abs(Jg - Jo).tocsc().max(axis=0).toarray().ravel() / colmax
    assert err.max() <= 1e-9


def test_linear_spline_degree_one():
    """traj_to_spline falls back to k=1 when a cubic fit throws (common.py:266-267)."""
    fl, truth, bakw = cases.make('gs_plain')
    t = fl.spline['tck'][0]
    kn = np.linspace(t[0][0], t[0][-1], 12)
    knots = np.concatenate(([kn[0]], kn, [kn[-1]]))
    rng = np.random.default_rng(0)
    from mvus_b200 import synth
    c = synth.gt_trajectory(kn) + rng.normal(size=(3, len(kn))) * 0.01
    fl.spline['tck'][0] = [knots, [c[0], c[1], c[2]], 1]
    fp = FlatProblem(fl, fl.numCam, **bakw)
    prob = ba_oracle.Problem(fl, fl.numCam, **bakw)
    r, span, J, mbase, mJ = helpers.emul_resjac(fp, prob.x0)
    ro = prob.residual(prob.x0)
    assert np.abs(r - ro).max() <= 1e-9 * max(1.0, np.abs(ro).max())
    from scipy import interpolate
    tt = np.linspace(kn[0], kn[-1] - 1e-9, 50)
    ref = np.asarray(interpolate.splev(tt, fl.spline['tck'][0]))
    l, B = ba_oracle.bspline_basis(None, knots, 1, tt)
    mine = np.array([B[0] * c[a][l - 1] + B[1] * c[a][l] for a in range(3)])
    assert np.abs(mine - ref).max() <= 1e-12


def test_oracle_bspline_vs_fitpack():
    from scipy import interpolate
    fl, truth, bakw = cases.make('rs_F_gap')
    for tck in fl.spline['tck']:
        tt = np.linspace(tck[0][0], tck[0][-1] - 1e-9, 333)
        ref = np.asarray(interpolate.splev(tt, tck))
        dref = np.asarray(interpolate.splev(tt, tck, der=1))
        l, B, dB = ba_oracle.bspline_basis(None, tck[0], 3, tt, nder=1)
        mine = np.array([sum(B[q] * tck[1][a][l - 3 + q] for q in range(4)) for a in range(3)])
        dmine = np.array([sum(dB[q] * tck[1][a][l - 3 + q] for q in range(4)) for a in range(3)])
        assert np.abs(mine - ref).max() <= 1e-12 * np.abs(ref).max()
        assert np.abs(dmine - dref).max() <= 1e-10 * max(np.abs(dref).max(), 1.0)


def test_oracle_undistort_vs_opencv():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(1)
    K = np.array([[1400.0, 0, 960], [0, 1380.0, 540], [0, 0, 1]])
    d = np.array([-0.26, 0.07, -1e-4, 2e-4, -0.009])
    pts = rng.uniform([0, 0], [1920, 1080], size=(200, 2))
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d).reshape(-1, 2)
    xn, yn = ba_oracle.undistort5(pts[:, 0], pts[:, 1], [K[0, 0], K[1, 1], K[0, 2], K[1, 2]], d)
    assert np.abs(xn - ref[:, 0]).max() <= 1e-12 and np.abs(yn - ref[:, 1]).max() <= 1e-12
