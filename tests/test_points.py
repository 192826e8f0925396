"""Scene.BA(motion_prior=True), the discrete-trajectory mode (common.py:466-467, 527-550,
587-605, 631-634, 681-687).  CPU: oracle/points_oracle.py against the reference's own error_BA closure (golden
fixtures tests/golden/points_*.npz).  GPU: mvus_ba_points_eval / mvus_ba_solve_points / Scene.BA against the
oracle and the reference's shipped solve."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_golden_points as mg                      # noqa: E402  (case table + flight factory, no reference needed)
from oracle import points_oracle                     # noqa: E402

CASES = list(mg.CASES)


def _case(case):
    fl, kw = mg.make(case)
    gold = np.load(os.path.join(HERE, 'golden', case + '.npz'))
    return fl, kw, gold, points_oracle.PointsProblem(fl, fl.numCam, **kw)


def _to_lib(pp, x):
    return np.concatenate((x[:pp.n_other], np.ravel(x[pp.n_other:].reshape(-1, 3).T)))


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_the_reference_closure(case):
    fl, kw, gold, pp = _case(case)
    assert pp.G == gold['global_traj'].shape[1] and np.array_equal(pp.pt_frame, gold['global_traj'][2])
    assert np.array_equal(pp.x0, gold['x0'])
    for x, r in zip(gold['xs'], gold['rs']):
        assert np.abs(pp.residual(x) - r).max() <= 1e-9 * max(1.0, np.abs(r).max())


def test_oracle_jacobian_is_the_derivative():
    """complex-step Jacobian against central differences of the restated function (columns of every kind)."""
    fl, kw, gold, pp = _case('points_KE_calib')
    x = gold['xs'][3]
    cols = np.array([0, pp.nc, 2 * pp.nc, 3 * pp.nc + 2, pp.n_other + 7, pp.n_other + 3 * 100 + 1, pp.n - 1])
    J = pp.jacobian_cs(x, cols)
    for q, c in enumerate(cols):
        h = 1e-6 * max(1.0, abs(x[c]))
        e = np.zeros(pp.n); e[c] = h
        fd = (pp.residual(x + e, reference_order=False) - pp.residual(x - e, reference_order=False)) / (2 * h)
        assert np.abs(J[:, q] - fd).max() <= 1e-4 * max(1.0, np.abs(fd).max())


# ---------------------------------------------------------------------------------------------- GPU
def _handles(fl, kw, pp):
    from mvus_b200 import _cabi, ba, points
    from mvus_b200.problem import FlatProblem
    fps = FlatProblem(fl, fl.numCam, rs=kw.get('rs', False), motion_reg=False)
    hs = _cabi.Handle(fps)
    ba._all_detect_to_traj(fl, fps, hs, fps.x0)
    slot, frame, yh = points.point_meta(fl, fps)
    assert np.array_equal(slot, pp.pt_cam) and np.array_equal(frame, pp.pt_frame) and np.abs(yh - pp.pt_yH).max() < 1e-15
    fpp = FlatProblem(points._PointScene(fl, pp.G), fl.numCam, rs=kw.get('rs', False), motion_reg=False)
    hp = _cabi.Handle(fpp, max_nfev=10)
    hp.points_set(slot, frame, yh)
    return fps, hs, hp


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_device_residual_and_gradient_match_oracle(case, built_lib):
    fl, kw, gold, pp = _case(case)
    fps, hs, hp = _handles(fl, kw, pp)
    mt = 1 if pp.motion_type == 'F' else 2
    try:
        free = pp.free_mask()
        for k, x in enumerate(gold['xs']):
            r, g, cost = hp.points_eval(hs, mt, pp.w, fps.x0, _to_lib(pp, x), want_g=(k in (0, 3)))
            ro = pp.residual(x, reference_order=False)
            assert np.abs(r - ro).max() <= 1e-9 * max(1.0, np.abs(ro).max()), (case, k)
            assert abs(cost - 0.5 * ro @ ro) <= 1e-10 * cost
            if g is not None:
                go = (pp.jacobian_cs(x).T @ ro) * free
                gl = _to_lib(pp, go)
                assert np.abs(g - gl).max() <= 1e-8 * np.abs(gl).max(), (case, k, np.abs(g - gl).max(), np.abs(gl).max())
    finally:
        hp.close()
        hs.close()


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_scene_ba_motion_prior_mode(case, built_lib, capsys):
    """Scene.BA(numCam, motion_prior=True, ...) on the mirror Scene: cost not above the reference's own
    10-evaluation solve, reported cost / fun = the oracle's at the returned x (reference layout and row
    placement), post-conditions of common.py:672-695."""
    fl, kw, gold, pp = _case(case)
    spline_before = [np.array(t[1]) for t in fl.spline['tck']]
    res = fl.BA(fl.numCam, max_iter=10, motion_prior=True, **kw)
    out = capsys.readouterr().out
    assert 'Number of BA parameters is %d' % pp.n in out and 'Doing BA with' in out
    assert res.x.shape == (pp.n,) and res.fun.shape == (pp.m,) and res.nfev <= 10
    # against the reference's own 10-evaluation solve: on the least-force flights SciPy's TRF does not move at all
    # (cost = cost0, finite differences on the truncated pattern) and the GPU solve ends far below; on the
    # kinetic-energy flight TRF's 2-D subspace steps are the better match for the quartic valley and reach 263
    # where the LM / trust-region driver is at 505 after the same 10 evaluations -- hence a one-sided
    # bound where the reference stalls and a progress bound (>= 98 % of the way) where it does not
    cost0 = pp.cost(pp.x0)
    if gold['shipped_cost'] > 0.5 * cost0:
        assert res.cost <= gold['shipped_cost'] * (1 + 1e-9), (res.cost, float(gold['shipped_cost']))
        assert res.cost < 0.05 * cost0
    else:
        assert cost0 - res.cost >= 0.98 * (cost0 - gold['shipped_cost']), (res.cost, float(gold['shipped_cost']))
    ro = pp.residual(res.x)
    assert abs(0.5 * ro @ ro - res.cost) <= 1e-9 * res.cost
    assert np.abs(res.fun - ro).max() <= 1e-8 * max(1.0, np.abs(ro).max())
    # post-conditions
    G = pp.G
    assert fl.global_traj.shape == (7, G)
    assert np.array_equal(fl.global_traj[4:7], res.x[pp.n_other:].reshape(G, 3).T)
    assert np.abs(fl.global_traj[3] - pp.timestamps(res.x)).max() <= 1e-9
    nc = pp.nc
    assert np.array_equal(np.asarray(fl.alpha)[fl.sequence[:nc]], res.x[:nc])
    assert len(fl.spline['tck']) == fl.spline['int'].shape[1] >= 1
    assert len(fl.detections_global) == fl.numCam and fl.visible is not None
    for t in fl.spline['tck']:
        assert len(t[1]) == 3 and len(t[1][0]) == len(t[0]) - t[2] - 1
    del spline_before


@pytest.mark.gpu
def test_motion_prior_mode_through_the_dropin(built_lib):
    """The same call on the REFERENCE's own Scene with mvus_b200.dropin.install: same solution as on the mirror."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip('reference tree not present')
    from mvus_b200 import dropin
    fl, kw, gold, pp = _case('points_KE_calib')
    fl2, _ = mg.make('points_KE_calib')
    common = ref_shim.load()
    ref = ref_shim.to_reference_scene(fl2)
    dropin.install(common)
    try:
        res_ref = ref.BA(fl2.numCam, max_iter=10, motion_prior=True, **kw)
    finally:
        dropin.uninstall(common)
    res = fl.BA(fl.numCam, max_iter=10, motion_prior=True, **kw)
    assert abs(res.cost - res_ref.cost) <= 1e-9 * res.cost and res.nfev == res_ref.nfev
    assert np.abs(res.x - res_ref.x).max() <= 1e-9 * np.abs(res.x).max()
    assert np.allclose(np.asarray(ref.spline['int']), np.asarray(fl.spline['int']), rtol=1e-9, atol=1e-9)


def test_reference_placement_matches_oracle():
    """mvus_b200.points.reference_placement (host logic of res.fun's row order) against the oracle's restatement of
    the reference's np.intersect1d scatter, on time stamps with crossings."""
    from mvus_b200 import points
    fl, kw, gold, pp = _case('points_F')
    x = gold['xs'][2]                                   # perturbed alpha / beta: some time stamps have crossed
    ts = pp.timestamps(x)
    assert (np.diff(ts) <= 0).any()
    gid, prev, nxt = pp.neighbours(ts)
    rows = np.nonzero((gid > 0) & (prev >= 0) & (nxt >= 0))[0]
    assert np.array_equal(points.reference_placement(ts, rows, gid), pp.placement(ts, rows, gid))
