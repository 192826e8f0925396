"""Spline fit (Scene.traj_to_spline, common.py:224-270 = scipy.interpolate.splprep = FITPACK parcur):
  * the oracle (oracle/fitpack_oracle.py) is pinned against the INSTALLED splprep: identical knots,
    coefficients <= 1e-8 relative, same ier;
  * the product's host loop (mvus_b200/splfit.py) is run here with a NumPy stand-in for the device solves
    (test infrastructure; the product itself only has the CUDA solves) and must agree with splprep too;
  * find_intervals against the reference's own util.find_intervals when the reference is present.
The device solves themselves are compared in tests/test_gpu_reference.py (-m gpu)."""
import numpy as np
import pytest
from scipy import interpolate

from mvus_b200 import splfit, synth
from oracle import fitpack_oracle as fo, ref_shim

CASES = [  # (m, noise, s factor relative to m * noise^2 (or absolute if noise == 0), degree, jitter)
    (300, 0.01, 1.0, 3, True), (300, 0.01, 0.5, 3, True), (2000, 0.02, 1.0, 3, True), (1500, 0.0, 6e-4, 3, False),
    (500, 0.05, 4.0, 3, True), (500, 0.05, 4e8, 3, True), (64, 0.0, 0.0, 3, True), (400, 0.02, 3.0, 1, True),
    (50, 0.02, 0.05, 1, True), (5, 0.0, 1e-6, 3, False), (4, 0.0, 1e-6, 3, False)]


def _data(m, noise, sf, seed=0, jitter=True):
    rng = np.random.default_rng(seed)
    u = np.sort(rng.uniform(0, 600.0, m)) if jitter else np.linspace(0.0, 600.0, m)
    x = synth.gt_trajectory(u) + rng.normal(size=(3, m)) * noise
    s = sf * m * noise ** 2 if noise > 0 else sf
    return u, x, s


def _check(t, c, ier, tck, ier_ref, tol=1e-8):
    assert len(t) == len(tck[0]) and np.array_equal(t, tck[0])
    scale = max(np.abs(np.asarray(tck[1])).max(), 1.0)
    assert max(np.abs(c[d][:len(tck[1][d])] - tck[1][d]).max() for d in range(3)) <= tol * scale
    assert ier == ier_ref


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_installed_splprep(case):
    m, noise, sf, k, jitter = case
    u, x, s = _data(m, noise, sf, jitter=jitter)
    (tck, _), fp, ier, msg = interpolate.splprep(x, u=u, s=s, k=k, full_output=1)
    t, c, fpo, iero = fo.parcur_fit(u, x, s, k)
    _check(t, c, iero, tck, ier)
    assert abs(fpo - fp) <= 1e-6 * max(fp, 1e-12) + 1e-18


class _NumpySolves:
    """Stand-in for _cabi.SplHandle (what spl_fit.cuh computes), built from the oracle's NumPy pieces."""

    def __init__(self, u, x, k=3, device=0):
        if len(u) <= k:
            raise TypeError('m > k must hold')
        self.u, self.x, self.k = u, x, k

    def solve(self, t, pen=None, pscale=0.0):
        from scipy.linalg import cholesky_banded, cho_solve_banded
        k = self.k
        G, rhs, l, h = fo.normal_equations(t, k, self.u, self.x)
        nk1 = len(t) - k - 1
        A = np.zeros((k + 2, nk1))
        A[1:] = G
        if pen is not None:
            A += pscale * pen[::-1]                  # pen[d][j] = P[j-d][j] -> scipy's upper band rows
        cb = cholesky_banded(A, lower=False)
        c = np.array([cho_solve_banded((cb, False), r) for r in rhs])
        fp, fpint, _ = fo.residual_sums(t, k, self.u, self.x, c, l, h)
        return c, fp, fpint, float(np.sum(cb[-1]))

    def close(self):
        pass


@pytest.mark.parametrize('case', CASES)
def test_product_host_loop_matches_installed_splprep(case, monkeypatch):
    from mvus_b200 import _cabi
    monkeypatch.setattr(_cabi, 'SplHandle', _NumpySolves)
    m, noise, sf, k, jitter = case
    u, x, s = _data(m, noise, sf, jitter=jitter)
    (tck, _), fp, ier, msg = interpolate.splprep(x, u=u, s=s, k=k, full_output=1)
    t, c, fpo, iero = splfit.fit(u, x, s, k)
    _check(t, c, iero, tck, ier)
    tck2, u2 = splfit.splprep(x, u, s, k=k)
    assert tck2[2] == k and len(tck2[1]) == 3 and all(len(a) == len(tck[1][0]) for a in tck2[1])


def test_cubic_fit_with_too_few_points_raises_like_splprep(monkeypatch):
    from mvus_b200 import _cabi
    monkeypatch.setattr(_cabi, 'SplHandle', _NumpySolves)
    u = np.arange(3.0)
    x = np.zeros((3, 3))
    with pytest.raises(TypeError, match='m > k must hold'):
        interpolate.splprep(x, u=u, s=1e-6, k=3)
    with pytest.raises(TypeError, match='m > k must hold'):
        splfit.splprep(x, u, 1e-6, k=3)


def test_helpers_match_the_oracle():
    u, x, s = _data(800, 0.02, 1.0)
    t, c, fp, ier = fo.parcur_fit(u, x, s, 3)
    assert np.array_equal(splfit._count_inside(t, 3, u), fo.count_data(t, 3, u))
    b = fo.disc_jumps(t, 3)
    nk1 = len(t) - 4
    BtB = np.zeros((nk1, nk1))
    for r in range(b.shape[0]):
        BtB[r:r + 5, r:r + 5] += np.outer(b[r], b[r])
    pen = splfit._jump_penalty(t, 3)
    for d in range(5):
        assert np.abs(pen[d][d:] - np.diag(BtB, d)).max() <= 1e-12 * np.abs(BtB).max()
    assert np.array_equal(splfit._interpolation_knots(u, 3), fo.interpolation_knots(u, 3))
    assert np.array_equal(splfit._interpolation_knots(u, 1), fo.interpolation_knots(u, 1))


def test_find_intervals_matches_the_reference():
    if not ref_shim.available():
        pytest.skip('reference not present')
    ref_shim.load()
    from tools import util
    rng = np.random.default_rng(0)
    x = np.cumsum(rng.choice([0.5, 1.0, 1.0, 1.0, 7.0, 30.0], size=400, p=[.2, .3, .2, .2, .07, .03]))
    a, ai = util.find_intervals(x, idx=True)
    b, bi = splfit.find_intervals(x, idx=True)
    assert np.array_equal(a, b) and np.array_equal(ai, bi)


def test_traj_to_spline_matches_the_reference(monkeypatch):
    """Scene.traj_to_spline through the product's code (NumPy stand-in for the device) against the
    reference's own method on the same discrete trajectory: intervals, knots, degree, coefficients."""
    if not ref_shim.available():
        pytest.skip('reference not present')
    from mvus_b200 import _cabi
    from mvus_b200.scene import Scene
    monkeypatch.setattr(_cabi, 'SplHandle', _NumpySolves)
    common = ref_shim.load()
    rng = np.random.default_rng(3)
    tt = np.concatenate((np.arange(0.0, 400.0, 0.5), np.arange(420.0, 700.0, 0.5), np.arange(720.0, 721.5, 0.5),
                         np.arange(800.0, 806.0, 0.5)))
    traj = np.vstack((tt, synth.gt_trajectory(tt) + rng.normal(size=(3, len(tt))) * 0.01))
    ref, mine = common.Scene(), Scene()
    ref.traj, mine.traj = traj.copy(), traj.copy()
    ref.traj_to_spline(smooth_factor=[10, 20])
    mine.traj_to_spline(smooth_factor=[10, 20])
    assert np.array_equal(ref.spline['int'], mine.spline['int'])
    assert len(ref.spline['tck']) == len(mine.spline['tck']) >= 3
    for a, b in zip(ref.spline['tck'], mine.spline['tck']):
        assert a[2] == b[2] and np.array_equal(a[0], b[0])
        for d in range(3):
            assert np.abs(a[1][d] - b[1][d]).max() <= 1e-8 * max(1.0, np.abs(a[1][d]).max())
