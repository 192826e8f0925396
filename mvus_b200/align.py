"""Ground-truth alignment of a reconstructed flight (analysis/compare_gt.py:73-151) with the fits on the GPU.

``align_gt(flight, f_gt, gt_path, visualize=False)`` has the reference's signature, prints and return value
(``{'align_param', 'reconst_tran', 'gt', 'tran_matrix', 'error'}``, the dict ``main.py:85-94`` stores in
``flight.out`` before pickling the Scene).  What runs where:

* every similarity fit (thirdparty/transformation.py:869-975) with its spline evaluations and point
  distances -- the coarse search over all integer time shifts in ONE launch, each residual evaluation of the
  fine stage in one launch -- is CUDA (csrc/align.cuh through ``mvus_ba_align``);
* the interpolating spline of the ground truth that ``util.match_overlap`` (util.py:119-135) refits for every
  shift is fitted once, on the device (``splfit.splprep(s=0)``);
* the 2-parameter robust least-squares driver of the fine stage is SciPy's ``least_squares`` exactly as the
  reference calls it (compare_gt.py:66: loss='cauchy', f_scale=1); only its residual function changed.

There is no CPU fallback: without the CUDA library ``_cabi`` raises.
"""
import numpy as np

from . import _cabi, ba, splfit
from .problem import FlatProblem


class _SplineScene:
    """Just enough of a Scene for ba._spline_only_problem: a spline dictionary."""

    def __init__(self, tck, interval):
        self.spline = {'tck': tck, 'int': np.asarray(interval, dtype=np.float64)}


def _spline_handle(scene):
    fp = ba._spline_only_problem(scene)
    return _cabi.Handle(fp, device=ba.DEVICE), fp


def _orient_gt(gt_ori):
    if gt_ori.shape[0] == 3 or gt_ori.shape[0] == 4:
        pass
    elif gt_ori.shape[1] == 3 or gt_ori.shape[1] == 4:
        gt_ori = gt_ori.T
    else:
        raise Exception('Ground truth data have an invalid shape')
    return gt_ori


def coarse_search(reconst, gt):
    """compare_gt.py:105-126: the integer shift j of the reconstruction's (GT-rate) time axis that gives the
    smallest mean distance after a similarity fit onto the ground truth; all shifts in one launch.
    -> (j, mean error per shift, shifts)."""
    thres = int(reconst[0, -1] / 2)
    if int(gt[0, -1] - thres) < 0:
        raise Exception('Ground truth too short!')
    shifts = np.arange(-thres, int(gt[0, -1] - thres), dtype=np.float64)
    # util.match_overlap: cubic interpolating spline through the ground truth, one per continuous part is NOT
    # what the reference does -- it fits ONE spline through all samples and only samples inside the parts
    tck, _ = splfit.splprep(gt[1:], gt[0], 0, k=3)
    interval = splfit.find_intervals(gt[0])
    hd, fp = _spline_handle(_SplineScene([tck], interval[:, :1]))
    try:
        # one spline, several membership intervals: the handle's intervals decide membership AND pick the
        # spline, so give it one spline per interval (same knots and coefficients)
        if interval.shape[1] > 1:
            hd.close()
            hd, fp = _spline_handle(_SplineScene([[tck[0], [c.copy() for c in tck[1]], tck[2]]
                                                   for _ in range(interval.shape[1])], interval))
        mean_err, count, M, _ = hd.align(fp.x0, reconst[0], reconst[1:], shifts, spline_is_src=False)
    finally:
        hd.close()
    if (count < 3).any():
        raise ValueError('input arrays are of wrong shape or type')     # transformation.py:915
    return int(shifts[int(np.argmin(mean_err))]), mean_err, shifts


def optimize(alpha, beta, flight, gt):
    """compare_gt.py:35-70: refine (alpha, beta) of t = alpha * t_gt + beta with a Cauchy-loss least-squares
    fit of the distances between the similarity-transformed reconstruction and the ground truth."""
    from scipy.optimize import least_squares
    hd, fp = _spline_handle(flight)
    try:
        def t_of(model):
            a, b = model[0], model[1]
            if gt.shape[0] == 3:
                return a * np.arange(gt.shape[1]) + b
            return a * (gt[0] - gt[0, 0]) + b

        def error_fn(model):
            _, _, _, err = hd.align(fp.x0, t_of(model), gt[-3:], [0.0], spline_is_src=True, want=0)
            return err

        ls = least_squares(error_fn, np.array([alpha, beta], dtype=np.float64), loss='cauchy', f_scale=1)
        t_gt = t_of(ls.x)
        mean_err, count, M, err = hd.align(fp.x0, t_gt, gt[-3:], [0.0], spline_is_src=True, want=0)
        if count[0] < 3:
            raise ValueError('input arrays are of wrong shape or type')
        _, idx = _sampling_mask(t_gt, flight.spline['int'])
        traj = hd.spline_to_traj(fp.x0, t_gt[idx])
        flight.traj = traj                     # the reference leaves its last spline_to_traj(t=...) here (common.py:299)
    finally:
        hd.close()
    M = M[0]
    tran = M @ np.vstack((traj[1:], np.ones(traj.shape[1])))
    tran /= tran[-1]
    return ls, (np.vstack((traj[0], tran[:3])), gt[-3:, idx], M, err[idx])


def _sampling_mask(t, interval):
    """util.sampling(t, interval) (util.py:90-116): members of any interval, a <= t < b."""
    interval = np.asarray(interval)
    idx = np.zeros(len(t), dtype=bool)
    for i in range(interval.shape[1]):
        idx |= np.logical_xor(t - interval[0, i] >= 0, t - interval[1, i] >= 0)
    return t[idx], idx


def align_gt(flight, f_gt, gt_path, visualize=False):
    if not len(gt_path):
        print('No ground truth data provided\n')
        return
    try:
        gt_ori = np.loadtxt(gt_path)
    except Exception:
        print('Ground truth not correctly loaded')
        return
    gt_ori = _orient_gt(gt_ori)

    # Pre-processing (compare_gt.py:95-105)
    f_reconst = flight.cameras[flight.settings['ref_cam']].fps
    alpha = f_reconst / f_gt
    reconst = ba.spline_to_traj(flight, sampling_rate=alpha)
    t0 = reconst[0, 0]
    reconst = np.vstack(((reconst[0] - t0) / alpha, reconst[1:]))
    if gt_ori.shape[0] == 3:
        gt = np.vstack((np.arange(len(gt_ori[0])), gt_ori))
    else:
        gt = np.vstack((gt_ori[0] - gt_ori[0, 0], gt_ori[1:]))

    j, _, _ = coarse_search(reconst, gt)
    beta = t0 - alpha * j

    ls, res = optimize(alpha, beta, flight, gt_ori)

    # Remove outliers by relative thresholding (compare_gt.py:131-135)
    thres = 10
    error_ = res[3]
    idx = error_ <= thres * np.mean(error_)
    reconst_, gt_, error_ = res[0][:, idx], res[1][:, idx], error_[idx]

    out = {'align_param': ls.x, 'reconst_tran': reconst_, 'gt': gt_, 'tran_matrix': res[2], 'error': error_}
    print('The mean error (distance) is {:.5f} meter\n'.format(np.mean(out['error'])))
    print('The median error (distance) is {:.5f} meter\n'.format(np.median(out['error'])))
    print(ls.x)
    if visualize:                              # plots stay with the reference (tools/visualization.py, matplotlib)
        import tools.visualization as vis
        vis.show_trajectory_3D(out['reconst_tran'][1:], out['gt'], line=False,
                               title='Reconstruction(left) vs Ground Truth(right)')
        vis.error_hist(out['error'])
        vis.error_traj(out['reconst_tran'][1:], out['error'])
    return out
