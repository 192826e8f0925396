"""CPU restatement of the ground-truth alignment (analysis/compare_gt.py:12-151) -- TEST INFRASTRUCTURE.

Only tests/ import this.  It restates, in NumPy/SciPy, the pieces the CUDA path replaces:
  similarity()      thirdparty/transformation.py:869-975 affine_matrix_from_points(shear=False, scale=True,
                    usesvd=True): Kabsch rotation from the SVD of v1 v0^T with the reflection fix, scale from
                    the RMS deviations, centroids moved back
  match_overlap()   tools/util.py:119-135
  coarse_errors()   compare_gt.py:112-126 (mean distance per integer shift)
  fine_error()      compare_gt.py:37-52 (error_fn of optimize)
Pinned against the live reference (oracle/_ref or /root/reference) in tests/test_align.py; "parity unpinned"
otherwise does not apply: the reference's own functions are importable wherever the tests run.
"""
import numpy as np
from scipy import interpolate


def similarity(v0, v1):
    v0 = np.array(v0, dtype=np.float64, copy=True)
    v1 = np.array(v1, dtype=np.float64, copy=True)
    if v0.shape[1] < 3 or v0.shape != v1.shape:
        raise ValueError('input arrays are of wrong shape or type')
    c0, c1 = v0.mean(axis=1), v1.mean(axis=1)
    v0 -= c0[:, None]
    v1 -= c1[:, None]
    u, s, vh = np.linalg.svd(v1 @ v0.T)
    R = u @ vh
    if np.linalg.det(R) < 0.0:
        R -= np.outer(u[:, 2], vh[2, :] * 2.0)
    sc = np.sqrt(np.sum(v1 * v1) / np.sum(v0 * v0))
    M = np.eye(4)
    M[:3, :3] = sc * R
    M[:3, 3] = c1 - sc * R @ c0
    return M


def distances(M, src, dst):
    tran = M @ np.vstack((src, np.ones(src.shape[1])))
    tran /= tran[-1]
    return np.sqrt(((dst - tran[:3]) ** 2).sum(axis=0))


def find_intervals(x, gap=5):
    x_s, x_e = np.append(-np.inf, x), np.append(x, np.inf)
    start = x_s[1:] - x_s[:-1] >= gap
    end = x_e[:-1] - x_e[1:] <= -gap
    interval = np.array([x[start], x[end]])
    return interval[:, interval[1] - interval[0] >= gap]


def members(t, interval):
    idx = np.zeros(len(t), dtype=bool)
    for i in range(interval.shape[1]):
        idx |= np.logical_xor(t - interval[0, i] >= 0, t - interval[1, i] >= 0)
    return idx


def match_overlap(x, y):
    interval = find_intervals(y[0])
    x_s = x[:, members(x[0], interval)]
    tck, _ = interpolate.splprep(y[1:], u=y[0], s=0, k=3)
    y_s = np.vstack((x_s[0], np.asarray(interpolate.splev(x_s[0], tck))))
    return x_s, y_s


def coarse_errors(reconst, gt):
    thres = int(reconst[0, -1] / 2)
    shifts = np.arange(-thres, int(gt[0, -1] - thres))
    out = np.empty(len(shifts))
    for k, i in enumerate(shifts):
        p1, p2 = match_overlap(np.vstack((reconst[0] + i, reconst[1:])), gt)
        M = similarity(p1[1:], p2[1:])
        out[k] = distances(M, p1[1:], p2[1:]).mean()
    return shifts, out


def fine_error(model, gt, tck_list, interval):
    """error_fn(model) of compare_gt.optimize for a flight with splines tck_list / interval."""
    a, b = model
    t_gt = a * np.arange(gt.shape[1]) + b if gt.shape[0] == 3 else a * (gt[0] - gt[0, 0]) + b
    idx = members(t_gt, interval)
    pts = np.empty((3, int(idx.sum())))
    t_part = t_gt[idx]
    for i in range(interval.shape[1]):
        m = (t_part >= interval[0, i]) & (t_part <= interval[1, i])
        if m.any():
            pts[:, m] = np.asarray(interpolate.splev(t_part[m], tck_list[i]))
    M = similarity(pts, gt[-3:, idx])
    err = np.zeros(gt.shape[1])
    err[idx] = distances(M, pts, gt[-3:, idx])
    return err, M


def spline_to_traj(tck_list, interval, sampling_rate=1, t=None):
    """Scene.spline_to_traj (common.py:273-301) with SciPy: unit/constant-rate samples (or the given ascending
    times) that lie inside an interval (closed ends), evaluated on that interval's spline."""
    ts = np.arange(interval[0, 0], interval[1, -1], sampling_rate) if t is None else np.asarray(t, dtype=np.float64)
    out = np.empty((4, 0))
    for i in range(interval.shape[1]):
        part = ts[(ts >= interval[0, i]) & (ts <= interval[1, i])]
        out = np.hstack((out, np.vstack((part, np.asarray(interpolate.splev(part, tck_list[i]))))))
    return out


def preprocess(tck_list, interval, fps_ref, f_gt, gt_ori):
    """compare_gt.py:95-105 -> (alpha, t0, reconst in GT-sample time, gt with a time row starting at 0)."""
    alpha = fps_ref / f_gt
    reconst = spline_to_traj(tck_list, interval, sampling_rate=alpha)
    t0 = reconst[0, 0]
    reconst = np.vstack(((reconst[0] - t0) / alpha, reconst[1:]))
    gt = np.vstack((np.arange(len(gt_ori[0])), gt_ori)) if gt_ori.shape[0] == 3 else \
        np.vstack((gt_ori[0] - gt_ori[0, 0], gt_ori[1:]))
    return alpha, t0, reconst, gt
