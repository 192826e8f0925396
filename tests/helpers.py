"""Shared test helpers: compact block-row Jacobian (include/mvus_ba.h layout) -> CSR in the
reference's (row, column) numbering, so it can be compared with the oracle's Jacobian."""
import ctypes
import os

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def expand_jacobian(fp, span, J, mbase=None, mJ=None):
    """fp: mvus_b200.problem.FlatProblem.  Returns CSR (m x n)."""
    N, P, nc, C, Pc = fp.N, fp.P, fp.nc, fp.C, fp.Pc
    M = 0 if mbase is None else len(mbase)
    m = 2 * N + M
    J = np.asarray(J).reshape(2 * P, N)
    cam_of = np.repeat(np.arange(nc), fp.N_cam)
    local = np.arange(N) - fp.cam_ptr[cam_of]
    row_u = 2 * fp.cam_ptr[cam_of] + local
    row_v = row_u + fp.N_cam[cam_of]
    rows, cols, vals = [], [], []
    cov = span >= 0
    # spline of each covered detection and local span l
    spl = np.searchsorted(fp.ctrl_off, np.where(cov, span, 0), side='right') - 1
    l = np.where(cov, span, 0) - fp.ctrl_off[spl]
    for p in range(P):
        if p == 0:
            col = cam_of
        elif p == 1:
            col = nc + cam_of
        elif p == 2:
            col = 2 * nc + cam_of
        elif p < Pc:
            col = 3 * nc + cam_of * C + (p - 3)
        else:
            mslot, ax = divmod(p - Pc, 3)
            j = l - 3 + mslot
            ok = cov & (j >= 0)
            col = fp.n_other + 3 * fp.ctrl_off[spl] + ax * fp.ncoef[spl] + j
            for rr, plane in ((row_u, J[p]), (row_v, J[P + p])):
                rows.append(rr[ok]); cols.append(col[ok]); vals.append(plane[ok])
            continue
        for rr, plane in ((row_u, J[p]), (row_v, J[P + p])):
            rows.append(rr[cov]); cols.append(col[cov]); vals.append(plane[cov])
    if M:
        mJ = np.asarray(mJ).reshape(10, M)
        act = np.nonzero(mbase >= 0)[0]
        b = mbase[act]
        spl = np.searchsorted(fp.ctrl_off, b, side='right') - 1
        for k in range(7):
            j = b + k - fp.ctrl_off[spl]
            ok = j < fp.ncoef[spl]
            for ax in range(3):
                col = fp.n_other + 3 * fp.ctrl_off[spl] + ax * fp.ncoef[spl] + j
                rows.append(2 * N + act[ok]); cols.append(col[ok]); vals.append((mJ[ax, act] * mJ[3 + k, act])[ok])
    rows = np.concatenate(rows); cols = np.concatenate(cols); vals = np.concatenate(vals)
    A = sp.coo_matrix((vals, (rows, cols)), shape=(m, fp.n)).tocsr()
    A.sum_duplicates()
    return A


def build_emul():
    """Build tests/emul/libmvus_emul.so (host compile of the kernels' math) if needed."""
    import subprocess
    d = os.path.join(ROOT, 'tests', 'emul')
    so = os.path.join(d, 'libmvus_emul.so')
    srcs = [os.path.join(d, 'emul.cpp'), os.path.join(ROOT, 'mvus_b200', 'csrc', 'ba_math.cuh'),
            os.path.join(ROOT, 'mvus_b200', 'csrc', 'ba_tables.hpp')]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-Wno-unknown-pragmas', '-o', so, srcs[0]])
    return ctypes.CDLL(so)


def emul_resjac(fp, x):
    """Run the host emulation of K1/K1m on FlatProblem fp at x -> (r, span, J, mbase, mJ)."""
    lib = build_emul()
    dp = ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(ctypes.c_int32)
    lp = ctypes.POINTER(ctypes.c_int64)

    def D(a):
        return a.ctypes.data_as(dp)
    lib.emul_motion_count.restype = ctypes.c_int64
    interval = np.ascontiguousarray(fp.interval.reshape(-1))
    M = 0
    if fp.motion_type:
        M = lib.emul_motion_count(fp.S, D(interval), fp.knot_ptr.ctypes.data_as(lp), D(fp.knots),
                                  fp.degree.ctypes.data_as(ip))
    r = np.zeros(2 * fp.N + M)
    span = np.zeros(fp.N, dtype=np.int32)
    J = np.zeros(2 * fp.P * fp.N)
    mbase = np.zeros(M, dtype=np.int32)
    mJ = np.zeros(10 * M)
    x = np.ascontiguousarray(x, dtype=np.float64)
    rc = lib.emul_resjac(fp.nc, int(fp.opt_calib), int(fp.undist), int(fp.opt_sync), int(fp.opt_rs),
                         fp.motion_type, ctypes.c_double(fp.motion_weight), fp.cam_ptr.ctypes.data_as(lp),
                         D(fp.frame), D(fp.x_raw), D(fp.y_raw), D(fp.height), D(fp.calib), fp.S,
                         D(interval), fp.knot_ptr.ctypes.data_as(lp), D(fp.knots),
                         fp.degree.ctypes.data_as(ip), D(x), D(r), span.ctypes.data_as(ip), D(J),
                         mbase.ctypes.data_as(ip), D(mJ))
    assert rc == 0, rc
    return r, span, J, mbase, mJ


def numpy_all_detect_to_traj(scene, cams):
    """NumPy restatement of Scene.all_detect_to_traj (common.py:887-944) on a scene whose
    detections_global is current; returns global_traj (7 x n)."""
    from scipy import interpolate
    ts = np.concatenate([scene.detections_global[i][0] for i in cams])
    fid = np.concatenate([np.asarray(scene.detections[i][0], dtype=np.float64) for i in cams])
    cid = np.concatenate([np.ones(scene.detections[i].shape[1]) * i for i in cams])
    tck, interval = scene.spline['tck'], np.asarray(scene.spline['int'])
    tsort = np.sort(ts)
    traj = np.empty([4, 0])
    for i in range(interval.shape[1]):
        part = tsort[np.logical_and(tsort >= interval[0, i], tsort <= interval[1, i])]
        traj = np.hstack((traj, np.vstack((part, np.asarray(interpolate.splev(part, tck[i]))))))
    gd = np.vstack((cid, fid, ts))
    tmp = gd[:, np.argsort(gd[2, :], kind='stable')]
    keep = np.isin(tmp[2], traj[0])
    tmp = np.vstack((tmp[:, keep], traj[1:]))
    return np.vstack((np.arange(tmp.shape[1]), tmp))


def check_normal_equations(fp, prob, x, A, g, Hss, Hcs, cost, rtol=1e-9, n_sample=3000):
    """Compare K2's output (mvus_ba_normal_equations layout) with J^T J / J^T r formed from the
    ORACLE's Jacobian, sparse all the way so that it works at the BASELINE sizes: camera blocks, the
    whole gradient, the camera x control-point coupling and sampled control-point band blocks."""
    ro = prob.residual(x)
    Jo = (prob.jacobian(x).tocsc() @ sp.diags(prob.free_mask().astype(float))).tocsc()
    assert abs(cost - 0.5 * ro @ ro) <= 1e-12 * cost
    go = Jo.T @ ro
    assert np.abs(g - go).max() <= rtol * np.abs(go).max()
    nc, C = fp.nc, fp.C
    cam_cols = [np.array([i, nc + i, 2 * nc + i] + list(range(3 * nc + i * C, 3 * nc + (i + 1) * C)))
                for i in range(nc)]
    for i in range(nc):
        Jc = Jo[:, cam_cols[i]]
        blk = (Jc.T @ Jc).toarray()
        assert np.abs(A[i] - blk).max() <= rtol * max(np.abs(blk).max(), 1e-300), i
    # control-point columns in control-point-major order (the solver's numbering)
    cols = np.concatenate([(fp.n_other + 3 * fp.ctrl_off[s] + np.arange(3)[None, :] * fp.ncoef[s]
                            + np.arange(fp.ncoef[s])[:, None]).ravel() for s in range(fp.S)])
    Js = Jo[:, cols]
    Hc = (Jo[:, np.concatenate(cam_cols)].T @ Js).toarray()
    assert Hcs.shape == Hc.shape
    assert np.abs(Hcs - Hc).max() <= rtol * np.abs(Hc).max()
    band = Hss.shape[1]
    rng = np.random.default_rng(1)
    pick = np.unique(np.concatenate(([0, 1, 2, fp.n_ctrl - 1, fp.n_ctrl - 2],
                                     rng.integers(0, fp.n_ctrl, n_sample))))
    scale = abs(Js).max() ** 2
    for i in pick:
        hi = min(fp.n_ctrl, i + band)
        blk = (Js[:, 3 * i:3 * i + 3].T @ Js[:, 3 * i:3 * hi]).toarray()        # 3 x 3 (hi - i)
        mine = np.concatenate([Hss[i, dj] for dj in range(hi - i)], axis=1)
        assert np.abs(mine - blk).max() <= rtol * max(np.abs(blk).max(), 1e-6 * scale), int(i)
    # nothing outside the band: the rows further than `band` control points apart are orthogonal
    far = Js[:, :3].T @ Js[:, 3 * band:] if fp.n_ctrl > band else None
    assert far is None or abs(far).max() == 0.0


def make_alignment_case(name='rs_F_gap', det_per_cam=400, f_gt=5.0, seed=7, with_time=True):
    """A flight plus a synthetic RTK ground truth for analysis/compare_gt.align_gt: the true trajectory at
    f_gt Hz over a wider time span than the flight, similarity-transformed (scale, rotation, translation) with
    2 cm noise.  -> (flight, gt array 4 x n [time in GT samples' own clock; X; Y; Z] or 3 x n, f_gt)."""
    import cases
    from mvus_b200 import hostmath, synth
    fl, truth, _ = cases.make(name, det_per_cam=det_per_cam, perturb=0.0)
    fl.settings['ref_cam'] = fl.ref_cam
    rng = np.random.default_rng(seed)
    fps = fl.cameras[fl.ref_cam].fps
    interval = np.asarray(fl.spline['int'])
    t_lo, t_hi = interval[0, 0] - 6.0 * fps, interval[1, -1] + 9.0 * fps          # frames of the reference camera
    tau = np.arange(t_lo, t_hi, fps / f_gt)
    X = synth.gt_trajectory(tau, fps_ref=fps)
    R = hostmath.rodrigues_to_matrix(np.array([0.3, -0.2, 0.5]))
    Y = 1.7 * R @ X + np.array([[10.0], [-4.0], [2.5]]) + rng.normal(size=X.shape) * 0.02
    if with_time:
        return fl, np.vstack((100.0 + np.arange(len(tau)), Y)), f_gt
    return fl, Y, f_gt
