"""Synthetic unsynchronised multi-camera flights of the shapes named in BASELINE.json.

The reference ships no data; every test, fixture and benchmark in
this repository runs on flights generated here (seeds: flight 0, noise 1, perturbation 2,
SURVEY.md 8d).  Two products:

* ``make_flight``  -> a BA-ready ``Scene`` (ground truth + perturbation) of any size,
  built directly because the shipped pipeline cannot reach configs 2-4.
* ``write_dataset`` -> dataset4-format files (``x y frame`` rows, README.md:143-162,
  camera JSON README.md:164-184, config JSON config.json) for running the reference's
  own ``main.py`` end to end (config 1).

Camera/time model used for simulation = the one the reference optimises
(common.py:125-126, 1072-1079): t = alpha (f + rho y/H) + beta ; pinhole P X with the
observation being the *distorted* pixel that ``cv2.undistortPoints`` later undoes.
"""
import json
import os

import numpy as np

from . import hostmath
from .scene import Camera, Scene

FPS_CHOICES = (30.0, 25.0, 50.0, 59.94, 29.97, 60.0, 24.0)
README_DIST = np.array([-0.260720634999793, 0.07494782427852716, -0.00013631462898833923,
                        0.00017484761775924765, -0.00906247784302948])


def gt_trajectory(tau, fps_ref=30.0, deriv=False):
    """Ground-truth 3D position (3xn, metres) at global time ``tau`` (reference-camera
    frames).  Bounded helix-like curve, smooth at the scale of a knot span.  ``deriv=True``
    returns d2X/dtau2 instead."""
    s = np.asarray(tau, dtype=np.float64) / fps_ref
    if deriv:
        return np.vstack((-1.25 * np.cos(0.5 * s) - 0.0169 * np.sin(0.13 * s),
                          -1.25 * np.sin(0.5 * s) - 0.0289 * np.cos(0.17 * s),
                          -0.08 * np.sin(0.2 * s) - 0.000961 * np.sin(0.031 * s))) / fps_ref ** 2
    X = np.vstack((5.0 * np.cos(0.5 * s) + 1.0 * np.sin(0.13 * s),
                   5.0 * np.sin(0.5 * s) + 1.0 * np.cos(0.17 * s),
                   3.0 + 2.0 * np.sin(0.2 * s) + 1.0 * np.sin(0.031 * s)))
    return X


def distort_normalized(xn, yn, d):
    """Forward Brown model (k1,k2,p1,p2,k3) on normalised coordinates."""
    k1, k2, p1, p2, k3 = d
    r2 = xn * xn + yn * yn
    radial = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3))
    xd = xn * radial + 2.0 * p1 * xn * yn + p2 * (r2 + 2.0 * xn * xn)
    yd = yn * radial + p1 * (r2 + 2.0 * yn * yn) + 2.0 * p2 * xn * yn
    return xd, yd


def make_cameras(nc, seed=0, distortion=True, radius=25.0):
    """``nc`` cameras on a ring looking at the flight volume; intrinsics ~1400 px,
    1920x1080."""
    rng = np.random.default_rng(seed)
    cams = []
    for i in range(nc):
        ang = 2.0 * np.pi * i / nc + rng.uniform(-0.1, 0.1)
        c = np.array([radius * np.cos(ang), radius * np.sin(ang), rng.uniform(0.5, 6.0)])
        R = hostmath.look_at(c, np.array([0.0, 0.0, 3.0]) + rng.uniform(-0.5, 0.5, 3))
        K = np.array([[1400.0 + rng.uniform(-60, 60), 0.0, 960.0 + rng.uniform(-15, 15)],
                      [0.0, 1400.0 + rng.uniform(-60, 60), 540.0 + rng.uniform(-15, 15)],
                      [0.0, 0.0, 1.0]])
        d = README_DIST * rng.uniform(0.7, 1.1, 5) if distortion else np.zeros(5)
        fps = FPS_CHOICES[i % len(FPS_CHOICES)] if i else 30.0
        cam = Camera(K=K, R=R, t=-R @ c, d=d, fps=fps, resolution=[1920, 1080])
        cam.compose()
        cams.append(cam)
    return cams


def simulate_detections(cam, alpha, beta, rho, frames, noise, rng, fps_ref=30.0):
    """Raw detections (3xN rows frame, x, y) of one camera for integer ``frames``.
    Rolling shutter is applied with the true ``rho`` by fixed-point iteration on the
    row-dependent capture time."""
    f = np.asarray(frames, dtype=np.float64)
    H = cam.resolution[1]
    fx, fy, cx, cy = cam.K[0, 0], cam.K[1, 1], cam.K[0, 2], cam.K[1, 2]
    t = alpha * f + beta
    for _ in range(6):
        Xc = cam.R @ gt_trajectory(t, fps_ref) + cam.t.reshape(3, 1)
        xd, yd = distort_normalized(Xc[0] / Xc[2], Xc[1] / Xc[2], cam.d)
        x, y = fx * xd + cx, fy * yd + cy
        t = alpha * (f + rho * y / H) + beta
    x = x + rng.normal(0.0, noise, x.shape)
    y = y + rng.normal(0.0, noise, y.shape)
    ok = (x >= 0) & (x < cam.resolution[0]) & (y >= 0) & (y < H) & (Xc[2] > 0.1)
    return np.vstack((f[ok], x[ok], y[ok]))


def fit_spline(t0, t1, n_coef, fps_ref=30.0):
    """Cubic B-spline ``tck = [t, [cx,cy,cz], 3]`` (the reference's list-of-lists form,
    common.py:247, 473) over [t0, t1] with ``n_coef`` coefficients per axis, fitted to
    the ground truth by banded least squares on uniform knots."""
    from scipy.interpolate import make_lsq_spline
    n_coef = max(int(n_coef), 4)
    interior = np.linspace(t0, t1, n_coef - 2)[1:-1]
    knots = np.concatenate((np.full(4, t0), interior, np.full(4, t1)))
    if n_coef > 1500:
        # O(n) quasi-interpolant (exact for quadratics): c_j = f(xi_j) + mu_j/2 f''(xi_j) with the
        # Greville abscissa xi_j and mu_j = blossom(t^2) - xi_j^2; make_lsq_spline is O(n^2).
        a, b, c3 = knots[1:n_coef + 1], knots[2:n_coef + 2], knots[3:n_coef + 3]
        xi = (a + b + c3) / 3.0
        mu = (a * b + a * c3 + b * c3) / 3.0 - xi * xi
        c = gt_trajectory(xi, fps_ref) + 0.5 * mu * gt_trajectory(xi, fps_ref, deriv=True)
        return [knots, [c[0].copy(), c[1].copy(), c[2].copy()], 3]
    ns = max(4 * n_coef, 64)
    ts = np.linspace(t0, t1, ns)
    spl = make_lsq_spline(ts, gt_trajectory(ts, fps_ref).T, knots, k=3)
    c = np.ascontiguousarray(spl.c.T)
    return [knots, [c[0].copy(), c[1].copy(), c[2].copy()], 3]


def make_flight(nc=4, det_per_cam=1500, frames_per_knot=15.0, seed=0, noise=0.5,
                rolling_shutter=False, distortion=True, opt_calib=False, motion_type=None,
                motion_weights=1e4, gaps=(), perturb=1.0, n_coef=None, uncovered=0.02,
                init_rs=None, rho_true=None):
    """Build a BA-ready Scene: ground truth + perturbation (pose 1 deg / 0.2 m, beta +-2
    frames, alpha +-min(1e-4, 2/T), rho +-0.1, control points +-0.05 m),
    scaled by ``perturb``.

    ``gaps``: list of (lo, hi) fractions of the time axis with no trajectory -> several
    spline intervals (common.py:234-269).  ``uncovered``: fraction of the time axis at each
    end where detections exist but no spline does (exercises util.py:105 membership).
    Returns (scene, truth) where truth holds the unperturbed parameters.
    """
    rng_f = np.random.default_rng(seed)
    rng_n = np.random.default_rng(seed + 1)
    rng_p = np.random.default_rng(seed + 2)
    cams = make_cameras(nc, seed=seed, distortion=distortion)
    fps_ref = cams[0].fps
    alpha = np.array([fps_ref / c.fps for c in cams])
    T = nc * det_per_cam / np.sum(1.0 / alpha)
    beta = np.zeros(nc)
    beta[1:] = rng_f.uniform(-40.0, 40.0, nc - 1)
    rho = rng_f.uniform(0.1, 0.8, nc) if rolling_shutter else np.zeros(nc)
    if rho_true is not None:                      # (tests: read-out speeds outside [0, 1] make the rs bounds active)
        rho = np.asarray(rho_true, dtype=np.float64).copy()

    flight = Scene()
    flight.numCam = nc
    flight.cameras = cams
    flight.ref_cam = 0
    flight.sequence = list(range(nc))
    flight.find_order = False
    for i, cam in enumerate(cams):
        f0 = int(np.ceil((0.0 - beta[i]) / alpha[i]))
        f1 = int(np.floor((T - beta[i]) / alpha[i]))
        flight.detections.append(simulate_detections(cam, alpha[i], beta[i], rho[i],
                                                     np.arange(f0, f1 + 1), noise, rng_n, fps_ref))

    # spline intervals: [uncovered*T, (1-uncovered)*T] minus the gaps
    lo, hi = uncovered * T, (1.0 - uncovered) * T
    edges = [lo]
    for g0, g1 in sorted(gaps):
        edges += [g0 * T, g1 * T]
    edges.append(hi)
    ints = np.array(edges).reshape(-1, 2).T.copy()          # 2 x S
    total = np.sum(ints[1] - ints[0])
    tck = []
    for s in range(ints.shape[1]):
        span = ints[1, s] - ints[0, s]
        ncf = n_coef * span / total if n_coef else span / frames_per_knot + 3
        tck.append(fit_spline(ints[0, s], ints[1, s], int(round(ncf)), fps_ref))
    flight.spline = {'tck': tck, 'int': ints}

    truth = {'alpha': alpha.copy(), 'beta': beta.copy(), 'rs': rho.copy(), 'T': T,
             'R': [c.R.copy() for c in cams], 't': [c.t.copy() for c in cams],
             'K': [c.K.copy() for c in cams], 'd': [c.d.copy() for c in cams],
             'coef': [[a.copy() for a in k[1]] for k in tck]}

    # perturbation
    p = perturb
    flight.alpha = alpha + p * rng_p.uniform(-1, 1, nc) * min(1e-4, 2.0 / T)
    flight.beta = beta + p * rng_p.uniform(-2, 2, nc)
    if rolling_shutter:
        flight.rs = np.clip(rho + p * rng_p.uniform(-0.1, 0.1, nc), 0.0, 1.0) if init_rs is None \
            else np.asarray(init_rs, dtype=np.float64)[:nc].copy()
    else:
        flight.rs = np.zeros(nc)
    for c in cams:
        w = rng_p.normal(size=3)
        w *= p * np.deg2rad(1.0) / np.linalg.norm(w)
        c.R = hostmath.rodrigues_to_matrix(w) @ c.R
        c.t = c.t + p * rng_p.uniform(-0.2, 0.2, 3)
        if opt_calib:
            c.K = c.K.copy()
            c.K[0, 0] += p * rng_p.uniform(-10, 10)
            c.K[1, 1] += p * rng_p.uniform(-10, 10)
            c.K[0, 2] += p * rng_p.uniform(-5, 5)
            c.K[1, 2] += p * rng_p.uniform(-5, 5)
            c.d = c.d * (1.0 + p * rng_p.uniform(-0.05, 0.05, 5))
        c.compose()
    for k in tck:
        k[1] = [a + p * rng_p.uniform(-0.05, 0.05, a.shape) for a in k[1]]

    flight.settings = {'num_detections': 10 ** 9, 'opt_calib': bool(opt_calib), 'cf_exact': True,
                       'undist_points': True, 'rolling_shutter': bool(rolling_shutter),
                       'init_rs': list(map(float, flight.rs)), 'rs_bounds': False,
                       'motion_reg': motion_type is not None, 'motion_weights': motion_weights,
                       'cut_detection_second': 0, 'camera_sequence': list(range(nc)), 'ref_cam': 0,
                       'thres_Fmatix': 30, 'thres_PnP': 30, 'thres_outlier': 10,
                       'thres_triangulation': 20, 'smooth_factor': [10, 20], 'sampling_rate': 0.5,
                       'path_output': ''}
    if motion_type is not None:
        flight.settings['motion_type'] = motion_type
    flight.cf = np.zeros(nc)
    return flight, truth


def write_dataset(out_dir, nc=4, det_per_cam=5000, seed=0, noise=0.5, rolling_shutter=False,
                  distortion=False, settings=None, ground_truth=None):
    """Write a dataset4-format flight (detections ``x y frame``, camera JSON, config JSON)
    that the reference ``main.py`` runs end to end; returns the config path.  ``ground_truth`` = a frequency in
    Hz adds an RTK-style ground-truth file (the true trajectory in another similarity frame, 2 cm noise, longer
    than the flight) under 'optional inputs' so that main.py:88-90 calls align_gt."""
    os.makedirs(out_dir, exist_ok=True)
    rng_f = np.random.default_rng(seed)
    rng_n = np.random.default_rng(seed + 1)
    cams = make_cameras(nc, seed=seed, distortion=distortion)
    fps_ref = cams[0].fps
    alpha = np.array([fps_ref / c.fps for c in cams])
    T = nc * det_per_cam / np.sum(1.0 / alpha)
    # integer corresponding frames: cf_ref - alpha*cf = beta  (common.py:1014)
    cf = np.zeros(nc)
    cf[1:] = np.round(rng_f.uniform(-60, 60, nc - 1))
    beta = cf[0] - alpha * cf
    rho = rng_f.uniform(0.1, 0.8, nc) if rolling_shutter else np.zeros(nc)
    det_paths, cam_paths = [], []
    for i, cam in enumerate(cams):
        f0 = int(np.ceil((0.0 - beta[i]) / alpha[i]))
        f1 = int(np.floor((T - beta[i]) / alpha[i]))
        det = simulate_detections(cam, alpha[i], beta[i], rho[i], np.arange(f0, f1 + 1), noise,
                                  rng_n, fps_ref)
        p = os.path.join(out_dir, 'cam%d.txt' % i)
        np.savetxt(p, np.column_stack((det[1], det[2], det[0])), fmt='%.6f %.6f %d')
        det_paths.append(p)
        p = os.path.join(out_dir, 'cam%d.json' % i)
        with open(p, 'w') as fh:
            json.dump({'K-matrix': cam.K.tolist(), 'distCoeff': cam.d.tolist(), 'fps': cam.fps,
                       'resolution': cam.resolution}, fh)
        cam_paths.append(p)
    st = {'num_detections': 100000, 'opt_calib': False, 'cf_exact': True, 'sync_method': 'iter',
          'undist_points': True, 'rolling_shutter': bool(rolling_shutter),
          'init_rs': [0.5] * nc if rolling_shutter else 0, 'rs_bounds': False,
          'motion_reg': False, 'motion_weights': 1e4, 'motion_type': 'F',
          'cut_detection_second': 0.5, 'camera_sequence': [], 'ref_cam': 0, 'thres_Fmatix': 30,
          'thres_PnP': 30, 'thres_outlier': 10, 'thres_triangulation': 20,
          'smooth_factor': [10, 20], 'sampling_rate': 0.5,
          'path_output': os.path.join(out_dir, 'result.pkl')}
    st.update(settings or {})
    cfg = {'comments': ['synthetic helix flight'],
           'necessary inputs': {'path_detections': det_paths, 'path_cameras': cam_paths,
                                'corresponding_frames': cf.tolist()},
           'settings': st}
    if ground_truth:
        tau = np.arange(-6.0 * fps_ref, T + 9.0 * fps_ref, fps_ref / float(ground_truth))
        # a proper rotation (Rodrigues of a fixed vector), scale 1.7, offset
        w = np.array([0.3, -0.2, 0.5])
        th = np.linalg.norm(w)
        k = w / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        Y = 1.7 * R @ gt_trajectory(tau, fps_ref=fps_ref) + np.array([[10.0], [-4.0], [2.5]])
        Y = Y + np.random.default_rng(seed + 5).normal(size=Y.shape) * 0.02
        gp = os.path.join(out_dir, 'gt.txt')
        np.savetxt(gp, Y.T)
        cfg['optional inputs'] = {'ground_truth': {'filepath': gp, 'frequency': float(ground_truth)}}
    p = os.path.join(out_dir, 'config.json')
    with open(p, 'w') as fh:
        json.dump(cfg, fh, indent=1)
    return p
