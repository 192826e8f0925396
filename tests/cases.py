"""Seeded synthetic flights shared by the CPU and GPU parity tests (small enough for the oracle
to finish in seconds).  Each case: (make_flight kwargs, BA kwargs)."""
CASES = {
    # global shutter, no distortion, no motion prior: BASELINE config 1 settings, reduced size
    'gs_plain': (dict(nc=4, det_per_cam=400, rolling_shutter=False, distortion=False), dict()),
    # rolling shutter + distortion + least-force prior, two spline intervals (config 2 settings)
    'rs_F_gap': (dict(nc=4, det_per_cam=400, rolling_shutter=True, distortion=True, motion_type='F',
                      gaps=[(0.4, 0.5)]), dict(rs=True, motion_reg=True, motion_weights=1e4)),
    # opt_calib + kinetic-energy prior, three intervals (config 3 settings)
    'calib_KE': (dict(nc=3, det_per_cam=300, rolling_shutter=True, distortion=True, opt_calib=True,
                      motion_type='KE', gaps=[(0.2, 0.3), (0.6, 0.62)]),
                 dict(rs=True, motion_reg=True, motion_weights=1e2)),
    # rs bounds + dense knots (wider motion band), 2 cameras only out of 3
    'rs_bounds_dense': (dict(nc=3, det_per_cam=250, rolling_shutter=True, distortion=True, motion_type='F',
                             frames_per_knot=3.0), dict(rs=True, motion_reg=True, motion_weights=1e2,
                                                        rs_bounds=True)),
    # everything strictly covered, low noise: well-posed optimum for the cost-parity test
    'covered': (dict(nc=4, det_per_cam=400, rolling_shutter=True, distortion=True, uncovered=-0.05,
                     noise=0.2, motion_type='KE'), dict(rs=True, motion_reg=True, motion_weights=1e2)),
    # well-posed optima for the two-sided cost-parity test (everything covered, mild start)
    'gs_margin': (dict(nc=4, det_per_cam=400, rolling_shutter=False, distortion=False, uncovered=-0.05,
                       noise=0.2, perturb=0.3), dict()),
    'rs_KE_fpk30': (dict(nc=5, det_per_cam=600, rolling_shutter=True, distortion=True, uncovered=-0.05,
                         noise=0.2, perturb=0.3, motion_type='KE', frames_per_knot=30.0),
                    dict(rs=True, motion_reg=True, motion_weights=1e2)),
}


def make(name, **over):
    from mvus_b200 import synth
    kw, bakw = CASES[name]
    kw = dict(kw)
    kw.update(over)
    fl, truth = synth.make_flight(**kw)
    return fl, truth, dict(bakw)
