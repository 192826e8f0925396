"""Import the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY -- used by
tests/golden/make_golden.py, by the live-reference tests and by bench.py's CPU arm.

Search order: $MVUS_REFERENCE, then /root/reference (the build container), then oracle/_ref
(the verbatim copy oracle/make_ref.py makes; git-ignored, it travels to the GPU box with the
snapshot, where /root/reference does not exist).  Nothing under mvus_b200/ imports this.

Two shims, no source edits:
  1. matplotlib is absent and reconstruction/common.py:16-18 imports it at module level
     -> stub modules ``matplotlib.pyplot`` / ``mpl_toolkits.mplot3d``.
  2. ``np.asfarray`` was removed in NumPy 2 and create_scene uses it (common.py:1205-1222).
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for root in (os.environ.get('MVUS_REFERENCE'), '/root/reference', os.path.join(_HERE, '_ref')):
        if root and os.path.isdir(os.path.join(root, 'multiviewunsynch', 'reconstruction')):
            return root
    return os.environ.get('MVUS_REFERENCE', '/root/reference')


REF_ROOT = _find_root()
REF_PKG = os.path.join(REF_ROOT, 'multiviewunsynch')


def available():
    return os.path.isdir(os.path.join(REF_PKG, 'reconstruction'))


def load():
    """Return the reference's ``reconstruction.common`` module."""
    import numpy as np
    if not available():
        raise ImportError('reference not present at %s' % REF_ROOT)
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = types.ModuleType('matplotlib')
            plt = types.ModuleType('matplotlib.pyplot')
            mpl.pyplot = plt
            tk = types.ModuleType('mpl_toolkits')
            m3 = types.ModuleType('mpl_toolkits.mplot3d')
            m3.Axes3D = object
            tk.mplot3d = m3
            sys.modules.update({'matplotlib': mpl, 'matplotlib.pyplot': plt, 'mpl_toolkits': tk,
                                'mpl_toolkits.mplot3d': m3})
    if not hasattr(np, 'asfarray'):
        np.asfarray = lambda a, dtype=float: np.asarray(a, dtype=dtype)
    if REF_PKG not in sys.path:
        sys.path.insert(0, REF_PKG)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from reconstruction import common
    return common


def to_reference_scene(scene):
    """Copy a mirror Scene (mvus_b200.scene.Scene) into a reference ``common.Scene`` so the
    reference's own BA / error functions can be run on it."""
    import copy
    import numpy as np
    common = load()
    ref = common.Scene()
    ref.numCam = scene.numCam
    for c in scene.cameras:
        rc = common.Camera(K=np.array(c.K, float), R=np.array(c.R, float), t=np.array(c.t, float),
                           d=np.array(c.d, float), fps=c.fps, resolution=list(c.resolution))
        rc.compose()
        ref.cameras.append(rc)
    ref.detections = [np.array(d, float) for d in scene.detections]
    ref.alpha = np.array(scene.alpha, float)
    ref.beta = np.array(scene.beta, float)
    ref.rs = np.array(scene.rs, float)
    ref.cf = np.array(scene.cf, float)
    ref.sequence = list(scene.sequence)
    ref.settings = copy.deepcopy(scene.settings)
    ref.ref_cam = scene.ref_cam
    ref.find_order = scene.find_order
    ref.spline = {'tck': [[np.array(t[0], float), [np.array(a, float) for a in t[1]], int(t[2])]
                          for t in scene.spline['tck']],
                  'int': np.array(scene.spline['int'], float)}
    ref.detection_to_global()
    return ref


def capture_ba(ref_scene, numCam, **kw):
    """Run the reference's Scene.BA with ``least_squares`` hooked (common.py:669-670) and
    return (error_BA, x0, jac_sparsity, kwargs) without solving."""
    common = load()
    box = {}

    class _Stop(Exception):
        pass

    def hook(fn, x0, **kwargs):
        box.update(fn=fn, x0=x0.copy(), kwargs=kwargs)
        raise _Stop()

    orig = common.least_squares
    common.least_squares = hook
    try:
        try:
            ref_scene.BA(numCam, **kw)
        except _Stop:
            pass
    finally:
        common.least_squares = orig
    return box['fn'], box['x0'], box['kwargs'].get('jac_sparsity'), box['kwargs']


def reference_error_BA(ref_scene, numCam, x, motion_reg=False, motion_weights=1):
    """The body of the reference's ``error_BA`` closure (common.py:448-487) evaluated with the
    reference's OWN methods (Camera.vector2P, Scene.all_detect_to_traj, Scene.error_cam,
    Scene.error_motion) on a reference Scene -- without ``jac_BA`` (common.py:490-610), whose
    dense int64 m x n pattern makes ``capture_ba`` infeasible beyond ~1e5 rows.
    Only the ten lines of glue that split x are restated here; every number comes from the
    reference.  Mutates ref_scene exactly as error_BA does."""
    import numpy as np
    s = ref_scene
    calib = s.settings['opt_calib']
    C = 15 if calib else 6
    seq = s.sequence[:numCam]
    parts = np.split(np.asarray(x, dtype=float), [numCam, 2 * numCam, 3 * numCam, 3 * numCam + numCam * C])
    s.alpha[seq], s.beta[seq], s.rs[seq] = parts[0], parts[1], parts[2]
    cams = np.split(parts[3], numCam)
    for i in range(numCam):
        s.cameras[seq[i]].vector2P(cams[i], calib=calib)                 # common.py:458-460
    if motion_reg:
        s.all_detect_to_traj(seq)                                        # common.py:462-464
    off = 0
    for i, t in enumerate(s.spline['tck']):                              # common.py:469-473
        nco = len(t[1][0])
        blk = parts[4][off:off + 3 * nco].reshape(3, -1)
        s.spline['tck'][i][1] = [blk[0], blk[1], blk[2]]
        off += 3 * nco
    err = [s.error_cam(seq[i], mode='each') for i in range(numCam)]      # common.py:476-479
    if motion_reg:
        err.append(s.error_motion(seq, motion_reg=True, motion_weights=motion_weights))   # common.py:483-485
    return np.concatenate(err)


def run_main(config_path, install=None, uninstall=None):
    """Run the reference's own main.py (main.py:18-97) on a config file, optionally between
    ``install(common)`` / ``uninstall(common)`` (the drop-in hook).  Returns the final Scene
    (``flight``) and the printed log."""
    import contextlib
    import io
    import runpy
    common = load()
    if install:
        install(common)
    buf = io.StringIO()
    argv = sys.argv
    sys.argv = ['main.py', config_path]
    try:
        with contextlib.redirect_stdout(buf):
            g = runpy.run_path(os.path.join(REF_PKG, 'main.py'), run_name='__main__')
    finally:
        sys.argv = argv
        if uninstall:
            uninstall(common)
    return g['flight'], buf.getvalue()


def reference_bookkeeping(scene, numCam, rs=False, motion_reg=False, motion_weights=1, rs_bounds=False, x=None):
    """The pickled bookkeeping arrays (README.md:205-231) exactly as the REFERENCE leaves them
    after one ``error_BA(x)`` at x (default: the scene's own parameters): ``visible``
    (compute_visibility, common.py:427-438, evaluated before, as jac_BA does at common.py:493),
    ``global_traj / global_detections / frame_id_all / global_time_stamps_all``
    (all_detect_to_traj, common.py:887-944) and ``traj`` (unit-step samples, common.py:379).
    Returns a dict of arrays."""
    import numpy as np
    ref = to_reference_scene(scene)
    ref.compute_visibility()
    out = {'visible_%d' % i: np.asarray(v).copy() for i, v in enumerate(ref.visible)}
    if x is None:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            _, x, _, _ = capture_ba(to_reference_scene(scene), numCam, rs=rs, motion_reg=motion_reg,
                                    motion_weights=motion_weights, rs_bounds=rs_bounds)
    reference_error_BA(ref, numCam, x, motion_reg=motion_reg, motion_weights=motion_weights)
    if motion_reg:
        for k in ('global_traj', 'global_detections', 'frame_id_all', 'global_time_stamps_all', 'traj'):
            out[k] = np.asarray(getattr(ref, k)).copy()
    return out
