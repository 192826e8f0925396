"""Import the UNMODIFIED reference (/root/reference) in the build container.  TEST
INFRASTRUCTURE ONLY -- used by tests/golden/make_golden.py and by the live-reference tests;
/root/reference does not exist on the GPU box, so nothing in ``-m gpu`` tests, smoke() or
bench.py goes through here.

Two shims, no source edits:
  1. matplotlib is absent and reconstruction/common.py:16-18 imports it at module level
     -> stub modules ``matplotlib.pyplot`` / ``mpl_toolkits.mplot3d``.
  2. ``np.asfarray`` was removed in NumPy 2 and create_scene uses it (common.py:1205-1222).
"""
import os
import sys
import types

REF_ROOT = os.environ.get('MVUS_REFERENCE', '/root/reference')
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
