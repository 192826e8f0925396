"""Install the GPU bundle adjustment under the UNMODIFIED reference: replaces the methods of
``reconstruction.common.Scene`` that sit on the BA path (BA, error_cam, remove_outliers and the
spline fit traj_to_spline either side of it, plus analysis.compare_gt.align_gt after it) so that the reference's own ``main.py`` runs with
its inner loop on the B200.  See INTEGRATION.md."""
from . import ba


def install(common_module, satellites=True):
    """common_module: the imported ``reconstruction.common``.  Returns the original BA."""
    Scene = common_module.Scene
    original = Scene.BA

    def BA(self, numCam, max_iter=10, rs=False, motion_prior=False, motion_reg=False, motion_weights=1,
           norm=False, rs_bounds=False):
        return ba.bundle_adjust(self, numCam, max_iter=max_iter, rs=rs, motion_prior=motion_prior,
                                motion_reg=motion_reg, motion_weights=motion_weights, norm=norm,
                                rs_bounds=rs_bounds)

    Scene.BA = BA
    Scene._reference_BA = original
    if satellites:
        orig_err, orig_rm = Scene.error_cam, Scene.remove_outliers

        def error_cam(self, cam_id, mode='dist', motion_prior=False, norm=False):
            if motion_prior or norm:
                return orig_err(self, cam_id, mode=mode, motion_prior=motion_prior, norm=norm)
            return ba.error_cam(self, cam_id, mode=mode)

        Scene.error_cam = error_cam
        Scene.remove_outliers = lambda self, cams, thres=30, verbose=False: ba.remove_outliers(
            self, cams, thres=thres, verbose=verbose)
        Scene._reference_error_cam, Scene._reference_remove_outliers = orig_err, orig_rm
        # the spline (re)fit either side of every BA: init (main.py:36) and triangulate's refit (common.py:812)
        from . import splfit
        Scene._reference_traj_to_spline = Scene.traj_to_spline
        Scene.traj_to_spline = lambda self, smooth_factor: splfit.traj_to_spline(self, smooth_factor)
        # ground-truth alignment at the end of main.py (main.py:88-90 binds analysis.compare_gt.align_gt by name
        # when it is imported, so this must run before main.py does)
        try:
            import importlib
            cg = importlib.import_module('analysis.compare_gt')
        except ImportError:
            cg = None
        if cg is not None and not hasattr(cg, '_reference_align_gt'):
            from . import align
            cg._reference_align_gt = cg.align_gt
            cg.align_gt = lambda flight, f_gt, gt_path, visualize=False: (
                cg._reference_align_gt(flight, f_gt, gt_path, visualize=True) if visualize
                else align.align_gt(flight, f_gt, gt_path))
    return original


def uninstall(common_module):
    """Put the reference's own methods back."""
    Scene = common_module.Scene
    for name in ('BA', 'error_cam', 'remove_outliers', 'traj_to_spline'):
        orig = Scene.__dict__.get('_reference_' + name)
        if orig is not None:
            setattr(Scene, name, orig)
            delattr(Scene, '_reference_' + name)
    import sys
    cg = sys.modules.get('analysis.compare_gt')
    if cg is not None and hasattr(cg, '_reference_align_gt'):
        cg.align_gt = cg._reference_align_gt
        del cg._reference_align_gt
