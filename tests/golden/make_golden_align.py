"""Golden fixtures for the ground-truth alignment (analysis/compare_gt.align_gt): outputs of the UNMODIFIED
reference on seeded synthetic cases (tests/helpers.make_alignment_case).  Run in the build container:
    python tests/golden/make_golden_align.py
Stores per case: the reference's align_param, tran_matrix, error, reconst_tran, gt, flight.traj after the call,
and the coarse-search mean errors recomputed with the reference's own functions (compare_gt.py:112-126)."""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import helpers                                 # noqa: E402
from oracle import ref_shim                    # noqa: E402

CASES = {'align_two_intervals': dict(name='rs_F_gap', with_time=True),
         'align_three_rows': dict(name='gs_plain', with_time=False)}


def reference_coarse(common, ref, f_gt, gt_ori):
    """compare_gt.py:95-126 with the reference's own util / transformation functions."""
    from tools import util
    from thirdparty import transformation
    alpha = ref.cameras[ref.settings['ref_cam']].fps / f_gt
    reconst = ref.spline_to_traj(sampling_rate=alpha)
    t0 = reconst[0, 0]
    reconst = np.vstack(((reconst[0] - t0) / alpha, reconst[1:]))
    gt = np.vstack((np.arange(len(gt_ori[0])), gt_ori)) if gt_ori.shape[0] == 3 else \
        np.vstack((gt_ori[0] - gt_ori[0, 0], gt_ori[1:]))
    thres = int(reconst[0, -1] / 2)
    shifts = np.arange(-thres, int(gt[0, -1] - thres))
    errs = np.empty(len(shifts))
    for k, i in enumerate(shifts):
        p1, p2 = util.match_overlap(np.vstack((reconst[0] + i, reconst[1:])), gt)
        M = transformation.affine_matrix_from_points(p1[1:], p2[1:], shear=False, scale=True)
        tran = M @ util.homogeneous(p1[1:])
        tran /= tran[-1]
        errs[k] = np.mean(np.sqrt(((p2[1:] - tran[:3]) ** 2).sum(axis=0)))
    return shifts, errs


def main():
    common = ref_shim.load()
    from analysis import compare_gt
    for case, kw in CASES.items():
        fl, gt, f_gt = helpers.make_alignment_case(**kw)
        ref = ref_shim.to_reference_scene(fl)
        path = os.path.join(HERE, '_gt_tmp.txt')
        np.savetxt(path, gt.T)
        with contextlib.redirect_stdout(io.StringIO()):
            shifts, errs = reference_coarse(common, ref, f_gt, gt)
            out = compare_gt.align_gt(ref, f_gt, path, visualize=False)
        os.remove(path)
        np.savez_compressed(os.path.join(HERE, case + '.npz'), align_param=out['align_param'],
                            tran_matrix=out['tran_matrix'], error=out['error'], reconst_tran=out['reconst_tran'],
                            gt=out['gt'], traj=ref.traj, shifts=shifts, coarse=errs)
        print(case, 'align_param', out['align_param'], 'mean error', out['error'].mean(), 'best shift',
              shifts[np.argmin(errs)])


if __name__ == '__main__':
    main()
