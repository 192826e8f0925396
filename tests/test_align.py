"""Ground-truth alignment (analysis/compare_gt.py:73-151).
CPU: the oracle restatement (oracle/align_oracle.py) against the golden outputs of the unmodified reference
(tests/golden/align_*.npz, made by tests/golden/make_golden_align.py) and against the live reference when
present.  GPU: mvus_ba_align / mvus_b200.align.align_gt against the oracle and the golden outputs."""
import os

import numpy as np
import pytest

import helpers
from oracle import align_oracle, ref_shim

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {'align_two_intervals': dict(name='rs_F_gap', with_time=True),
         'align_three_rows': dict(name='gs_plain', with_time=False)}


def _case(case):
    fl, gt, f_gt = helpers.make_alignment_case(**CASES[case])
    gold = np.load(os.path.join(HERE, 'golden', case + '.npz'))
    tck, interval = fl.spline['tck'], np.asarray(fl.spline['int'])
    return fl, gt, f_gt, gold, tck, interval


@pytest.mark.parametrize('case', list(CASES))
def test_oracle_coarse_search_matches_golden(case):
    fl, gt_ori, f_gt, gold, tck, interval = _case(case)
    alpha, t0, reconst, gt = align_oracle.preprocess(tck, interval, fl.cameras[fl.ref_cam].fps, f_gt, gt_ori)
    shifts, errs = align_oracle.coarse_errors(reconst, gt)
    assert np.array_equal(shifts, gold['shifts'])
    assert np.abs(errs - gold['coarse']).max() <= 1e-9 * np.abs(gold['coarse']).max()


@pytest.mark.parametrize('case', list(CASES))
def test_oracle_fine_error_matches_golden(case):
    fl, gt_ori, f_gt, gold, tck, interval = _case(case)
    err, M = align_oracle.fine_error(gold['align_param'], gt_ori, tck, interval)
    keep = err[err > 0]                    # the reference drops nothing here unless an error exceeds 10 x the mean
    assert len(keep) >= len(gold['error'])
    sel = keep <= 10 * keep.mean()
    assert np.abs(keep[sel] - gold['error']).max() <= 1e-9
    assert np.abs(M - gold['tran_matrix']).max() <= 1e-9 * np.abs(gold['tran_matrix']).max()


def test_oracle_similarity_matches_live_reference():
    if not ref_shim.available():
        pytest.skip('reference tree not present')
    ref_shim.load()
    from thirdparty import transformation
    rng = np.random.default_rng(0)
    for n in (3, 10, 500):
        v0 = rng.normal(size=(3, n)) * 4.0 + 7.0
        M0 = align_oracle.similarity(rng.normal(size=(3, 6)), rng.normal(size=(3, 6)))
        v1 = (M0 @ np.vstack((v0, np.ones(n))))[:3] + rng.normal(size=(3, n)) * 0.01
        Mr = transformation.affine_matrix_from_points(v0, v1, shear=False, scale=True)
        assert np.abs(align_oracle.similarity(v0, v1) - Mr).max() <= 1e-11 * np.abs(Mr).max()


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize('case', list(CASES))
def test_device_fits_match_oracle(case, built_lib):
    """mvus_ba_align: every shift of the coarse search (spline = interpolating spline of the ground truth) and
    fine-stage evaluations at several (alpha, beta) against the oracle."""
    from mvus_b200 import align, ba
    fl, gt_ori, f_gt, gold, tck, interval = _case(case)
    alpha, t0, reconst, gt = align_oracle.preprocess(tck, interval, fl.cameras[fl.ref_cam].fps, f_gt, gt_ori)
    j, mean_err, shifts = align.coarse_search(reconst, gt)
    assert np.array_equal(shifts, gold['shifts'])
    assert np.abs(mean_err - gold['coarse']).max() <= 1e-8 * np.abs(gold['coarse']).max()
    assert j == gold['shifts'][np.argmin(gold['coarse'])]
    hd, fp = align._spline_handle(fl)
    try:
        for model in (gold['align_param'], gold['align_param'] * [1.0005, 1.0], gold['align_param'] + [0.0, 3.7]):
            a, b = model
            t_gt = a * np.arange(gt_ori.shape[1]) + b if gt_ori.shape[0] == 3 else a * (gt_ori[0] - gt_ori[0, 0]) + b
            me, cnt, M, err = hd.align(fp.x0, t_gt, gt_ori[-3:], [0.0], spline_is_src=True, want=0)
            eo, Mo = align_oracle.fine_error(model, gt_ori, tck, interval)
            assert cnt[0] == np.count_nonzero(eo)
            assert np.abs(err - eo).max() <= 1e-9 * max(1.0, eo.max())
            assert np.abs(M[0] - Mo).max() <= 1e-9 * np.abs(Mo).max()
    finally:
        hd.close()


@pytest.mark.gpu
@pytest.mark.parametrize('case', list(CASES))
def test_align_gt_matches_the_reference(case, built_lib, tmp_path, capsys):
    """mvus_b200.align.align_gt against the golden output of the reference's align_gt on the same flight and
    ground-truth file (and against the live reference when it travelled to this box)."""
    from mvus_b200 import align
    fl, gt_ori, f_gt, gold, tck, interval = _case(case)
    path = str(tmp_path / 'gt.txt')
    np.savetxt(path, gt_ori.T)
    out = align.align_gt(fl, f_gt, path)
    assert 'The mean error (distance) is' in capsys.readouterr().out
    refs = [dict(gold)]
    if ref_shim.available():
        ref_shim.load()
        from analysis import compare_gt
        ref = ref_shim.to_reference_scene(fl)
        refs.append(compare_gt.align_gt(ref, f_gt, path, visualize=False))
        refs[-1]['traj'] = ref.traj
    for want in refs:
        # the optimum of a noisy 2-parameter robust fit found through finite-difference Jacobians: both drivers are
        # the same SciPy call, their residuals agree to 1e-9, the minimiser to ~1e-6
        assert np.abs(out['align_param'] - want['align_param']).max() <= 1e-5 * np.abs(want['align_param']).max()
        assert np.abs(out['tran_matrix'] - want['tran_matrix']).max() <= 1e-5 * np.abs(want['tran_matrix']).max()
        assert out['error'].shape == want['error'].shape
        assert abs(out['error'].mean() - want['error'].mean()) <= 1e-6
        assert np.abs(out['reconst_tran'] - want['reconst_tran']).max() <= 1e-3
        assert np.array_equal(out['gt'], want['gt'])
        assert fl.traj.shape == want['traj'].shape
    assert align.align_gt(fl, f_gt, '') is None
