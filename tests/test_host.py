"""Host-side logic: C-ABI library exports, loud failure without a GPU, packing round trip,
Rodrigues helpers, the NumPy model of the device solver."""
import ctypes
import os
import re

import numpy as np
import pytest

import cases
from mvus_b200 import _cabi, hostmath
from mvus_b200.problem import FlatProblem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, 'include', 'mvus_ba.h')).read()
    declared = sorted(set(re.findall(r'\b(mvus_ba_\w+)\s*\(', hdr)))
    assert len(declared) >= 15
    lib = ctypes.CDLL(built_lib)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(_cabi.EXPORTS) == declared
    lib.mvus_ba_version.restype = ctypes.c_char_p
    assert b'sm_100a' in lib.mvus_ba_version()


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product path must fail loudly (no silent CPU route)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip('GPU present')
    fl, truth, bakw = cases.make('gs_plain')
    fp = FlatProblem(fl, fl.numCam, **bakw)
    with pytest.raises(_cabi.MvusError, match='no usable CUDA device'):
        _cabi.Handle(fp)
    with pytest.raises(_cabi.MvusError):
        fl.BA(fl.numCam)
    import mvus_b200
    src = ''.join(open(os.path.join(ROOT, 'mvus_b200', f)).read()
                  for f in os.listdir(os.path.join(ROOT, 'mvus_b200')) if f.endswith('.py'))
    assert 'import oracle' not in src and 'from oracle' not in src


def test_pack_unpack_roundtrip():
    for name in ('gs_plain', 'calib_KE'):
        fl, truth, bakw = cases.make(name)
        fp = FlatProblem(fl, fl.numCam, **bakw)
        x = fp.x0 + 0.01
        fp.unpack_into(fl, x)
        fp2 = FlatProblem(fl, fl.numCam, **bakw)
        assert np.abs(fp2.x0 - x).max() <= 1e-12
        for c in fl.cameras:
            assert np.allclose(c.P, c.K @ np.hstack((c.R, c.t.reshape(3, 1))), atol=1e-12, rtol=0)


def test_rodrigues_matches_opencv():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(0)
    for _ in range(50):
        w = rng.normal(size=3) * rng.choice([1e-9, 0.1, 1.0, 3.0])
        R = hostmath.rodrigues_to_matrix(w)
        assert np.abs(R - cv2.Rodrigues(w)[0]).max() <= 1e-14
        assert np.abs(hostmath.matrix_to_rodrigues(R) - cv2.Rodrigues(R)[0].ravel()).max() <= 1e-10


def test_solver_model_matches_dense_solve():
    """tests/proto/bcr_proto.py (the NumPy model ba_solve.cuh mirrors): cyclic reduction + Schur
    complement == dense solve of the damped normal equations."""
    from proto import bcr_proto
    from oracle import ba_oracle
    for bw, name in ((3, 'gs_plain'), (4, 'rs_F_gap'), (6, 'rs_bounds_dense')):
        fl, truth, bakw = cases.make(name, det_per_cam=150)
        prob = ba_oracle.Problem(fl, fl.numCam, **bakw)
        x = prob.x0
        J = prob.jacobian(x).toarray() * prob.free_mask()[None, :]
        r = prob.residual(x)
        n_ctrl = sum(prob.ncoef)
        cols = np.zeros((n_ctrl, 3), int)
        j = 0
        for s in range(prob.S):
            for l in range(prob.ncoef[s]):
                for ax in range(3):
                    cols[j, ax] = prob.coef_off[s] + ax * prob.ncoef[s] + l
                j += 1
        A, bc, D, E, Wt, perm = bcr_proto.assemble(J, r, prob.n_other, n_ctrl, cols, bw)
        lam = 1e-3
        dc, ds = bcr_proto.solve_bcr(A, bc, D, E, Wt, lam)
        H = J.T @ J
        ref = np.linalg.solve(H + lam * np.diag(np.clip(np.diag(H), 1e-6, 1e32)), -J.T @ r)
        mine = np.zeros(prob.n)
        mine[:prob.n_other] = dc
        flat = ds.reshape(-1)
        ok = perm >= 0
        mine[perm[ok]] = flat[ok]
        assert np.abs(mine - ref).max() <= 1e-8 * np.abs(ref).max()


def test_chunk_prereduction_model_matches_dense_solve():
    """tests/proto/bcr_proto.py::solve_chunked (the NumPy model csrc/ba_chunk.cuh mirrors: sequential elimination
    inside chunks, head system, back substitution) == dense solve of the damped normal equations, for chunk
    lengths that divide the block count, ragged ones and one chunk spanning everything; plus a larger random
    block-tridiagonal system against the cyclic-reduction model."""
    from proto import bcr_proto
    from oracle import ba_oracle
    for bw, name in ((3, 'gs_plain'), (4, 'rs_F_gap'), (6, 'rs_bounds_dense')):
        fl, truth, bakw = cases.make(name, det_per_cam=150)
        prob = ba_oracle.Problem(fl, fl.numCam, **bakw)
        J = prob.jacobian(prob.x0).toarray() * prob.free_mask()[None, :]
        r = prob.residual(prob.x0)
        n_ctrl = sum(prob.ncoef)
        cols = np.zeros((n_ctrl, 3), int)
        j = 0
        for s in range(prob.S):
            for l in range(prob.ncoef[s]):
                for ax in range(3):
                    cols[j, ax] = prob.coef_off[s] + ax * prob.ncoef[s] + l
                j += 1
        A, bc, D, E, Wt, perm = bcr_proto.assemble(J, r, prob.n_other, n_ctrl, cols, bw)
        lam = 1e-3
        H = J.T @ J
        ref = np.linalg.solve(H + lam * np.diag(np.clip(np.diag(H), 1e-6, 1e32)), -J.T @ r)
        for Lc in (1, 2, 5, 64):
            dc, ds = bcr_proto.solve_chunked(A, bc, D, E, Wt, lam, Lc)
            mine = np.zeros(prob.n)
            mine[:prob.n_other] = dc
            flat = ds.reshape(-1)
            ok = perm >= 0
            mine[perm[ok]] = flat[ok]
            assert np.abs(mine - ref).max() <= 1e-8 * np.abs(ref).max(), (name, Lc)
    rng = np.random.default_rng(0)
    for nb, q, ncp, Lc in ((37, 9, 12, 8), (65, 12, 7, 32), (5, 9, 4, 8)):
        D = np.zeros((nb, q, q)); E = np.zeros((nb, q, q))
        for k in range(nb):
            Jb = rng.normal(size=(4 * q, 2 * q))
            Hb = Jb.T @ Jb
            D[k] += Hb[:q, :q]
            if k + 1 < nb:
                D[k + 1] += Hb[q:, q:]
                E[k] += Hb[:q, q:]
        Wt = rng.normal(size=(nb, q, ncp + 1))
        A = np.eye(ncp) * 1e3 * q * nb
        bc = rng.normal(size=ncp)
        dc0, ds0 = bcr_proto.solve_bcr(A, bc, D, E, Wt, 1e-2)
        dc1, ds1 = bcr_proto.solve_chunked(A, bc, D, E, Wt, 1e-2, Lc)
        assert np.abs(dc0 - dc1).max() <= 1e-12 * np.abs(dc0).max()
        assert np.abs(ds0 - ds1).max() <= 1e-12 * np.abs(ds0).max()


def test_dropin_installs_on_the_reference_module():
    """mvus_b200.dropin.install replaces Scene.BA (and the error_cam / remove_outliers / traj_to_spline / align_gt
    satellites) of the UNMODIFIED reference module with the same signatures, and uninstall puts every one of them
    back."""
    import inspect
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip('reference tree not present (GPU box)')
    from mvus_b200 import dropin
    common = ref_shim.load()
    sig_before = inspect.signature(common.Scene.BA)
    orig = dropin.install(common)
    try:
        assert common.Scene.BA is not orig and common.Scene._reference_BA is orig
        assert list(inspect.signature(common.Scene.BA).parameters) == list(sig_before.parameters)
        for name, p in inspect.signature(common.Scene.BA).parameters.items():
            assert p.default == sig_before.parameters[name].default
        # a reference Scene packs to the same x0 through the product's FlatProblem as through the
        # reference's own BA (duck typing of the two Scene classes)
        fl, truth, bakw = cases.make('calib_KE')
        ref = ref_shim.to_reference_scene(fl)
        fp_ref = FlatProblem(ref, ref.numCam, **bakw)
        fp_mir = FlatProblem(fl, fl.numCam, **bakw)
        assert np.abs(fp_ref.x0 - fp_mir.x0).max() <= 1e-12 * max(1.0, np.abs(fp_mir.x0).max())
        assert fp_ref.N == fp_mir.N and (fp_ref.knots == fp_mir.knots).all()
        # the satellites either side of the BA: spline (re)fit and the ground-truth alignment main.py ends with
        from analysis import compare_gt
        assert common.Scene.traj_to_spline is not common.Scene._reference_traj_to_spline
        assert compare_gt.align_gt is not compare_gt._reference_align_gt
        assert list(inspect.signature(compare_gt.align_gt).parameters) == \
            list(inspect.signature(compare_gt._reference_align_gt).parameters)
    finally:
        dropin.uninstall(common)
    from analysis import compare_gt
    assert common.Scene.BA is orig and not hasattr(common.Scene, '_reference_BA')
    assert not hasattr(common.Scene, '_reference_traj_to_spline') and not hasattr(compare_gt, '_reference_align_gt')
    assert compare_gt.align_gt.__module__ == 'analysis.compare_gt'


@pytest.mark.parametrize('name,scramble,bw', [('gs_plain', False, 3), ('rs_F_gap', False, 4), ('rs_F_gap', True, 3),
                                              ('calib_KE', True, 5), ('rs_bounds_dense', False, 6)])
def test_k2_streaming_model_matches_dense(name, scramble, bw):
    """tests/proto/k2_stream_proto.py (the bookkeeping ba_k2.cuh implements lane by lane: swizzled
    32-detection blocks, column layout X|Y|Z|r|camera, column -> plane mapping of the fragment loads,
    ping-pong phases, per-lane flush masks, D as upper triangle / E ordered by global row)
    == dense J^T J, J^T r of the reprojection rows.  `scramble` feeds the detections of each camera
    in a random order, i.e. arbitrary span sequences (jumps, returns, uncovered gaps)."""
    import helpers
    from proto import k2_stream_proto as k2
    from mvus_b200.problem import FlatProblem
    fl, truth, bakw = cases.make(name, det_per_cam=300)
    fp = FlatProblem(fl, fl.numCam, **bakw)
    r, span, J, mbase, mJ = helpers.emul_resjac(fp, fp.x0)
    N, P, Pc, nc, C = fp.N, fp.P, fp.Pc, fp.nc, fp.C
    J = np.asarray(J).reshape(2 * P, N)
    Jg = helpers.expand_jacobian(fp, span, J).toarray()[:2 * N]
    rr = np.asarray(r)[:2 * N]
    H, g = Jg.T @ Jg, Jg.T @ rr
    nb = (fp.n_ctrl + bw - 1) // bw
    q = 3 * bw
    out = None
    rng = np.random.default_rng(5)
    for i in range(nc):
        a, b = int(fp.cam_ptr[i]), int(fp.cam_ptr[i + 1])
        order = np.arange(a, b)
        if scramble:                         # blocks of 1-7 detections in random order
            cuts = np.cumsum(rng.integers(1, 8, size=b - a))
            blocks = np.split(order, cuts[cuts < b - a])
            order = np.concatenate([blocks[k] for k in rng.permutation(len(blocks))])
        n_i = b - a
        ru = rr[2 * a + (order - a)]
        rv = rr[2 * a + n_i + (order - a)]
        # the kernels store abs(residual) rows with the sign folded into J, so r and J are consistent as they are
        out = k2.accumulate_camera(J[:P, order].T, J[P:, order].T, ru, rv, span[order], Pc, i, nc, bw, nb, out)
    A, bc, D, E, W = out
    assert np.abs(W[nb * q:]).max() == 0.0 and np.abs(D[nb:]).max() == 0.0      # nothing lands in the ghost block
    W = W[:nb * q]
    cam_cols = lambda i: np.array([i, nc + i, 2 * nc + i] + list(range(3 * nc + i * C, 3 * nc + (i + 1) * C)))
    ctrl_col = np.full(nb * q, -1)
    for s in range(fp.S):
        for l in range(int(fp.ncoef[s])):
            for ax in range(3):
                ctrl_col[3 * (int(fp.ctrl_off[s]) + l) + ax] = fp.n_other + 3 * fp.ctrl_off[s] + ax * fp.ncoef[s] + l
    Hm, gm = np.zeros_like(H), np.zeros_like(g)
    for i in range(nc):
        cc = cam_cols(i)
        Hm[np.ix_(cc, cc)] = A[i]
        gm[cc] = -bc[i]
    ok = ctrl_col >= 0
    for i in range(nc):
        cc = cam_cols(i)
        Hm[np.ix_(ctrl_col[ok], cc)] = W[ok][:, i * Pc:(i + 1) * Pc]
        Hm[np.ix_(cc, ctrl_col[ok])] = W[ok][:, i * Pc:(i + 1) * Pc].T
    gm[ctrl_col[ok]] = -W[ok, -1]
    for k in range(nb):
        rows = ctrl_col[k * q:(k + 1) * q]
        Dk = np.triu(D[k]) + np.triu(D[k], 1).T
        assert np.abs(np.tril(D[k], -1)).max() == 0.0          # upper triangle only
        rk = rows >= 0
        Hm[np.ix_(rows[rk], rows[rk])] = Dk[np.ix_(rk, rk)]
        if k + 1 < nb:
            nxt = ctrl_col[(k + 1) * q:(k + 2) * q]
            nk = nxt >= 0
            Hm[np.ix_(rows[rk], nxt[nk])] = E[k][np.ix_(rk, nk)]
            Hm[np.ix_(nxt[nk], rows[rk])] = E[k][np.ix_(rk, nk)].T
    assert np.abs(Hm - H).max() <= 1e-11 * np.abs(H).max()
    assert np.abs(gm - g).max() <= 1e-11 * np.abs(g).max()


def test_lm_driver_model_converges_like_the_shipped_solve():
    """tests/proto/lm_proto.py (NumPy model of mvus_ba_solve's LM / trust-region control with a dense
    solve): from the same start and with the same 10-evaluation cap it must not end above the shipped
    SciPy call (oracle A), and the lambda search must stay cheap (linear solves per LM step)."""
    from proto import lm_proto
    from oracle import ba_oracle
    for name in ('gs_plain', 'rs_KE_fpk30'):
        fl, truth, bakw = cases.make(name, det_per_cam=150)
        prob = ba_oracle.Problem(fl, fl.numCam, **bakw)
        ra = prob.shipped_solve(prob.x0, max_nfev=10)
        out = lm_proto.run(prob, prob.x0, max_nfev=10)
        assert out['nfev'] <= 10
        assert out['cost'] <= ra.cost * (1 + 1e-9), (name, out['cost'], ra.cost)
        assert out['solves'] <= 2 * out['steps']
