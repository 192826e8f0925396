"""Diagnostic driver for a gpurun call: prints parity numbers instead of asserting."""
import sys, os, time, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases, helpers
from mvus_b200 import _cabi
from mvus_b200.problem import FlatProblem
from oracle import ba_oracle
import scipy.sparse as sp

for name in cases.CASES:
    try:
        fl, truth, bakw = cases.make(name)
        fp = FlatProblem(fl, fl.numCam, **bakw)
        prob = ba_oracle.Problem(fl, fl.numCam, **bakw)
        hd = _cabi.Handle(fp, max_nfev=10)
        x = prob.x0
        r = hd.residual(x); ro = prob.residual(x)
        print(name, 'n', fp.n, 'm', hd.m, 'resid rel', np.abs(r - ro).max() / np.abs(ro).max())
        r, span, J, mbase, mJ = hd.residual_jacobian(x)
        Jg = helpers.expand_jacobian(fp, span, J, mbase, mJ)
        free = prob.free_mask()
        Jo = prob.jacobian(x).tocsc() @ sp.diags(free.astype(float))
        colmax = np.maximum(abs(Jo).max(axis=0).toarray().ravel(), 1e-300)
        err = abs(Jg - Jo).tocsc().max(axis=0).toarray().ravel() / colmax
        print('  J rel', err.max(), 'resid(J call) rel', np.abs(r - ro).max() / np.abs(ro).max())
        A, g, Hss, Hcs, cost = hd.normal_equations(x)
        Jd = Jo.toarray(); go = Jd.T @ ro
        print('  cost rel', abs(cost - 0.5 * ro @ ro) / cost, 'g rel', np.abs(g - go).max() / np.abs(go).max())
        t0 = time.time(); xs, rs, st = hd.solve(fp.x0); t1 = time.time()
        print('  solve: cost0 %.6g -> %.6g (oracle at x*: %.6g) nfev %d njev %d status %d lm_it %d lam %.3g opt %.3g' % (
            st.cost0, st.cost, prob.cost(xs), st.nfev, st.njev, st.status, st.lm_iterations, st.lam, st.optimality))
        print('  ms total %.3f resjac %.3f accum %.3f solve %.3f trial %.3f launches %d wall %.3f' % (
            st.ms_total, st.ms_resjac, st.ms_accum, st.ms_solve, st.ms_trial, st.launches, t1 - t0))
        ra = prob.shipped_solve(prob.x0, max_nfev=10)
        print('  shipped@10 cost %.6g' % ra.cost)
        hd.close()
    except Exception:
        traceback.print_exc()
