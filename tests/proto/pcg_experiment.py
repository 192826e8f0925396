"""Why the damped normal equations are solved directly and not by PCG: spectrum of
the BA normal matrix and convergence history of preconditioned CG in FP64 on the oracle's Jacobian of
seeded test flights (TEST INFRASTRUCTURE; run on the CPU: python tests/proto/pcg_experiment.py).
Reported per flight: eigenvalue extremes / condition of J^T J and of the Marquardt-damped, diagonally
scaled matrix at the lambda the LM driver typically works at; then, for block-Jacobi PCG (camera blocks +
3x3 control-point blocks, the preconditioner a matrix-free GPU PCG could afford) the relative residual
and -- what LM needs -- the relative error of the STEP against the exact solve after k iterations."""
import os
import sys

import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases                              # noqa: E402
from oracle import ba_oracle              # noqa: E402


def pcg(H, b, Minv, iters, x_exact):
    x = np.zeros_like(b)
    r = b.copy()
    z = Minv(r)
    p = z.copy()
    rz = r @ z
    hist = []
    for k in range(1, iters + 1):
        Hp = H @ p
        a = rz / (p @ Hp)
        x += a * p
        r -= a * Hp
        z = Minv(r)
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
        if k in (10, 30, 100, 300, 1000, 3000):
            hist.append((k, np.linalg.norm(r) / np.linalg.norm(b), np.linalg.norm(x - x_exact) / np.linalg.norm(x_exact)))
    return hist


def main():
    for name in ('gs_margin', 'rs_F_gap', 'rs_KE_fpk30'):
        fl, truth, bakw = cases.make(name)
        prob = ba_oracle.Problem(fl, fl.numCam, **bakw)
        free = prob.free_mask()
        J = prob.jacobian(prob.x0).toarray()[:, free]
        r = prob.residual(prob.x0)
        H = J.T @ J
        g = J.T @ r
        w = np.linalg.eigvalsh(H)
        d = np.clip(np.diag(H), 1e-6, None)
        print('%s: n = %d free unknowns, m = %d rows' % (name, H.shape[0], J.shape[0]))
        print('  J^T J eigenvalues: max %.3e, 8 smallest %s' % (w[-1], ' '.join('%.2e' % v for v in w[:8])))
        print('  condition without the 7 gauge modes: %.2e' % (w[-1] / max(w[7], 1e-300)))
        for lam in (1e-4, 1e-2):
            Hd = H + lam * np.diag(d)
            Hs = Hd / np.sqrt(np.outer(d, d))                      # Jacobi-scaled: what diagonal PCG sees
            ws = np.linalg.eigvalsh(Hs)
            x_exact = sla.cho_solve(sla.cho_factor(Hd), -g)
            # block-Jacobi: camera blocks (alpha, beta, rho, pose of one camera are spread in x: use the
            # exact camera-parameter index sets) + 3x3 control-point blocks
            nc, C = prob.nc, prob.C
            idx = np.cumsum(free) - 1
            blocks = []
            for i in range(nc):
                cols = [i, nc + i, 2 * nc + i] + list(range(3 * nc + i * C, 3 * nc + (i + 1) * C))
                blocks.append([idx[c] for c in cols if free[c]])
            off = prob.n_other
            for s, t in enumerate(fl.spline['tck']):
                nco = len(t[1][0])
                for l in range(nco):
                    blocks.append([idx[off + a * nco + l] for a in range(3)])
                off += 3 * nco
            inv = [(np.array(bk), np.linalg.inv(Hd[np.ix_(bk, bk)])) for bk in blocks if len(bk)]

            def Minv(v):
                out = np.zeros_like(v)
                for bk, Bi in inv:
                    out[bk] = Bi @ v[bk]
                return out
            hist = pcg(Hd, -g, Minv, 3000, x_exact)
            print('  lambda %.0e: condition of the damped, Jacobi-scaled matrix %.2e' % (lam, ws[-1] / ws[0]))
            print('    block-Jacobi PCG   ' + '   '.join('k=%d: |r|/|b| %.1e, step error %.1e' % h for h in hist))


if __name__ == '__main__':
    main()
