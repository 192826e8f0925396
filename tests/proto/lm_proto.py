"""NumPy model of the LM / trust-region driver of mvus_ba_solve (TEST INFRASTRUCTURE).

Same control flow as mvus_b200/csrc/mvus_ba.cu (damped normal equations with Marquardt scaling,
lambda searched so that the scaled step length lies in [lo, hi] x Delta, SciPy's radius update),
with a dense solve instead of the device solver.  Used to study how many linear solves the lambda
search needs per LM step (the device solve is 2/3 of a step at config 4) without a GPU.

search='powerlaw' is the shipped search (running power-law model + log-log bracket interpolation);
search='hebden' replaces the model by the rational one of More / Hebden, |delta|_D ~ a / (b + lambda),
fitted through the last two samples of the CURRENT normal equations."""
import numpy as np
import scipy.linalg as sla

DIAG_MIN, DIAG_MAX, FLOOR = 1e-6, 1e32, 1e-2


class Model:
    def __init__(self, prob):
        self.prob = prob
        self.free = prob.free_mask().astype(float)
        self.n_other = prob.n_other

    def normal(self, x):
        J = self.prob.jacobian(x).toarray() * self.free[None, :]
        r = self.prob.residual(x)
        H = J.T @ J
        g = J.T @ r
        d = np.clip(np.diag(H), DIAG_MIN, DIAG_MAX)
        ds = d[self.n_other:]
        d[self.n_other:] = np.maximum(ds, FLOOR * ds.sum() / len(ds))
        return H, g, d, 0.5 * r @ r


def run(prob, x0, max_nfev=10, search='powerlaw', band=(0.5, 1.5), verbose=False):
    m = Model(prob)
    x = x0.copy()
    H, g, d, F = m.normal(x)
    nfev, solves, steps = 1, 0, 0
    lam, Delta, pexp = 1e-4, -1.0, 2.0 / 3.0
    last_l, last_n = -1.0, 0.0
    lam_min = 1e-10
    hist = []                                   # (lambda, norm) samples of the current normal equations

    def solve(l):
        nonlocal solves
        solves += 1
        try:
            c = sla.cho_factor(H + l * np.diag(d))
        except sla.LinAlgError:
            return None, 1e300
        dl = -sla.cho_solve(c, g)
        nrm = float(np.sqrt(dl @ (d * dl)))
        hist.append((l, nrm))
        return dl, nrm

    def hebden(target):
        (l1, n1), (l2, n2) = hist[-2], hist[-1]
        if n1 == n2 or l1 == l2:
            return None
        b = (n2 * l2 - n1 * l1) / (n1 - n2)
        a = n1 * (b + l1)
        l = a / target - b
        return l if np.isfinite(l) and l > 0 else None

    while nfev < max_nfev:
        if Delta > 0 and last_l > 0 and 0 < last_n < 1e299:
            if search == 'hebden':
                lam = min(max(last_l * last_n / Delta, lam_min), 1e30) if last_n > Delta else \
                    min(max(last_l * (last_n / Delta) ** (1.0 / pexp), lam_min), 1e30)
            else:
                lam = min(max(last_l * (last_n / Delta) ** (1.0 / pexp), lam_min), 1e30)
        dl, nrm = solve(lam)
        ok = dl is not None
        if ok and Delta < 0:
            Delta = nrm
        prev_l, prev_n = (lam if ok else -1.0), nrm
        lo_l = lo_n = hi_l = hi_n = -1.0
        for its in range(10):
            if not ok or nrm > band[1] * Delta:
                lo_l, lo_n = lam, nrm
            elif nrm < band[0] * Delta and lam > lam_min:
                hi_l, hi_n = lam, nrm
            else:
                break
            new = None
            if search == 'hebden' and len(hist) >= 2 and ok and hist[-2][1] < 1e299:
                new = hebden(Delta)
                if new is not None:
                    if lo_l > 0 and hi_l > 0:       # stay inside the bracket's middle 80 %
                        a_, b_ = np.log(lo_l), np.log(hi_l)
                        new = float(np.exp(min(max(np.log(new), a_ + 0.1 * (b_ - a_)), a_ + 0.9 * (b_ - a_))))
                    elif lo_l > 0:
                        new = max(new, lam * 1.5)
                    else:
                        new = max(min(new, lam / 1.5), lam_min)
            if new is not None:
                lam = new
            elif lo_l > 0 and hi_l > 0 and lo_n < 1e299:
                a_, b_ = np.log(lo_l), np.log(hi_l)
                w = (np.log(lo_n) - np.log(Delta)) / (np.log(lo_n) - np.log(hi_n))
                w = min(max(w, 0.25), 0.75)
                lam = float(np.exp(a_ + w * (b_ - a_)))
            elif lo_l > 0 and hi_l > 0:
                lam = float(np.sqrt(lo_l * hi_l))
            elif lo_l > 0:
                f = (nrm / Delta) ** (1.0 / pexp) if ok else 10.0
                lam = max(lam * f, lam * 2.0)
                if lam > 1e30:
                    break
            else:
                lam = max(min(lam * (nrm / Delta) ** (1.0 / pexp), lam * 0.5), lam_min)
            dl, nrm = solve(lam)
            ok = dl is not None
            if ok and Delta < 0:
                Delta = nrm
            if ok and prev_l > 0 and prev_n < 1e299 and lam != prev_l and nrm > 0 and prev_n > 0:
                pe = -np.log(nrm / prev_n) / np.log(lam / prev_l)
                if np.isfinite(pe):
                    pexp = min(max(pe, 0.15), 1.0)
            if ok:
                prev_l, prev_n = lam, nrm
        if not ok:
            break
        last_l, last_n = lam, nrm
        pred = 0.5 * (lam * nrm ** 2 - g @ dl)
        xt = x + dl
        Fn = prob.cost(xt)
        nfev += 1
        steps += 1
        actual = F - Fn
        ratio = actual / pred if pred > 0 and np.isfinite(Fn) else -1.0
        if verbose:
            print('nfev %d F %.6e Fn %.6e ratio %.3f lam %.3e |d| %.3e Delta %.3e solves %d' % (nfev, F, Fn, ratio, lam, nrm, Delta, solves))
        if ratio < 0.25:
            Delta = 0.25 * nrm
        elif ratio > 0.75 and nrm > 0.7 * Delta:
            Delta *= 2.0
        if np.isfinite(Fn) and actual > 0:
            x = xt
            if nfev >= max_nfev:
                F = Fn
                break
            H, g, d, F = m.normal(x)
            hist = []
    return dict(x=x, cost=F, nfev=nfev, solves=solves, steps=steps)
