// Host-side construction of the static spline tables the kernels read (built once per
// mvus_ba_set_splines; knots are fixed during a BA, common.py:470-473 only rewrites
// coefficients).  Pure C++ (no CUDA) so that tests/emul can reuse it.
//
//   spanpoly : per knot span, the 4 active B-spline basis functions as cubic polynomials
//              in (t - t_l)  (Taylor coefficients from the Piegl-Tiller derivative
//              recurrence, "The NURBS Book" A2.3) -- replaces FITPACK's per-point de Boor
//              recursion (scipy splev, called at common.py:331) by 4 Horner evaluations.
//   lut      : uniform-bucket table  bucket -> first candidate span, so a detection finds
//              its span in O(1) instead of a bisection over up to 2e5 knots.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

namespace mvus {

struct HostSplineTables {
    int S = 0;
    int64_t n_ctrl = 0;
    std::vector<double> int_a, int_b, knots, spanpoly, span_t0, lut_t0, lut_invh;
    std::vector<int64_t> knot_off, ctrl_off, xoff, lut_off;
    std::vector<int> ncoef, deg, lut_n, lut;
};

// ders[d][m], d = 0..p, m = 0..p: d-th derivative of basis function (span - p + m) at u.
inline void basis_ders(const double* U, int span, int p, double u, double ders[4][4]) {
    double ndu[4][4], a[2][4], left[4], right[4];
    ndu[0][0] = 1.0;
    for (int j = 1; j <= p; ++j) {
        left[j] = u - U[span + 1 - j];
        right[j] = U[span + j] - u;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            ndu[j][r] = right[r + 1] + left[j - r];
            const double temp = ndu[r][j - 1] / ndu[j][r];
            ndu[r][j] = saved + right[r + 1] * temp;
            saved = left[j - r] * temp;
        }
        ndu[j][j] = saved;
    }
    for (int j = 0; j <= p; ++j) ders[0][j] = ndu[j][p];
    for (int r = 0; r <= p; ++r) {
        int s1 = 0, s2 = 1;
        a[0][0] = 1.0;
        for (int k = 1; k <= p; ++k) {
            double d = 0.0;
            const int rk = r - k, pk = p - k;
            if (r >= k) {
                a[s2][0] = a[s1][0] / ndu[pk + 1][rk];
                d = a[s2][0] * ndu[rk][pk];
            }
            const int j1 = (rk >= -1) ? 1 : -rk;
            const int j2 = (r - 1 <= pk) ? k - 1 : p - r;
            for (int j = j1; j <= j2; ++j) {
                a[s2][j] = (a[s1][j] - a[s1][j - 1]) / ndu[pk + 1][rk + j];
                d += a[s2][j] * ndu[rk + j][pk];
            }
            if (r <= pk) {
                a[s2][k] = -a[s1][k - 1] / ndu[pk + 1][r];
                d += a[s2][k] * ndu[r][pk];
            }
            ders[k][r] = d;
            const int t = s1; s1 = s2; s2 = t;
        }
    }
    int r = p;
    for (int k = 1; k <= p; ++k) {
        for (int j = 0; j <= p; ++j) ders[k][j] *= r;
        r *= (p - k);
    }
}

// interval: 2*S (starts then ends); knot_ptr[S+1]; degree[S]; n_other = first spline
// coefficient index in x.  Returns false on malformed input.
inline bool build_spline_tables(int S, const double* interval, const int64_t* knot_ptr,
                                const double* knots, const int32_t* degree, int64_t n_other,
                                HostSplineTables& T) {
    T = HostSplineTables();
    T.S = S;
    T.int_a.assign(interval, interval + S);
    T.int_b.assign(interval + S, interval + 2 * S);
    T.knots.assign(knots, knots + knot_ptr[S]);
    T.knot_off.resize(S); T.ncoef.resize(S); T.deg.resize(S);
    T.ctrl_off.resize(S + 1); T.xoff.resize(S);
    T.lut_off.resize(S); T.lut_n.resize(S); T.lut_t0.resize(S); T.lut_invh.resize(S);
    int64_t ctrl = 0, xo = n_other;
    for (int s = 0; s < S; ++s) {
        const int k = degree[s];
        if (k != 1 && k != 3) return false;
        const int64_t nk = knot_ptr[s + 1] - knot_ptr[s];
        const int64_t nco = nk - k - 1;
        if (nco < k + 1) return false;
        T.knot_off[s] = knot_ptr[s];
        T.ncoef[s] = (int)nco;
        T.deg[s] = k;
        T.ctrl_off[s] = ctrl;
        T.xoff[s] = xo;
        ctrl += nco;
        xo += 3 * nco;
    }
    T.ctrl_off[S] = ctrl;
    T.n_ctrl = ctrl;
    T.spanpoly.assign((size_t)ctrl * 16, 0.0);
    T.span_t0.assign((size_t)ctrl, 0.0);
    for (int s = 0; s < S; ++s) {
        const double* U = knots + knot_ptr[s];
        const int k = T.deg[s], nco = T.ncoef[s];
        // (200 000 spans at config 4: 57 ms on one host core, inside every Scene.BA call -> split over threads)
        auto span_range = [&](int l0, int l1) {
            for (int l = l0; l < l1; ++l) {
                const int64_t g = T.ctrl_off[s] + l;
                T.span_t0[g] = U[l];
                if (!(U[l + 1] > U[l])) continue;      // empty span: never selected
                double ders[4][4];
                basis_ders(U, l, k, U[l], ders);
                double fact = 1.0;
                for (int d = 0; d <= k; ++d) {
                    if (d > 0) fact *= d;
                    for (int m = 0; m <= k; ++m)          // basis (l-k+m) -> slot (3-k+m)
                        T.spanpoly[(size_t)g * 16 + (3 - k + m) * 4 + d] = ders[d][m] / fact;
                }
            }
        };
        const int nsp = nco - k;
        const int nthr = nsp >= 32768 ? (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency())) : 1;
        if (nthr <= 1) span_range(k, nco);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nthr; ++t)
                th.emplace_back(span_range, k + (int)((int64_t)nsp * t / nthr), k + (int)((int64_t)nsp * (t + 1) / nthr));
            for (auto& x : th) x.join();
        }
        // lookup table: ~2 buckets per span
        const int nspans = nco - k;
        const int nb = 2 * nspans + 1;
        const double t0 = U[k], t1 = U[nco];
        const double h = (t1 - t0) / nb;
        T.lut_off[s] = (int64_t)T.lut.size();
        T.lut_n[s] = nb;
        T.lut_t0[s] = t0;
        T.lut_invh[s] = h > 0 ? 1.0 / h : 0.0;
        int l = k;
        for (int b = 0; b < nb; ++b) {
            // one bucket of slack to the left: the device rounds (t - t0) * invh itself
            const double tb = t0 + (b > 0 ? b - 1 : 0) * h;
            while (l < nco - 1 && U[l + 1] <= tb) ++l;
            T.lut.push_back(l);
        }
    }
    return true;
}

// Motion-prior sample grid: Scene.spline_to_traj() (common.py:289-297) samples the global
// unit grid np.arange(int[0,0], int[1,-1], 1) and keeps, per interval, a <= tau <= b;
// error_motion (common.py:380, 411) then groups the samples by util.sampling membership
// (a <= tau < b).  flags: bit0 in a group, bit1 predecessor in the same group, bit2 successor.
inline void build_motion_samples(const HostSplineTables& T, std::vector<double>& tau,
                                 std::vector<int>& spl, std::vector<unsigned char>& flags) {
    tau.clear(); spl.clear(); flags.clear();
    if (T.S == 0) return;
    const double start = T.int_a[0], stop = T.int_b[T.S - 1];
    const int64_t cnt = (int64_t)std::ceil((stop - start) / 1.0);
    std::vector<int> member;
    tau.reserve((size_t)cnt + 2); spl.reserve((size_t)cnt + 2); member.reserve((size_t)cnt + 2);
    for (int s = 0; s < T.S; ++s) {
        int64_t i0 = (int64_t)std::floor(T.int_a[s] - start) - 1, i1 = (int64_t)std::ceil(T.int_b[s] - start) + 1;
        if (i0 < 0) i0 = 0;
        if (i1 > cnt - 1) i1 = cnt - 1;
        for (int64_t i = i0; i <= i1; ++i) {
            const double v = start + (double)i * 1.0;
            if (v >= T.int_a[s] && v <= T.int_b[s]) {
                tau.push_back(v);
                spl.push_back(s);
                int hit = -1;
                for (int q = 0; q < T.S; ++q) {
                    const bool ga = (v - T.int_a[q]) >= 0.0, gb = (v - T.int_b[q]) >= 0.0;
                    if (ga != gb) hit = q;
                }
                member.push_back(hit);
            }
        }
    }
    const int64_t M = (int64_t)tau.size();
    flags.assign((size_t)M, 0);
    for (int64_t j = 0; j < M; ++j) {
        if (member[j] < 0) continue;
        spl[j] = member[j];
        unsigned char f = 1;
        if (j > 0 && member[j - 1] == member[j]) f |= 2;
        if (j + 1 < M && member[j + 1] == member[j]) f |= 4;
        flags[j] = f;
    }
}

// Largest number of consecutive control points a motion-prior row touches (4 = one sample's support): the rows
// of sample j involve the samples j-1 (KE, F) and j+1 (F) of the same group.  The samples are ascending inside a
// spline, so the knot span follows by walking forward (a bisection per sample over 2e5 knots was 140 ms at
// config 4 -- host time inside every Scene.BA call).
inline int motion_spread(const HostSplineTables& T, const std::vector<double>& tau, const std::vector<int>& spl,
                         const std::vector<unsigned char>& fl, bool least_force) {
    const size_t M = tau.size();
    std::vector<int> span(M, 0);
    int cur_s = -1, l = 0;
    double last_t = 0.0;
    for (size_t j = 0; j < M; ++j) {
        const int s = spl[j];
        const double* kn = T.knots.data() + T.knot_off[s];
        const int k = T.deg[s], lmax = T.ncoef[s] - 1;
        if (s != cur_s || tau[j] < last_t) { cur_s = s; l = k; }
        while (l < lmax && kn[l + 1] <= tau[j]) ++l;
        span[j] = l;
        last_t = tau[j];
    }
    int spread = 4;
    for (size_t j = 0; j < M; ++j) {
        if (!(fl[j] & 1)) continue;
        int lo = span[j], hi = lo;
        if (fl[j] & 2) { lo = std::min(lo, span[j - 1]); hi = std::max(hi, span[j - 1]); }
        if ((fl[j] & 4) && least_force) { lo = std::min(lo, span[j + 1]); hi = std::max(hi, span[j + 1]); }
        spread = std::max(spread, hi - lo + 4);
    }
    return spread;
}

}  // namespace mvus
