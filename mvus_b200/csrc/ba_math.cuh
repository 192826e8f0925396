// Per-detection / per-sample mathematics of the mvus BA error function and its analytic
// Jacobian, written once as __host__ __device__ code: the CUDA kernels in mvus_ba.cu call
// it on the device; tests/emul/emul.cu compiles the same functions for the host so the
// arithmetic can be checked against the oracle in the GPU-less build container (that
// emulation library is test infrastructure and is never loaded by the product).
//
// Reference lines restated here (see SURVEY.md 8a "per-detection mathematics"):
//   time stamp        t = alpha (f + rho y/H) + beta                  common.py:125
//   observation       K * undistortPoints(raw; K, d), 5 iterations    common.py:126, 1147-1157
//   membership        (t >= a) xor (t >= b), last interval wins       util.py:103-106
//   spline            FITPACK splev (cubic or linear B-spline)        common.py:331
//   projection        x = P X / (P X)_z, P = K [R | t], R = Rodrigues common.py:1072-1079, 1127-1144
//   residual          |u - u_obs|, |v - v_obs|                        common.py:349-359
//   motion prior      F / KE on unit-step samples                     common.py:362-424, 959-1001
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define MV_HD __host__ __device__ __forceinline__
#else
#define MV_HD inline
#endif

namespace mvus {

// ---------------------------------------------------------------------------------------
// Views of the static problem data (device pointers on the device, host pointers in the
// host emulation).
struct SplineView {
    int S;
    const double* int_a;      // [S] interval starts   (Scene.spline['int'][0])
    const double* int_b;      // [S] interval ends     (Scene.spline['int'][1])
    const double* knots;      // concatenated knot vectors
    const int64_t* knot_off;  // [S]
    const int* ncoef;         // [S] coefficients per axis
    const int* deg;           // [S] 1 or 3
    const int64_t* ctrl_off;  // [S+1] first global control-point index of spline s
    const int64_t* xoff;      // [S] offset of spline s coefficients in x (cx | cy | cz)
    const double* spanpoly;   // [n_ctrl][16]: Taylor coefficients of the 4 basis functions of
                              //   span l (global index ctrl_off[s]+l): poly[m*4+d] * dt^d
    const double* span_t0;    // [n_ctrl] start knot of the span
    const int64_t* lut_off;   // [S] uniform-bucket span lookup
    const int* lut_n;         // [S]
    const double* lut_t0;     // [S]
    const double* lut_invh;   // [S]
    const int* lut;
};

// Per-camera quantities derived from x once per evaluation (cam_prep kernel).
struct CamPrep {
    double alpha, beta, rho, invH;
    double K4[4];    // fx fy cx cy
    double d[5];     // k1 k2 p1 p2 k3
    double R[9];     // row-major
    double T[3];
    double dR[27];   // dR/dw_k, k = 0..2, each row-major 3x3
};
static const int CAMPREP_DOUBLES = sizeof(CamPrep) / sizeof(double);

// ---------------------------------------------------------------------------------------
// Rodrigues and its derivative (cv2.Rodrigues semantics, common.py:1136/1140; derivative:
// dR/dw_k = (w_k [w]x + [w x (I - R) e_k]x) R / |w|^2, [e_k]x at w = 0).
MV_HD void rodrigues(const double w[3], double R[9], double dR[27]) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double th = sqrt(th2);
    if (th < 2.220446049250313e-16) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        for (int i = 0; i < 27; ++i) dR[i] = 0.0;
        dR[0 * 9 + 5] = -1; dR[0 * 9 + 7] = 1;     // [e_x]x
        dR[1 * 9 + 2] = 1;  dR[1 * 9 + 6] = -1;    // [e_y]x
        dR[2 * 9 + 1] = -1; dR[2 * 9 + 3] = 1;     // [e_z]x
        return;
    }
    const double c = cos(th), s = sin(th), c1 = 1.0 - c, it = 1.0 / th;
    const double kx = w[0] * it, ky = w[1] * it, kz = w[2] * it;
    R[0] = c + c1 * kx * kx;      R[1] = c1 * kx * ky - s * kz; R[2] = c1 * kx * kz + s * ky;
    R[3] = c1 * kx * ky + s * kz; R[4] = c + c1 * ky * ky;      R[5] = c1 * ky * kz - s * kx;
    R[6] = c1 * kx * kz - s * ky; R[7] = c1 * ky * kz + s * kx; R[8] = c + c1 * kz * kz;
    const double ith2 = 1.0 / th2;
    for (int k = 0; k < 3; ++k) {
        // q = (I - R) e_k ; p = w x q ; M = w_k [w]x + [p]x ; dR_k = M R / th2
        double q[3] = {-R[0 * 3 + k], -R[1 * 3 + k], -R[2 * 3 + k]};
        q[k] += 1.0;
        const double p0 = w[1] * q[2] - w[2] * q[1];
        const double p1 = w[2] * q[0] - w[0] * q[2];
        const double p2 = w[0] * q[1] - w[1] * q[0];
        const double a0 = w[k] * w[0] + p0, a1 = w[k] * w[1] + p1, a2 = w[k] * w[2] + p2;
        // M = [a]x
        const double M[9] = {0, -a2, a1, a2, 0, -a0, -a1, a0, 0};
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc)
                dR[k * 9 + r * 3 + cc] = (M[r * 3 + 0] * R[0 * 3 + cc] + M[r * 3 + 1] * R[1 * 3 + cc] +
                                          M[r * 3 + 2] * R[2 * 3 + cc]) * ith2;
    }
}

// Build CamPrep for camera i from x (reference layout) and the constant calibration.
MV_HD void cam_prep_one(const double* x, int i, int nc, int C, bool calib, const double* calib9,
                        double height, CamPrep& c) {
    c.alpha = x[i];
    c.beta = x[nc + i];
    c.rho = x[2 * nc + i];
    c.invH = 1.0 / height;
    const double* cam = x + 3 * nc + (int64_t)i * C;
    const double* w;
    if (calib) {
        for (int k = 0; k < 4; ++k) c.K4[k] = cam[k];
        w = cam + 4;
        for (int k = 0; k < 3; ++k) c.T[k] = cam[7 + k];
        for (int k = 0; k < 5; ++k) c.d[k] = cam[10 + k];
    } else {
        for (int k = 0; k < 4; ++k) c.K4[k] = calib9[i * 9 + k];
        for (int k = 0; k < 5; ++k) c.d[k] = calib9[i * 9 + 4 + k];
        w = cam;
        for (int k = 0; k < 3; ++k) c.T[k] = cam[3 + k];
    }
    rodrigues(w, c.R, c.dR);
}

// ---------------------------------------------------------------------------------------
// cv2.undistortPoints for (k1,k2,p1,p2,k3): 5 fixed-point iterations (common.py:1154).
MV_HD void undistort5(double x, double y, const double K4[4], const double d[5], double& xn,
                      double& yn) {
    const double x0 = (x - K4[2]) / K4[0], y0 = (y - K4[3]) / K4[1];
    xn = x0; yn = y0;
    for (int it = 0; it < 5; ++it) {
        const double r2 = xn * xn + yn * yn;
        const double ic = 1.0 / (1.0 + ((d[4] * r2 + d[1]) * r2 + d[0]) * r2);
        const double dx = 2.0 * d[2] * xn * yn + d[3] * (r2 + 2.0 * xn * xn);
        const double dy = d[2] * (r2 + 2.0 * yn * yn) + 2.0 * d[3] * xn * yn;
        xn = (x0 - dx) * ic;
        yn = (y0 - dy) * ic;
    }
}

// Same with forward-mode derivatives w.r.t. p = (fx, fy, cx, cy, k1, k2, p1, p2, k3) of
// u_obs = fx xn + cx, v_obs = fy yn + cy.
MV_HD void undistort5_jac(double x, double y, const double K4[4], const double d[5], double& uo,
                          double& vo, double du[9], double dv[9]) {
    const double fx = K4[0], fy = K4[1];
    const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
    const double x0 = (x - K4[2]) / fx, y0 = (y - K4[3]) / fy;
    double dx0[9], dy0[9], dxn[9], dyn[9];
    for (int q = 0; q < 9; ++q) { dx0[q] = 0; dy0[q] = 0; }
    dx0[0] = -x0 / fx; dx0[2] = -1.0 / fx;
    dy0[1] = -y0 / fy; dy0[3] = -1.0 / fy;
    for (int q = 0; q < 9; ++q) { dxn[q] = dx0[q]; dyn[q] = dy0[q]; }
    double xn = x0, yn = y0;
    for (int it = 0; it < 5; ++it) {
        const double r2 = xn * xn + yn * yn;
        const double dpoly = (3.0 * k3 * r2 + 2.0 * k2) * r2 + k1;
        const double ic = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2);
        const double dX = 2.0 * p1 * xn * yn + p2 * (r2 + 2.0 * xn * xn);
        const double dY = p1 * (r2 + 2.0 * yn * yn) + 2.0 * p2 * xn * yn;
        const double ax = x0 - dX, ay = y0 - dY;
        for (int q = 0; q < 9; ++q) {
            const double dr2 = 2.0 * (xn * dxn[q] + yn * dyn[q]);
            double dden = dpoly * dr2;
            if (q == 4) dden += r2;
            if (q == 5) dden += r2 * r2;
            if (q == 8) dden += r2 * r2 * r2;
            const double dic = -dden * ic * ic;
            const double cross = dxn[q] * yn + xn * dyn[q];
            double ddX = 2.0 * p1 * cross + p2 * (dr2 + 4.0 * xn * dxn[q]);
            double ddY = p1 * (dr2 + 4.0 * yn * dyn[q]) + 2.0 * p2 * cross;
            if (q == 6) { ddX += 2.0 * xn * yn; ddY += r2 + 2.0 * yn * yn; }
            if (q == 7) { ddX += r2 + 2.0 * xn * xn; ddY += 2.0 * xn * yn; }
            dxn[q] = (dx0[q] - ddX) * ic + ax * dic;
            dyn[q] = (dy0[q] - ddY) * ic + ay * dic;
        }
        xn = ax * ic;
        yn = ay * ic;
    }
    uo = fx * xn + K4[2];
    vo = fy * yn + K4[3];
    for (int q = 0; q < 9; ++q) { du[q] = fx * dxn[q]; dv[q] = fy * dyn[q]; }
    du[0] += xn; du[2] += 1.0;
    dv[1] += yn; dv[3] += 1.0;
}

// ---------------------------------------------------------------------------------------
// Interval membership (util.sampling belong=True): 0-based spline id or -1.
MV_HD int find_interval(const SplineView& sp, double t) {
    int s_hit = -1;
    for (int s = 0; s < sp.S; ++s) {
        const bool ge_a = (t - sp.int_a[s]) >= 0.0, ge_b = (t - sp.int_b[s]) >= 0.0;
        if (ge_a != ge_b) s_hit = s;
    }
    return s_hit;
}

// Knot span l with knots[l] <= t < knots[l+1], clamped to [k, ncoef-1] (FITPACK fpbspl).
MV_HD int find_span(const SplineView& sp, int s, double t) {
    const double* kn = sp.knots + sp.knot_off[s];
    const int k = sp.deg[s], lmax = sp.ncoef[s] - 1;
    int b = (int)floor((t - sp.lut_t0[s]) * sp.lut_invh[s]);
    const int nb = sp.lut_n[s];
    b = b < 0 ? 0 : (b >= nb ? nb - 1 : b);
    int l = sp.lut[sp.lut_off[s] + b];
    while (l < lmax && t >= kn[l + 1]) ++l;
    while (l > k && t < kn[l]) --l;
    return l;
}

// Values (and first derivatives) of the 4 basis-function slots of global span g at time t.
// Slot m belongs to control point (l - 3 + m); for a linear spline slots 0,1 are zero.
template <bool DERIV>
MV_HD void span_basis(const SplineView& sp, int64_t g, double t, double B[4], double dB[4]) {
    const double* p = sp.spanpoly + g * 16;
    const double dt = t - sp.span_t0[g];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const double c0 = p[m * 4 + 0], c1 = p[m * 4 + 1], c2 = p[m * 4 + 2], c3 = p[m * 4 + 3];
        B[m] = ((c3 * dt + c2) * dt + c1) * dt + c0;
        if (DERIV) dB[m] = (3.0 * c3 * dt + 2.0 * c2) * dt + c1;
    }
}

// Spline position (and time derivative) from x.  Returns global span index.
template <bool DERIV>
MV_HD int64_t spline_eval(const SplineView& sp, const double* x, int s, double t, double X[3],
                          double dX[3], double B[4]) {
    const int l = find_span(sp, s, t);
    const int64_t g = sp.ctrl_off[s] + l;
    double dB[4];
    span_basis<DERIV>(sp, g, t, B, dB);
    const int nco = sp.ncoef[s];
    const double* cx = x + sp.xoff[s];
    X[0] = X[1] = X[2] = 0.0;
    if (DERIV) dX[0] = dX[1] = dX[2] = 0.0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        int j = l - 3 + m;
        j = j < 0 ? 0 : j;              // slot unused (zero basis) when clamped
        const double c0 = cx[j], c1 = cx[nco + j], c2 = cx[2 * nco + j];
        X[0] += B[m] * c0; X[1] += B[m] * c1; X[2] += B[m] * c2;
        if (DERIV) { dX[0] += dB[m] * c0; dX[1] += dB[m] * c1; dX[2] += dB[m] * c2; }
    }
    return g;
}

// ---------------------------------------------------------------------------------------
// One detection: residual (abs) and, if WANTJ, the sign-applied Jacobian block row.
// Compact column order: 0 alpha, 1 beta, 2 rho, 3..3+C-1 camera vector (reference order),
// then 4 control-point slots x (x,y,z).  The sink receives (column, du, dv).
//   obs_u/obs_v: pre-undistorted observation (used when !CALIB); xr/yr raw pixel.
struct FreeMask { bool sync, rs; };

template <bool CALIB, bool WANTJ, class Sink>
MV_HD void resjac_one(const CamPrep& c, bool undist, FreeMask fm, double f, double xr, double yr,
                      double obs_u, double obs_v, const SplineView& sp, const double* x,
                      double& ru, double& rv, int& span_out, Sink& sink) {
    const int C = CALIB ? 15 : 6;
    const double yH = yr * c.invH;
    const double tau = f + c.rho * yH;
    const double t = c.alpha * tau + c.beta;
    const int s = find_interval(sp, t);
    if (s < 0) {
        ru = 0.0; rv = 0.0; span_out = -1;
        if (WANTJ) for (int p = 0; p < 3 + C + 12; ++p) sink.put(p, 0.0, 0.0);
        return;
    }
    double X[3], dX[3], B[4];
    const int64_t g = spline_eval<WANTJ>(sp, x, s, t, X, dX, B);
    span_out = (int)g;
    const double* R = c.R;
    const double Xc0 = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + c.T[0];
    const double Xc1 = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + c.T[1];
    const double Xc2 = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + c.T[2];
    const double iz = 1.0 / Xc2;
    const double fx = c.K4[0], fy = c.K4[1];
    double uo = obs_u, vo = obs_v, duo[9], dvo[9];
    if (CALIB) {
        if (undist) {
            if (WANTJ) undistort5_jac(xr, yr, c.K4, c.d, uo, vo, duo, dvo);
            else {
                double xn, yn;
                undistort5(xr, yr, c.K4, c.d, xn, yn);
                uo = fx * xn + c.K4[2]; vo = fy * yn + c.K4[3];
            }
        } else {
            uo = xr; vo = yr;
            if (WANTJ) for (int q = 0; q < 9; ++q) { duo[q] = 0.0; dvo[q] = 0.0; }
        }
    }
    const double xz = Xc0 * iz, yz = Xc1 * iz;
    const double eu = fx * xz + c.K4[2] - uo;
    const double ev = fy * yz + c.K4[3] - vo;
    const double su = eu < 0.0 ? -1.0 : 1.0, sv = ev < 0.0 ? -1.0 : 1.0;
    ru = su * eu; rv = sv * ev;
    if (!WANTJ) return;
    // G = d(u,v)/dXc, sign-applied
    const double gu0 = su * fx * iz, gu2 = -su * fx * xz * iz;
    const double gv1 = sv * fy * iz, gv2 = -sv * fy * yz * iz;
    // (G R): rows
    const double GRu[3] = {gu0 * R[0] + gu2 * R[6], gu0 * R[1] + gu2 * R[7], gu0 * R[2] + gu2 * R[8]};
    const double GRv[3] = {gv1 * R[3] + gv2 * R[6], gv1 * R[4] + gv2 * R[7], gv1 * R[5] + gv2 * R[8]};
    const double vu = GRu[0] * dX[0] + GRu[1] * dX[1] + GRu[2] * dX[2];
    const double vv = GRv[0] * dX[0] + GRv[1] * dX[1] + GRv[2] * dX[2];
    const double ms = fm.sync ? 1.0 : 0.0, mr = fm.rs ? 1.0 : 0.0;
    sink.put(0, ms * vu * tau, ms * vv * tau);
    sink.put(1, ms * vu, ms * vv);
    sink.put(2, mr * vu * c.alpha * yH, mr * vv * c.alpha * yH);
    const int ro = CALIB ? 4 : 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double* D = c.dR + k * 9;
        const double q0 = D[0] * X[0] + D[1] * X[1] + D[2] * X[2];
        const double q1 = D[3] * X[0] + D[4] * X[1] + D[5] * X[2];
        const double q2 = D[6] * X[0] + D[7] * X[1] + D[8] * X[2];
        sink.put(3 + ro + k, gu0 * q0 + gu2 * q2, gv1 * q1 + gv2 * q2);
    }
    sink.put(3 + ro + 3, gu0, 0.0);
    sink.put(3 + ro + 4, 0.0, gv1);
    sink.put(3 + ro + 5, gu2, gv2);
    if (CALIB) {
        // d e / d (fx, fy, cx, cy): projection part minus observation part; d e / d d_k
        sink.put(3 + 0, su * (xz - duo[0]), sv * (-dvo[0]));
        sink.put(3 + 1, su * (-duo[1]), sv * (yz - dvo[1]));
        sink.put(3 + 2, su * (1.0 - duo[2]), sv * (-dvo[2]));
        sink.put(3 + 3, su * (-duo[3]), sv * (1.0 - dvo[3]));
#pragma unroll
        for (int k = 0; k < 5; ++k) sink.put(3 + 10 + k, -su * duo[4 + k], -sv * dvo[4 + k]);
    }
    const int cb = 3 + C;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) sink.put(cb + m * 3 + ax, GRu[ax] * B[m], GRv[ax] * B[m]);
    }
}

// ---------------------------------------------------------------------------------------
// Motion-prior rows (common.py:959-1001 on the unit-step samples of common.py:273-301).
// flags: bit0 = sample is in a group (a <= tau < b), bit1 = has a predecessor in the same
// group, bit2 = has a successor in the same group.
// Output: r (>= 0), base control point (or -1), axis factors fa[3], control factors fc[7]
// with d r / d C_{base+k, ax} = fa[ax] * fc[k].  Returns false if the row needs more than 7
// consecutive control points (unsupported knot density).
template <bool WANTJ>
MV_HD bool motion_one(int type, double w, const SplineView& sp, const double* x, const double* tau,
                      const int* tau_spl, const unsigned char* flags, int64_t j, double& r,
                      int& base, double fa[3], double fc[7]) {
    const double eps = 1e-20;
    r = 0.0; base = -1;
    if (WANTJ) { fa[0] = fa[1] = fa[2] = 0.0; for (int k = 0; k < 7; ++k) fc[k] = 0.0; }
    const unsigned char fl = flags[j];
    if (!(fl & 1)) return true;
    const int s = tau_spl[j];
    if (type == 2) {            // KE: rows j with a predecessor
        if (!(fl & 2)) return true;
        double P0[3], Pm[3], dmy[3], B0[4], Bm[4];
        const int64_t g0 = spline_eval<false>(sp, x, s, tau[j], P0, dmy, B0);
        const int64_t gm = spline_eval<false>(sp, x, s, tau[j - 1], Pm, dmy, Bm);
        const double dt = tau[j] - tau[j - 1];
        const double idt = 1.0 / (dt + eps);
        const double aw = fabs(w);
        double sum = 0.0, v[3];
        for (int ax = 0; ax < 3; ++ax) { v[ax] = (P0[ax] - Pm[ax]) * idt; sum += fabs(aw * 0.5 * (v[ax] * v[ax] * dt)); }
        r = sum;
        if (WANTJ) {
            const int64_t lo = (gm < g0 ? gm : g0) - 3;
            const int64_t lo_c = lo < sp.ctrl_off[s] ? sp.ctrl_off[s] : lo;
            const int64_t hi = (gm < g0 ? g0 : gm);
            if (hi - lo_c + 1 > 7) return false;
            base = (int)lo_c;
            const double sg = dt < 0 ? -1.0 : 1.0;
            for (int ax = 0; ax < 3; ++ax) fa[ax] = aw * v[ax] * dt * idt * sg;
            for (int m = 0; m < 4; ++m) {
                const int64_t c0 = g0 - 3 + m, cm = gm - 3 + m;
                if (c0 >= lo_c) fc[c0 - lo_c] += B0[m];
                if (cm >= lo_c) fc[cm - lo_c] -= Bm[m];
            }
        }
        return true;
    }
    // F: rows j with predecessor and successor
    if ((fl & 6) != 6) return true;
    double P0[3], Pm[3], Pp[3], dmy[3], B0[4], Bm[4], Bp[4];
    const int64_t g0 = spline_eval<false>(sp, x, s, tau[j], P0, dmy, B0);
    const int64_t gm = spline_eval<false>(sp, x, s, tau[j - 1], Pm, dmy, Bm);
    const int64_t gp = spline_eval<false>(sp, x, s, tau[j + 1], Pp, dmy, Bp);
    const double dt1 = tau[j] - tau[j - 1], dt2 = tau[j + 1] - tau[j], dt3 = dt1 + dt2;
    const double i1 = 1.0 / (dt1 + eps), i2 = 1.0 / (dt2 + eps), i3 = 1.0 / (dt3 + eps);
    double sum = 0.0, sg[3];
    for (int ax = 0; ax < 3; ++ax) {
        const double v1 = (P0[ax] - Pm[ax]) * i1, v2 = (Pp[ax] - P0[ax]) * i2;
        const double a = w * ((v2 - v1) * i3 * dt3);
        sg[ax] = a < 0.0 ? -1.0 : 1.0;
        sum += sg[ax] * a;
    }
    r = sum;
    if (WANTJ) {
        int64_t lo = gm < g0 ? gm : g0; lo = gp < lo ? gp : lo; lo -= 3;
        int64_t hi = gm > g0 ? gm : g0; hi = gp > hi ? gp : hi;
        const int64_t lo_c = lo < sp.ctrl_off[s] ? sp.ctrl_off[s] : lo;
        if (hi - lo_c + 1 > 7) return false;
        base = (int)lo_c;
        for (int ax = 0; ax < 3; ++ax) fa[ax] = w * sg[ax] * dt3 * i3;
        for (int m = 0; m < 4; ++m) {
            const int64_t c0 = g0 - 3 + m, cm = gm - 3 + m, cp = gp - 3 + m;
            if (cp >= lo_c) fc[cp - lo_c] += Bp[m] * i2;
            if (c0 >= lo_c) fc[c0 - lo_c] -= B0[m] * (i1 + i2);
            if (cm >= lo_c) fc[cm - lo_c] += Bm[m] * i1;
        }
    }
    return true;
}

}  // namespace mvus
