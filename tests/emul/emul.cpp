// TEST INFRASTRUCTURE: host build of the per-detection math in mvus_b200/csrc/ba_math.cuh
// (the very functions the CUDA kernels call), so tests can check the arithmetic against the
// oracle in the GPU-less container.  Never loaded by the product package.
#include <vector>
#include <cstring>
#include "../../mvus_b200/csrc/ba_math.cuh"
#include "../../mvus_b200/csrc/ba_tables.hpp"

using namespace mvus;

struct ArraySink {
    double* ju; double* jv;
    inline void put(int p, double a, double b) { ju[p] = a; jv[p] = b; }
};

extern "C" int emul_resjac(int nc, int opt_calib, int undist, int opt_sync, int opt_rs,
                           int motion_type, double motion_weight, const int64_t* cam_ptr,
                           const double* frame, const double* xr, const double* yr,
                           const double* height, const double* calib9, int S,
                           const double* interval, const int64_t* knot_ptr, const double* knots,
                           const int32_t* degree, const double* x, double* r, int32_t* span,
                           double* J, int32_t* mbase, double* mJ) {
    const int C = opt_calib ? 15 : 6, P = 3 + C + 12;
    HostSplineTables T;
    if (!build_spline_tables(S, interval, knot_ptr, knots, degree, (int64_t)nc * (3 + C), T)) return -1;
    SplineView sp{S, T.int_a.data(), T.int_b.data(), T.knots.data(), T.knot_off.data(), T.ncoef.data(),
                  T.deg.data(), T.ctrl_off.data(), T.xoff.data(), T.spanpoly.data(), T.span_t0.data(),
                  T.lut_off.data(), T.lut_n.data(), T.lut_t0.data(), T.lut_invh.data(), T.lut.data()};
    const int64_t N = cam_ptr[nc];
    FreeMask fm{opt_sync != 0, opt_rs != 0};
    int64_t row = 0;
    for (int i = 0; i < nc; ++i) {
        CamPrep c;
        cam_prep_one(x, i, nc, C, opt_calib != 0, calib9, height[i], c);
        const int64_t n0 = cam_ptr[i], Ni = cam_ptr[i + 1] - n0;
        for (int64_t d = n0; d < n0 + Ni; ++d) {
            double ju[32], jv[32], ru, rv, ou = xr[d], ov = yr[d];
            if (!opt_calib && undist) {
                double xn, yn;
                undistort5(xr[d], yr[d], c.K4, c.d, xn, yn);
                ou = c.K4[0] * xn + c.K4[2]; ov = c.K4[1] * yn + c.K4[3];
            }
            ArraySink sink{ju, jv};
            int sp_out;
            if (opt_calib) resjac_one<true, true>(c, undist != 0, fm, frame[d], xr[d], yr[d], ou, ov, sp, x, ru, rv, sp_out, sink);
            else resjac_one<false, true>(c, undist != 0, fm, frame[d], xr[d], yr[d], ou, ov, sp, x, ru, rv, sp_out, sink);
            r[row + (d - n0)] = ru;
            r[row + Ni + (d - n0)] = rv;
            span[d] = sp_out;
            for (int p = 0; p < P; ++p) { J[(int64_t)p * N + d] = ju[p]; J[(int64_t)(P + p) * N + d] = jv[p]; }
        }
        row += 2 * Ni;
    }
    std::vector<double> tauv; std::vector<int> splv; std::vector<unsigned char> flv;
    if (motion_type) build_motion_samples(T, tauv, splv, flv);
    const int64_t M = (int64_t)tauv.size();
    const double* tau = tauv.data(); const int* tau_spl = splv.data(); const unsigned char* tau_flag = flv.data();
    for (int64_t j = 0; j < M; ++j) {
        double rr, fa[3], fc[7]; int base;
        if (!motion_one<true>(motion_type, motion_weight, sp, x, tau, tau_spl, tau_flag, j, rr, base, fa, fc)) return -4;
        r[row + j] = rr; mbase[j] = base;
        for (int k = 0; k < 3; ++k) mJ[(int64_t)k * M + j] = fa[k];
        for (int k = 0; k < 7; ++k) mJ[(int64_t)(3 + k) * M + j] = fc[k];
    }
    return 0;
}

extern "C" int64_t emul_motion_count(int S, const double* interval, const int64_t* knot_ptr,
                                     const double* knots, const int32_t* degree) {
    HostSplineTables T;
    if (!build_spline_tables(S, interval, knot_ptr, knots, degree, 0, T)) return -1;
    std::vector<double> tauv; std::vector<int> splv; std::vector<unsigned char> flv;
    build_motion_samples(T, tauv, splv, flv);
    return (int64_t)tauv.size();
}
