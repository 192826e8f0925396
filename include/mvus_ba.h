/*
 * mvus_ba.h -- C ABI of the B200-native bundle-adjustment path for mvus.
 *
 * The reference (CenekAlbl/mvus) is pure Python and has NO foreign-function interface;
 * this header is the boundary a maintainer binds from Python (ctypes, see
 * INTEGRATION.md) to replace the body of
 *
 *     Scene.BA(numCam, max_iter, rs, motion_prior, motion_reg, motion_weights, norm,
 *              rs_bounds)                     multiviewunsynch/reconstruction/common.py:441-697
 *
 * Every entry point cites the reference lines it replaces.  Conventions:
 *   - plain C, FP64 throughout (the reference is NumPy float64 end to end);
 *   - all array arguments are HOST pointers unless the name ends in _dev; the handle owns
 *     all device memory, the caller owns all host memory;
 *   - every function returns 0 on success and a negative mvus_status on error;
 *     mvus_ba_last_error() gives the message.  No exceptions cross the ABI;
 *   - one handle = one CUDA device + one stream; a handle is not thread-safe, distinct
 *     handles are independent;
 *   - there is NO CPU fallback: without a usable CUDA device mvus_ba_create fails with
 *     MVUS_ERR_CUDA.
 *
 * Parameter vector x (length n), exactly the reference layout (common.py:616-650):
 *   [ alpha(nc) | beta(nc) | rho(nc) | cam_0 .. cam_{nc-1} | spline_0: cx|cy|cz | spline_1 .. ]
 *   cam_i = [rvec(3), t(3)]                         (6)   opt_calib = 0   common.py:1124
 *         = [fx, fy, cx, cy, rvec(3), t(3), d(5)]   (15)  opt_calib = 1   common.py:1122
 * Residual vector r (length m), exactly the reference order (common.py:476-487, 359):
 *   for each camera i: [ |e_u| (N_i) , |e_v| (N_i) ], then the M motion-prior rows.
 */
#ifndef MVUS_BA_H
#define MVUS_BA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mvus_ba_ctx* mvus_ba_handle;

enum mvus_status {
    MVUS_OK = 0,
    MVUS_ERR_ARG = -1,       /* bad argument / call order */
    MVUS_ERR_CUDA = -2,      /* CUDA runtime error or no device */
    MVUS_ERR_NONFINITE = -3, /* "Residuals are not finite in the initial point." (scipy least_squares.py:945) */
    MVUS_ERR_UNSUPPORTED = -4,
    MVUS_ERR_NCCL = -5
};

enum mvus_motion_type { MVUS_MOTION_NONE = 0, MVUS_MOTION_F = 1, MVUS_MOTION_KE = 2 };

/* Problem description: what Scene.BA reads from its arguments and self.settings
 * (common.py:441, 454-460, 512-525, 621, 655-662; main.py:49-52). */
typedef struct mvus_ba_desc {
    int32_t num_cams;        /* numCam                                                   */
    int32_t opt_calib;       /* settings['opt_calib']  -> 15 instead of 6 camera unknowns */
    int32_t undist_points;   /* settings['undist_points'] (common.py:126)                */
    int32_t opt_sync;        /* settings.get('opt_sync', True): alpha/beta free (512-515) */
    int32_t opt_rs;          /* BA(rs=...): rho free (518-521)                            */
    int32_t rs_bounds;       /* BA(rs_bounds=...): rho in [0,1] (655-660)                 */
    int32_t motion_type;     /* mvus_motion_type; NONE when motion_reg is False           */
    int32_t device;          /* CUDA device ordinal                                       */
    double  motion_weight;   /* motion_weights (483-485)                                  */
    int32_t max_nfev;        /* max_iter -> least_squares(max_nfev=...) (670)             */
    int32_t solver_chunk;    /* super-blocks per chunk of the pre-reduction in front of the cyclic reduction;
                                0 = chosen from the problem size, 1 = cyclic reduction only */
    double  ftol, xtol, gtol;/* SciPy defaults 1e-8 / reference xtol=1e-12 / 1e-8 (670)   */
} mvus_ba_desc;

/* What scipy's OptimizeResult carries back (common.py:697) plus timing. */
typedef struct mvus_ba_stats {
    double  cost0;           /* 0.5*|r(x0)|^2                                             */
    double  cost;            /* 0.5*|r(x*)|^2                                             */
    double  optimality;      /* |J^T r|_inf at the last point a Jacobian was evaluated at  */
    double  lambda;          /* final LM damping                                          */
    int32_t nfev;            /* residual evaluations                                      */
    int32_t njev;            /* Jacobian evaluations                                      */
    int32_t status;          /* 0 max_nfev, 1 gtol, 2 ftol, 3 xtol, 4 ftol&xtol (scipy codes); -1 failure */
    int32_t lm_iterations;   /* accepted + rejected linear solves                         */
    double  ms_total;        /* device time of the whole solve (CUDA events)              */
    double  ms_resjac;       /* summed device time in K1/K1m with Jacobian                */
    double  ms_accum;        /* summed device time in K2/K2m                              */
    double  ms_solve;        /* summed device time in the Schur/Cholesky step             */
    double  ms_trial;        /* summed device time in residual-only evaluations           */
    int32_t launches;        /* kernels launched by this library during the solve         */
    int32_t n_resjac;        /* number of K1(with J) launches timed in ms_resjac          */
    double  ms_syrk;         /* of ms_solve: the Schur-complement SYRK (+ rhs GEMV)       */
    double  ms_bcr;          /* of ms_solve: cyclic-reduction levels (elimination)        */
    double  ms_reduce;       /* of ms_accum: multi-GPU exchange of the normal equations   */
    double  ms_k2;           /* of ms_accum: K2 + K2m (with their memsets)                */
} mvus_ba_stats;

const char* mvus_ba_version(void);

/* Create / destroy a problem handle.  Replaces nothing in the reference by itself; holds
 * what common.py:612-665 derives from the Scene. */
int  mvus_ba_create(const mvus_ba_desc* desc, mvus_ba_handle* out);
void mvus_ba_destroy(mvus_ba_handle h);
const char* mvus_ba_last_error(mvus_ba_handle h);   /* h may be NULL: last create error */

/* Detections of the nc optimised cameras, concatenated in sequence order.
 * cam_ptr[nc+1]: offsets; frame/x/y: rows 0/1/2 of Scene.detections[i] (common.py:1190,
 * sorted by frame); height[i] = cameras[i].resolution[1] (common.py:125).
 * calib: nc x 9 = (fx, fy, cx, cy, k1, k2, p1, p2, k3) -- used as constants when
 * opt_calib = 0 (common.py:126, 1147-1157) and ignored otherwise. */
int mvus_ba_set_detections(mvus_ba_handle h, const int64_t* cam_ptr, const double* frame,
                           const double* x, const double* y, const double* height,
                           const double* calib);

/* Same, without a host-side concatenation: one (frame, x, y) row-pointer triple per camera
 * (rows 0/1/2 of Scene.detections[i] when they are contiguous), count[i] detections each. */
int mvus_ba_set_detections_rows(mvus_ba_handle h, const int64_t* count, const double* const* frame,
                                const double* const* x, const double* const* y, const double* height,
                                const double* calib);

/* Splines: Scene.spline['tck'] / ['int'] (common.py:224-270).  interval: 2*S doubles
 * (row 0 = starts, row 1 = ends, i.e. the 2 x S array flattened); knot_ptr[S+1] into
 * knots[]; degree[s] in {1,3} (common.py:247, 267).  Coefficients travel inside x. */
int mvus_ba_set_splines(mvus_ba_handle h, int32_t num_splines, const double* interval,
                        const int64_t* knot_ptr, const double* knots, const int32_t* degree);

/* Problem sizes after the two setters: n = len(x), m = len(r), N detections, M motion
 * rows, P = 3 + C + 12 compact Jacobian columns per detection row. */
int mvus_ba_dims(mvus_ba_handle h, int64_t* n, int64_t* m, int64_t* N, int64_t* M, int32_t* P);

/* r = error_BA(x)   (common.py:448-487).  x, r host pointers. */
int mvus_ba_residual(mvus_ba_handle h, const double* x, double* r);

/* r and the analytic Jacobian of error_BA at x -- replaces the (1 + n_groups) finite
 * difference evaluations scipy does per Jacobian (scipy/optimize/_numdiff.py:770-895) and
 * the pattern of jac_BA (common.py:490-610).  Compact block-row form:
 *   span[N]     : global index of the LAST active control point of the detection
 *                 (control points are numbered consecutively over the splines), -1 if the
 *                 detection is covered by no interval (row is zero, common.py:565-566);
 *   J[2*P*N]    : column planes, J[p*N + d] = d|e_u|(d)/d q_p and J[(P+p)*N + d] for e_v,
 *                 q = (alpha_i, beta_i, rho_i, cam_i (C), then 4 control points
 *                 span-3..span, each (x, y, z));
 *   mbase[M]    : first control point touched by motion row j, -1 if the row is zero;
 *   mJ[10*M]    : planes; axis factors a[0..2] then control-point factors c[0..6]:
 *                 d r_j / d C_{mbase+k, axis} = a[axis] * c[k].
 * Any of span/J/mbase/mJ may be NULL. */
int mvus_ba_residual_jacobian(mvus_ba_handle h, const double* x, double* r, int32_t* span,
                              double* J, int32_t* mbase, double* mJ);

/* Full solve: replaces least_squares(error_BA, x0, jac_sparsity=A, tr_solver='lsmr',
 * xtol=1e-12, max_nfev=max_iter, bounds=...) (common.py:670) -- Levenberg-Marquardt with
 * exact steps from the Schur-reduced normal equations.  x0 in, x* out (host, length n);
 * r_out (length m, may be NULL) receives error_BA(x*). */
int mvus_ba_solve(mvus_ba_handle h, const double* x0, double* x_out, double* r_out,
                  mvus_ba_stats* stats);

/* detections_global of the optimised cameras at parameters x (common.py:105-127, 695):
 * t = alpha (f + rho y/H) + beta and the (undistorted) observation.  out[3*N]: for camera i
 * a contiguous 3 x N_i row-major block [t; u; v] starting at 3*cam_ptr[i] -- exactly the
 * array the reference stores in Scene.detections_global[i] (np.vstack, common.py:127). */
int mvus_ba_detections_global(mvus_ba_handle h, const double* x, double* out);

/* Scene.visible (compute_visibility, common.py:427-438) of the optimised cameras at parameters
 * x: 1-based spline-interval id per detection, 0 = in no interval (util.sampling belong=True,
 * util.py:103-106).  visible[N] int64, concatenated in camera order. */
int mvus_ba_visibility(mvus_ba_handle h, const double* x, int64_t* visible);

/* Device memory: handles allocate from a memory pool PRIVATE to this library (one per device; the
 * process-wide default pool is not touched).  Destroyed handles leave their memory cached in that
 * pool for the next handle (config 4 needs ~35 GB; main.py makes two BA calls per camera);
 * mvus_ba_trim returns everything above keep_bytes to the driver.  The NCCL communicator
 * (mvus_ba_comm_init) and the page-locked staging buffers live until the process exits. */
int mvus_ba_trim(int32_t device, uint64_t keep_bytes);

/* Page-locked host buffers for large outputs (device->host copies at PCIe speed). */
void* mvus_ba_host_alloc(size_t bytes);
void  mvus_ba_host_free(void* p);

/* Bookkeeping output of Scene.all_detect_to_traj (common.py:887-944) at parameters x:
 * `global_traj` = every detection of the optimised cameras whose global time stamp lies inside
 * a spline interval (closed ends, common.py:292), sorted by time stamp.  out: 7 x n_out
 * row-major (rows: running index, camera id taken from cam_ids[nc], frame id, time stamp,
 * X, Y, Z of the spline at that time); the buffer must hold 7*N doubles.
 * gd_out (may be NULL): `global_detections` (common.py:927), 3 x N row-major = camera id, frame
 * id, global time stamp of every detection in concatenation order. */
int mvus_ba_global_traj(mvus_ba_handle h, const double* x, const int32_t* cam_ids, int64_t* n_out,
                        double* out, double* gd_out);

/* Scene.spline_to_traj (common.py:273-301): evaluate the splines (coefficients inside x) at the
 * ascending times t[n]; a time is kept iff it lies inside a spline interval (closed ends,
 * common.py:292).  out: 4 x n_out row-major (time, X, Y, Z); the buffer must hold 4*n doubles.
 * Needs mvus_ba_set_splines only (no detections). */
int mvus_ba_spline_to_traj(mvus_ba_handle h, const double* x, const double* t, int64_t n,
                           int64_t* n_out, double* out);

/* Ground-truth alignment (analysis/compare_gt.py:35-70, 112-126; thirdparty/transformation.py:869-975
 * affine_matrix_from_points(shear=False, scale=True)): nshift independent similarity fits between the
 * splines of the handle (coefficients inside x), evaluated at tau[j] + shift[b], and the fixed points
 * pts (3 x n row-major).  A point takes part iff its shifted time lies in a spline interval by the rule
 * of util.sampling (a <= t < b, util.py:105).  spline_is_src != 0: the transform maps the spline points
 * onto pts (fine stage, error_fn of compare_gt.optimize); 0: pts onto the spline points (coarse search,
 * where the spline is the interpolating spline of the ground truth, util.match_overlap).
 * Outputs per shift: mean_err (mean point distance after the fit; +inf when fewer than 3 points take
 * part, where the reference raises), count, M (4 x 4 row-major).  err (may be NULL): the n point
 * distances of shift number `want`, 0 where the point takes no part.  Needs mvus_ba_set_splines only. */
int mvus_ba_align(mvus_ba_handle h, const double* x, int64_t n, const double* tau, const double* pts,
                  int32_t nshift, const double* shift, int32_t spline_is_src, int32_t want,
                  double* mean_err, int64_t* count, double* M, double* err);

/* Scene.BA(motion_prior=True): the discrete-trajectory mode (common.py:466-467, 527-550, 587-605, 631-634,
 * 681-687).  Unknowns = camera side + the G points of global_traj (common.py:887-944); the splines are
 * constants; motion rows of error_motion(motion_prior=True) (common.py:386-403) tie consecutive points of a
 * spline interval together, with time stamps that follow alpha/beta/rho of the camera that saw each point.
 * Two handles: hs = the ordinary handle of the flight (detections + splines, motion_type NONE), hp = a handle
 * with the same cameras, NO detections and ONE pseudo-spline of degree 1 with G coefficients (it only sizes the
 * block-tridiagonal solver for G points; the LM controls -- max_nfev, tolerances, rs_bounds -- are hp's).
 *   mvus_ba_points_set   per point: camera slot (position in sequence[:numCam]), frame id and raw y / image
 *                        height of its detection.
 *   x layout of hp       [alpha, beta, rho, camera vectors | X_0..X_G-1 | Y_0.. | Z_0..]  (the reference's
 *                        vector interleaves the points, common.py:631-634; the binding converts)
 *   xs0                  a parameter vector of hs (reference layout): its spline coefficients are the constants
 *   r_out                2N reprojection rows in the reference's order, then G motion rows in global_traj order
 *                        (row j = the triple centred at point j for F, the pair ending at j for KE)
 * A motion row may span at most 4 consecutive points (MVUS_ERR_UNSUPPORTED otherwise).  One GPU. */
int mvus_ba_points_set(mvus_ba_handle hp, int64_t G, const int32_t* cam_slot, const double* frame,
                       const double* y_over_height);
int mvus_ba_solve_points(mvus_ba_handle hs, mvus_ba_handle hp, int32_t motion_type, double motion_weight,
                         const double* xs0, const double* x0, double* x_out, double* r_out,
                         mvus_ba_stats* stats);
/* diagnostics for the parity tests: residual vector, gradient J^T r (hp layout) and cost at x */
int mvus_ba_points_eval(mvus_ba_handle hs, mvus_ba_handle hp, int32_t motion_type, double motion_weight,
                        const double* xs0, const double* x, double* r_out, double* g_out, double* cost);

/* Diagnostics used by the parity tests: the normal equations K2 assembles at x.
 *   A    [nc*Pc*Pc]  camera diagonal blocks (Pc = 3 + C), row-major per camera
 *   g    [n]         J^T r in the reference's x layout
 *   Hss  : spline-spline block returned as a dense band: for control points i <= j <
 *          i + bw (bw = *band_ctrl), Hss[((i*bw)+(j-i))*9 + a*3 + b]
 *   Hcs  [nc*Pc * 3*n_ctrl] camera x spline coupling, row-major, spline column = 3*j+axis
 * Any output may be NULL. */
int mvus_ba_normal_equations(mvus_ba_handle h, const double* x, double* A, double* g,
                             double* Hss, int32_t* band_ctrl, double* Hcs, double* cost);

/* ---- smoothing-spline fit (Scene.traj_to_spline, common.py:224-270; triangulate's refit :754-815) ----
 * The per-data-point arithmetic of scipy.interpolate.splprep (FITPACK parcur/fppara) on the device; the
 * knot-placement / smoothing-parameter decisions are host logic (mvus_b200/splfit.py).  A handle holds the
 * data of ONE interval: u[m] ascending parameter values (time stamps), x[idim][m] row-major, degree k (1 or 3). */
typedef struct mvus_spl_ctx* mvus_spl_handle;
int mvus_ba_spl_create(int32_t device, int64_t m, int32_t idim, int32_t k, const double* u, const double* x,
                       mvus_spl_handle* out);
void mvus_ba_spl_destroy(mvus_spl_handle h);
const char* mvus_ba_spl_last_error(mvus_spl_handle h);
/* One solve on the knots t[n] (n >= 2k+2, k+1-fold end knots):  min sum |x_i - s(u_i)|^2 + pscale * c^T P c,
 * P given by its upper band pen[(k+2)][n-k-1] (pen[d][j] = P[j-d][j]; NULL = plain least squares).
 * Outputs, any may be NULL: c[idim][n-k-1] B-spline coefficients, fp = residual sum of squares,
 * fpint[n-2k-1] = FITPACK's per-knot-interval residual sums (a point on an interior knot counts half on
 * each side), diag_sum = sum of the diagonal of the triangular factor of the normal matrix. */
int mvus_ba_spl_solve(mvus_spl_handle h, int32_t n, const double* t, const double* pen, double pscale,
                      double* c, double* fp, double* fpint, double* diag_sum);

/* Multi-GPU (one process per GPU): join an NCCL communicator whose unique id was
 * produced by mvus_ba_nccl_unique_id on rank 0 and broadcast by the host framework
 * (torch.distributed).  After this, detections set on each rank are that rank's shard
 * and the normal equations are summed over ranks. */
int mvus_ba_nccl_unique_id(char id_out[128]);
int mvus_ba_comm_init(mvus_ba_handle h, int32_t world_size, int32_t rank, const char id[128]);

/* Control-point bounds of the block ranges the ranks of a world_size-GPU solve own
 * (bounds[world_size + 1], bounds[0] = 0, bounds[world_size] = number of control points).  A
 * detection whose knot span (index of its last active control point at the start parameters)
 * lies in [bounds[r], bounds[r+1]) belongs on rank r: the exchange of the normal equations then
 * only moves the 3-control-point halos at the range boundaries.  Any other
 * partition of the detections stays correct, it only moves more.  Needs set_splines only. */
int mvus_ba_shard_bounds(mvus_ba_handle h, int32_t world_size, int64_t* bounds);

/* Timing helper for benchmarks: run `reps` back-to-back residual+Jacobian evaluations
 * (K1 + K1m) at x on the handle's stream, return the mean device ms (CUDA events). */
int mvus_ba_time_resjac(mvus_ba_handle h, const double* x, int32_t reps, double* ms_mean);
int mvus_ba_time_accumulate(mvus_ba_handle h, int32_t reps, double* ms_mean);

#ifdef __cplusplus
}
#endif
#endif /* MVUS_BA_H */
