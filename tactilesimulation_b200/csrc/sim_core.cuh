// Per-lane simulation math of the B200 tactile simulator.
//
// Everything here is templated on the scalar T (double, or Dual = value + one tangent) and
// is free of CUDA intrinsics, so the same source is (a) compiled into the sm_100a kernels in
// kernels.cu and (b) compiled by g++ into the test-only host harness (tests/emu) that lets
// the parity tests run without a GPU.  Cross-lane work (LU solve, reductions, marker
// striding) goes through a Tile policy: DevTile<LPE> = warp shuffles over a tile of LPE
// lanes, HostTile = one lane that owns everything.
//
// Formulation (differs from the reference on purpose; reference = DiffRedMax, cited as
// DH/... = externals/DiffHand/core/projects/redmax/...):
//   * residual of the implicit BDF1 step, DH/Simulation.cpp:1227-1251,
//       g(q1) = M(q1) (q1-q0-h qd0) - h^2 f(q1,(q1-q0)/h)
//     is evaluated matrix-free in WORLD-frame spatial algebra over the MOVING joints only
//     (fixed joints are folded into constant transforms by scene_lower.h):
//       outward sweep   E_0j, V_j = J qd (twist), X_j = J (q1-q0-h qd0) + h^2 Jdot qd
//       per body        a_i = I chi_i - h^2 (coriolis + gravity) in the body frame, pushed to the
//                       world-frame wrench accumulator of its joint; contact wrenches likewise
//       inward sweep    g_k = S_k . (sum of wrenches in the subtree of joint k) - h^2 (joint forces)
//     -- the reference instead builds J (42x7), Jdot, Mm, Km, Dm, dJ_dq ...
//     (DH/Simulation.cpp:256-454, DH/Robot.cpp:803-885).
//   * derivatives: H = dg/dq1 and the adjoint block G0 = dg/dq0 come from the same code on Dual
//     numbers, one reduced coordinate per lane (see dual.cuh); G1 = dg/dqdot0 = -h M is built
//     from mass-matrix columns (value arithmetic only).  With FORCE actuators G0 = -M + hD and
//     G1 = -hM in the reference's notation (DH/Simulation.cpp:1652,1657,1670,1692).
//   * Newton with backtracking line search (DH/Simulation.cpp:1150-1225): every line-search trial is
//     evaluated WITH its Jacobian columns, so an accepted trial is at once the next iterate's
//     (g, H) and -- when converged -- the tape's H; no residual is evaluated twice.
//   * joints: revolute, prismatic, planar, translational, free2d, free3d / spherical in the XYZ-Euler and in the
//     exponential chart; primitives: cuboid, cylinder, sphere, capsule; integrators: BDF1 (with the adjoint), BDF2 and
//     SDIRK2 as stages of one residual (stage_inputs / stage_coef).  Features are gated per kernel variant (kernel_layout.h).
//   * tactile readout: per-marker penalty force in the contacted box's frame
//     (DH/Sensor/TactileSensor.cpp:29-87); its adjoint is hand-written reverse mode
//     accumulating cotangents on (R2^T R1, R2^T(p1-p2), phi1, phi2) per candidate body
//     (replaces the 3M x 12 blocks of TactileSensor.cpp:89-235 and the 3M x n products of
//     Simulation.cpp:811-838).
#pragma once
#include "dual.cuh"
#include "kernel_layout.h"
#include "scene_layout.h"

#define TS_EPS 1e-8        // constants::eps of the reference (DH/Common.h)
#define TS_CULL_MARGIN 1e-9  // slack (metres) of the bounding-sphere culls; rounding is ~1e-17

#if defined(__CUDACC__)
#define TS_NOINLINE __noinline__
#else
#define TS_NOINLINE __attribute__((noinline))
#endif

// optional clock64 instrumentation of the forward kernel (development builds only: -DTS_PROFILE)
#if defined(TS_PROFILE) && defined(__CUDA_ARCH__)
#define TS_TIC(tl) const long long ts_t0_ = clock64()
#define TS_TOC(tl, slot) (tl).acc[slot] += clock64() - ts_t0_
#define TS_TIC2(tl) const long long ts_t1_ = clock64()
#define TS_TOC2(tl, slot) (tl).acc[slot] += clock64() - ts_t1_
#ifdef TS_PROFILE_GP
#define TS_GPT(tl, slot) do { const long long ts_now_ = clock64(); (tl).acc[slot] += ts_now_ - ts_gp_; ts_gp_ = ts_now_; } while (0)
#define TS_GPT0() long long ts_gp_ = clock64()
#endif
#else
#define TS_TIC(tl)
#define TS_TOC(tl, slot)
#define TS_TIC2(tl)
#define TS_TOC2(tl, slot)
#endif
#ifndef TS_GPT
#define TS_GPT(tl, slot)
#define TS_GPT0()
#endif
// cooperative contact-point phase (slots 8..15): items, phases with items, batches | cycles: publish, compute, wait, accumulate
#if defined(TS_PROFILE) && defined(__CUDA_ARCH__)
#define TS_CPT0() long long ts_cp_ = clock64()
#define TS_CPT(tl, slot) do { const long long ts_now_ = clock64(); (tl).acc[slot] += ts_now_ - ts_cp_; ts_cp_ = ts_now_; } while (0)
#define TS_CPN(tl, slot, n) (tl).acc[slot] += (n)
#else
#define TS_CPT0()
#define TS_CPT(tl, slot)
#define TS_CPN(tl, slot, n)
#endif

// Output streams (tactile field, tape, trajectory) are written once and never re-read by the forward kernel:
// evict-first stores keep them from displacing the per-lane tangent work space (local memory) in L2, whose
// write-back was the bulk of the DRAM traffic above the algorithmic bytes (profiles/r01_traffic.json).
HD void st_stream(double* p, double v) {
#ifdef __CUDA_ARCH__
  __stcs(p, v);
#else
  *p = v;
#endif
}
HD void st_stream(int* p, int v) {
#ifdef __CUDA_ARCH__
  __stcs(p, v);
#else
  *p = v;
#endif
}

#define TS_MAXJ KT_MAXJ
// Block-wide line-search service (ls_service_round below): 1 = the batched step lengths of a struggling line search are
// evaluated by lanes spread over all warps of the block instead of by the sixteen neighbouring lanes of the searching
// tile (16-dof variants; TactilePush hardly batches any more, and its step loop has its own block-wide phase).
#ifndef TS_LS_SERVICE
#define TS_LS_SERVICE (KT_MAXN > 8)
#endif
#define TS_LS_CAP 24          // step lengths per request (the reference's default line-search cap is 20)
#define TS_MAXN KT_MAXN
#define TS_MAXU KT_MAXU
#define TS_MAXCAND KT_MAXCAND

struct SceneView {
  const int* ib;
  const double* db;
  const double* mk;      // marker table (KM_STRIDE doubles per marker): global memory on the GPU
  int cmw;               // words of the contact bitmask output per env-step
  int ntape;             // doubles of the adjoint tape per env-step: H, G0, G1 (n x n each) + d f_r / d u per control
  int nj, n, nu, nee, nmark, nground, ngp, nact, nsens, max_iter, max_ls, nbody;
  int o_joint, o_body, o_ground, o_gp, o_act, o_ee, o_sensor;
  int d_joint, d_body, d_ground, d_gp, d_act, d_ee, d_sensor, d_points, d_markers;
  double h, tol, grav[3], gn[3], gx[3];
};

HDN inline void scene_view_init(SceneView& S, const int* ib, const double* db) {
  S.ib = ib; S.db = db;
  S.nj = ib[KI_NMJ]; S.n = ib[KI_N]; S.nu = ib[KI_NU]; S.nee = ib[KI_NEE];
  S.nmark = ib[KI_NMARK]; S.nground = ib[KI_NGROUND]; S.ngp = ib[KI_NGP];
  S.nact = ib[KI_NACT]; S.nsens = ib[KI_NSENS]; S.nbody = ib[KI_NBODY];
  S.max_iter = ib[KI_MAX_ITER]; S.max_ls = ib[KI_MAX_LS];
  S.cmw = ib[KI_CMW];
  S.ntape = 3 * S.n * S.n + S.nu;
  S.mk = db + ib[KI_D_MARKERS];
  S.o_joint = ib[KI_O_JOINT]; S.o_body = ib[KI_O_BODY]; S.o_ground = ib[KI_O_GROUND]; S.o_gp = ib[KI_O_GP];
  S.o_act = ib[KI_O_ACT]; S.o_ee = ib[KI_O_EE]; S.o_sensor = ib[KI_O_SENSOR];
  S.d_joint = ib[KI_D_JOINT]; S.d_body = ib[KI_D_BODY]; S.d_ground = ib[KI_D_GROUND]; S.d_gp = ib[KI_D_GP];
  S.d_act = ib[KI_D_ACT]; S.d_ee = ib[KI_D_EE]; S.d_sensor = ib[KI_D_SENSOR];
  S.d_points = ib[KI_D_POINTS]; S.d_markers = ib[KI_D_MARKERS];
  S.h = db[KD_H]; S.tol = db[KD_TOL];
  for (int i = 0; i < 3; ++i) { S.grav[i] = db[KD_GRAV + i]; S.gn[i] = db[KD_GN + i]; S.gx[i] = db[KD_GX + i]; }
}

// ------------------------------------------------------------------ small vector algebra
template <class A, class B, class C> HD void cross3(const A* a, const B* b, C* o) {
  C x = a[1] * b[2] - a[2] * b[1];
  C y = a[2] * b[0] - a[0] * b[2];
  C z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
template <class A, class B> HD auto dot3(const A* a, const B* b) -> decltype(a[0] * b[0]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
// o = R v   (R row-major 3x3)
template <class A, class B, class C> HD void mv3(const A* R, const B* v, C* o) {
  C x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  C y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  C z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
// o = R^T v
template <class A, class B, class C> HD void mtv3(const A* R, const B* v, C* o) {
  C x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  C y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  C z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
// O = A B
template <class A, class B, class C> HD void mm3(const A* a, const B* b, C* o) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
// world twist (w, v_O) -> twist of the frame (R,p) in its own coordinates:
//   w' = R^T w,  v' = R^T (v_O + w x p)
template <class A, class B, class C> HD void twist_to_frame(const A* R, const A* p, const B* in, C* out) {
  C t[3];
  cross3(in, p, t);
  t[0] = t[0] + in[3]; t[1] = t[1] + in[4]; t[2] = t[2] + in[5];
  mtv3(R, in, out);
  mtv3(R, t, out + 3);
}

// momentum-like product of the composite spatial inertia of joint j with a joint-frame twist:
//   h_ang = Ibar w + mc x v,  h_lin = m v + w x mc
template <class T>
HD void inertia_mul(const double* jd, const T* tw, T* hm) {
  const double* Ib = jd + KJ_IBAR;
  const double* mc = jd + KJ_MC;
  const double m = jd[KJ_MASS];
  T a[3], b[3];
  mv3(Ib, tw, hm);
  cross3(mc, tw + 3, a);
  cross3(tw, mc, b);
  for (int i = 0; i < 3; ++i) { hm[i] = hm[i] + a[i]; hm[3 + i] = m * tw[3 + i] + b[i]; }
}

// ------------------------------------------------------------------ per-lane work space
// One record per MOVING joint, world frame: pose, twist V = J qd, X = J dl + h^2 Jdot qd and the
// wrench accumulator of the joint's own bodies (summed over the subtree by the inward sweep).
// Two storage policies behind the same accessors:
//   Work<T>    plain per-lane arrays (host harness, value-only readout kernel);
//   WorkSplit  Dual numbers on the GPU: the VALUES are identical in every lane of a tile (same code on the
//              same inputs), so they are kept ONCE per tile in shared memory; only the tangents are per-lane
//              (local memory).  This halves the local-memory working set, which is what spilled out of L2.
enum { WK_R0 = 0, WK_P0 = 9, WK_V = 12, WK_X = 18, WK_WA = 24, WK_REC = 30 };

// world frames of the sensor body (slot 0) and its candidate bodies (slots 1..ncand), values only
struct Frames {
  double R[1 + TS_MAXCAND][9], p[1 + TS_MAXCAND][3], ph[1 + TS_MAXCAND][6];
  double R21[TS_MAXCAND][9], r21[TS_MAXCAND][3];   // pad frame in each candidate's frame (fast classification)
  bool near[1 + TS_MAXCAND];   // candidate may touch a marker (bounding-sphere test)
};

// Step state that is identical in every lane of a tile (the lanes run the same value arithmetic):
// kept once per tile (shared memory on the GPU).  Read-modify-write updates go through registers
// with a tile_sync between the reads and the writes.
struct TileState {
  double q[TS_MAXN], qd[TS_MAXN], u[TS_MAXU];       // state at the start of the step, controls of the step
  double x[TS_MAXN], xn[TS_MAXN], dx[TS_MAXN];     // Newton iterate, line-search trial, Newton direction
  double xq[TS_MAXN], xv[TS_MAXN], xl[TS_MAXN];    // inputs (q1, qd1, dl) of the evaluation in flight
  double g[TS_MAXN];                               // residual of the last evaluation
  double hs[TS_MAXN * TS_MAXN];                    // transposition scratch of the cooperative LU
#if KT_MULTISTEP
  double pq[TS_MAXN], pqd[TS_MAXN];                // second state of the two-state formulas: BDF2 (q, qd) one step back,
                                                   // SDIRK2 second stage (q_alpha, qd_alpha)
#endif
#if TS_LS_SERVICE
  // request of this tile to the block's line-search service (ls_service_round): ||g(x + 2^-i alpha0 dx)||, i < ls_n
  double lsn[TS_LS_CAP];                           // answers (HUGE_VAL = not evaluated)
  double ls_alpha0, ls_gnorm;                      // first step length, the norm a trial has to beat
  const double* ls_db;                             // double table of the requester's scene view (per-environment parameters)
  int ls_req, ls_n, ls_mode, ls_found;             // request pending, step lengths asked for, implicit stage, one accepted
#endif
};

// ---- implicit stages (DH/Simulation.cpp:1076-1092, 1227-1235, 1353-1364, 1425-1450, 1466-1475, 1536-1548).
// Every stage solves g(x) = M(x) dl(x) - beta f(x, qd(x)) = 0 with affine qd(x), dl(x):
//   TS_ST_BDF1    qd = (x - q0) / h                        dl = x - q0 - h qd0                               beta = h^2
//   TS_ST_SDIRK_A the same with h := alpha h, alpha = (2 - sqrt 2) / 2  (first stage of SDIRK2 -> q_alpha, qd_alpha)
//   TS_ST_SDIRK_B qd = (x + (1/alpha - 2) q0 - (1 - alpha)/alpha q_alpha) / (alpha h)
//                 dl = x - q0 - (2 alpha - 1) h qd0 - 2 (1 - alpha) h qd_alpha                            beta = alpha^2 h^2
//   TS_ST_BDF2    qd = 3/(2h) (x - 4/3 q1 + 1/3 q0)        dl = x - 4/3 q1 + 1/3 q0 - 8/9 h qd1 + 2/9 h qd0   beta = 4/9 h^2
// (q, qd) of the tile state is the state at the start of the step (q1 of BDF2), (pq, pqd) the second state.
enum { TS_ST_BDF1 = 0, TS_ST_SDIRK_A = 1, TS_ST_SDIRK_B = 2, TS_ST_BDF2 = 3 };
#define TS_SDIRK_ALPHA ((2. - sqrt(2.)) / 2.)

// d qd / d x of the stage and its force scale
HD void stage_coef(const SceneView& S, int mode, double& tv, double& beta) {
  const double h = S.h;
  if (KT_MULTISTEP && mode == TS_ST_SDIRK_A) { const double ha = TS_SDIRK_ALPHA * h; tv = 1.0 / ha; beta = ha * ha; }
  else if (KT_MULTISTEP && mode == TS_ST_SDIRK_B) { const double al = TS_SDIRK_ALPHA; tv = 1.0 / (al * h); beta = al * al * h * h; }
  else if (KT_MULTISTEP && mode == TS_ST_BDF2) { tv = 3. / (2. * h); beta = 4. / 9. * h * h; }
  else { tv = (1.0 - 0.0) / h; beta = h * h; }
}

// (qd, dl) of coordinate i at the stage point xi
HD void stage_inputs(const SceneView& S, const TileState& ts, int mode, int i, double xi, double& xv, double& xl) {
  const double h = S.h, qi = ts.q[i], vi = ts.qd[i];
#if KT_MULTISTEP
  if (mode == TS_ST_SDIRK_A) {
    const double ha = TS_SDIRK_ALPHA * h;
    xv = (xi - qi) / ha;
    xl = xi - qi - ha * vi;
    return;
  }
  if (mode == TS_ST_SDIRK_B) {
    const double al = TS_SDIRK_ALPHA;
    xv = (xi + (1. / al - 2.) * qi - (1. - al) / al * ts.pq[i]) / (al * h);
    xl = xi - qi - (2. * al - 1.) * h * vi - 2. * (1. - al) * h * ts.pqd[i];
    return;
  }
  if (mode == TS_ST_BDF2) {
    xv = 3. / (2. * h) * (xi - 4. / 3. * qi + 1. / 3. * ts.pq[i]);
    xl = xi - 4. / 3. * qi + 1. / 3. * ts.pq[i] - 8. / 9. * h * vi + 2. / 9. * h * ts.pqd[i];
    return;
  }
#endif
  xv = (xi - qi) / h;
  xl = xi - qi - h * vi;
}

// inputs (q, qd, dl) of one evaluation: plain arrays ...
template <class T> struct ArrIn {
  const T *q_, *qd_, *dl_;
  const double *q0_, *qd0_;                 // state at the start of the step (position-controlled motors)
  HD T q(int i) const { return q_[i]; }
  HD T qd(int i) const { return qd_[i]; }
  HD T dl(int i) const { return dl_[i]; }
  HD T q0(int i) const { return T(q0_[i]); }
  HD T qd0(int i) const { return T(qd0_[i]); }
};
// ... or Dual numbers whose values live in the tile state and whose tangent is a unit seed on dof k
struct SeedIn {
  const double *xq, *xv, *xl;
  const double *q0v, *qd0v;                 // state at the start of the step (position-controlled motors)
  int k;
  double tq, tv, tl, tq0, tqd0;
  HD Dual q(int i) const { return mkdual(xq[i], i == k ? tq : 0.0); }
  HD Dual qd(int i) const { return mkdual(xv[i], i == k ? tv : 0.0); }
  HD Dual dl(int i) const { return mkdual(xl[i], i == k ? tl : 0.0); }
  HD Dual q0(int i) const { return mkdual(q0v[i], i == k ? tq0 : 0.0); }
  HD Dual qd0(int i) const { return mkdual(qd0v[i], i == k ? tqd0 : 0.0); }
};

template <class T> struct WorkRec {          // the joint records alone (value-only line-search trials)
  typedef T Scalar;
  double beta;                               // force scale of the residual in flight (h^2 for BDF1), set by eval_g
  int gp_any;                                // the last evaluation had an active general-primitive contact point
  T rec[TS_MAXJ][WK_REC];
  HD T get(int j, int o) const { return rec[j][o]; }
  HD double getv(int j, int o) const { return val(rec[j][o]); }
  HD void put(int j, int o, const T& x) { rec[j][o] = x; }
};
template <class T> struct Work : WorkRec<T> {
  Frames fr;
  TileState ts;
  HD TileState& state() { return ts; }
  HD double* scratch() { return ts.hs; }
  HD Frames& frames() { return fr; }
};

struct WorkSplit {
  typedef Dual Scalar;
  double* sv;                      // [nj][WK_REC] values, shared by the lanes of the tile
  Frames* fr;                      // per tile, shared memory
  TileState* ts;                   // per tile, shared memory
  HD TileState& state() { return *ts; }
  HD double* scratch() { return ts->hs; }
  double beta;                     // force scale of the residual in flight (h^2 for BDF1), set by eval_g
  int gp_any;                      // the last evaluation had an active general-primitive contact point
  double dt[TS_MAXJ][WK_REC];      // tangents of this lane
  HD Dual get(int j, int o) const { return mkdual(sv[j * WK_REC + o], dt[j][o]); }
  HD double getv(int j, int o) const { return sv[j * WK_REC + o]; }
  HD void put(int j, int o, const Dual& x) { sv[j * WK_REC + o] = x.v; dt[j][o] = x.d; }
  HD Frames& frames() { return *fr; }
};
// ... or values only, kept once per tile in the same shared region (the tactile read-out pass: kinematics of a recorded
// state without derivatives -- a third of the arithmetic of the dual-number work space and no local-memory tangents)
struct WorkSharedV {
  typedef double Scalar;
  double* sv;
  Frames* fr;
  TileState* ts;
  double beta;
  int gp_any;
  HD TileState& state() { return *ts; }
  HD double get(int j, int o) const { return sv[j * WK_REC + o]; }
  HD double getv(int j, int o) const { return sv[j * WK_REC + o]; }
  HD void put(int j, int o, double x) { sv[j * WK_REC + o] = x; }
  HD Frames& frames() { return *fr; }
};
template <class WK, class T> HD void wk_ld(const WK& W, int j, int o, int n, T* out) {
  for (int i = 0; i < n; ++i) out[i] = W.get(j, o + i);
}
template <class WK> HD void wk_ldv(const WK& W, int j, int o, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = W.getv(j, o + i);
}
template <class WK, class T> HD void wk_st(WK& W, int j, int o, int n, const T* in) {
  for (int i = 0; i < n; ++i) W.put(j, o + i, in[i]);
}

// ------------------------------------------------------------------ exponential coordinates (free3d-exp joints)
// Coefficient functions of t = |r|^2:  A = sin(th)/th,  B = (1 - cos th)/th^2,  C = (th - sin th)/th^3  and dB/dt,
// dC/dt.  R(r) = 1 + A [r] + B [r]^2 is math::exp of the reference (DH/Utils.h:82-98); the left Jacobian
// JL(r) = 1 + B [r] + C [r]^2 has the columns unskew(dR/dr_j R^T) that DH/Joint/JointSphericalExp.cpp:98-108
// writes as (r_j [r] + [[r](1 - R) e_j]) / |r|^2.  Power series below t = 1e-4: no 0/0 at r = 0 (where the ball of
// the rolling-ball scene starts) and differentiable there by the same dual-number code.
template <class T>
HD void so3_coefs(T t, T& A, T& B, T& C, T& dB, T& dC) {
  if (val(t) < 1e-4) {
    A = 1.0 + t * (-1.0 / 6.0 + t * (1.0 / 120.0 + t * (-1.0 / 5040.0 + t * (1.0 / 362880.0))));
    B = 0.5 + t * (-1.0 / 24.0 + t * (1.0 / 720.0 + t * (-1.0 / 40320.0)));
    C = 1.0 / 6.0 + t * (-1.0 / 120.0 + t * (1.0 / 5040.0 + t * (-1.0 / 362880.0)));
    dB = -1.0 / 24.0 + t * (1.0 / 360.0 + t * (-1.0 / 13440.0));
    dC = -1.0 / 120.0 + t * (1.0 / 2520.0 + t * (-1.0 / 120960.0));
  } else {
    T th = dsqrt(t), sn, cs;
    dsincos(th, sn, cs);
    A = sn / th;
    B = (1.0 - cs) / t;
    C = (th - sn) / (t * th);
    dB = (A - 2.0 * B) / (2.0 * t);
    dC = (B - 3.0 * C) / (2.0 * t);
  }
}
// R = 1 + A [r] + B (r r^T - t 1), row-major
template <class T>
HD void so3_exp(const T* r, T A, T B, T t, T* R) {
  R[0] = 1.0 + B * (r[0] * r[0] - t); R[1] = B * (r[0] * r[1]) - A * r[2]; R[2] = B * (r[0] * r[2]) + A * r[1];
  R[3] = B * (r[1] * r[0]) + A * r[2]; R[4] = 1.0 + B * (r[1] * r[1] - t); R[5] = B * (r[1] * r[2]) - A * r[0];
  R[6] = B * (r[2] * r[0]) - A * r[1]; R[7] = B * (r[2] * r[1]) + A * r[0]; R[8] = 1.0 + B * (r[2] * r[2] - t);
}
// o = (1 + sB [r] + C [r]^2) v :  sB = +B left Jacobian (angular velocity in the pre-motion frame per unit rdot),
// sB = -B right Jacobian (the same in the child frame: the S_j of DH/Joint/JointSphericalExp.cpp:172-175)
template <class T, class V>
HD void so3_jac_mul(const T* r, T sB, T C, const V* v, T* o) {
  T c1[3], c2[3];
  cross3(r, v, c1);
  cross3(r, c1, c2);
  for (int i = 0; i < 3; ++i) o[i] = v[i] + sB * c1[i] + C * c2[i];
}
// World axes of the six coordinates of a free3d-exp joint whose child frame is (R0, .): translations along
// Ra e_i = R0 (row i of R(r)), rotations about R0 JR(r) e_i.
template <class T>
HD void exp_world_axes(const T* R0, T r1, T r2, T r3, T (*tax)[3], T (*rax)[3]) {
  T r[3] = {r1, r2, r3};
  T t = r1 * r1 + r2 * r2 + r3 * r3, A, B, C, dB, dC;
  so3_coefs(t, A, B, C, dB, dC);
  T Rq[9];
  so3_exp(r, A, B, t, Rq);
  for (int i = 0; i < 3; ++i) {
    T row[3] = {Rq[3 * i], Rq[3 * i + 1], Rq[3 * i + 2]};
    mv3(R0, row, tax[i]);
    double ei[3] = {i == 0 ? 1.0 : 0.0, i == 1 ? 1.0 : 0.0, i == 2 ? 1.0 : 0.0};
    T jc[3];
    so3_jac_mul(r, -B, C, ei, jc);
    mv3(R0, jc, rax[i]);
  }
}

// Outward sweep over the moving joints.  dyn=false computes poses and twists only.
// Joint models: DH/Joint/JointRevolute.cpp:38-69, JointPrismatic.cpp:22-45, JointPlanar.cpp:7-33,
// JointTranslational.cpp:9-40; recursion DH/Joint/Joint.cpp:119-165.
template <class WK, class In>
HDN void kinematics(const SceneView& S, const In& in, WK& W, bool dyn) {
  typedef typename WK::Scalar T;
  const double h2 = W.beta;             // only read when dyn
  for (int j = 0; j < S.nj; ++j) {
    const int* ji = S.ib + S.o_joint + j * KJ_ISTRIDE;
    const double* jd = S.db + S.d_joint + j * KJ_DSTRIDE;
    const int jt = ji[0], par = ji[1], qo = ji[2];
    const double* Rc = jd + KJ_RA;
    const double* pc = jd + KJ_PA;
    const double* a0 = jd + KJ_AX0;
    const double* a1 = jd + KJ_AX1;
    // pre-motion frame = parent moving frame * constant offset
    T Ra[9], pa[3], Vp[6], Xp[6];
    if (par < 0) {
      for (int i = 0; i < 9; ++i) Ra[i] = Rc[i];
      for (int i = 0; i < 3; ++i) pa[i] = pc[i];
      for (int i = 0; i < 6; ++i) { Vp[i] = 0.0; Xp[i] = 0.0; }
    } else {
      T Rp[9], pp[3];
      wk_ld(W, par, WK_R0, 9, Rp);
      wk_ld(W, par, WK_P0, 3, pp);
      mm3(Rp, Rc, Ra);
      mv3(Rp, pc, pa);
      for (int i = 0; i < 3; ++i) pa[i] = pa[i] + pp[i];
      wk_ld(W, par, WK_V, 6, Vp);
      if (dyn) wk_ld(W, par, WK_X, 6, Xp);
      else for (int i = 0; i < 6; ++i) Xp[i] = 0.0;
    }
    // world screw sums  sq = sum_d S_d qd_d,  sl = sum_d S_d dl_d
    T sq[6], sl[6];
    for (int i = 0; i < 6; ++i) { sq[i] = 0.0; sl[i] = 0.0; }
    T R0[9], p0[3];
    if (jt == TS_JT_REVOLUTE) {          // Q = exp([axis] q)
      T s, c;
      dsincos(in.q(qo), s, c);
      T c1 = 1.0 - c;
      T Rq[9];
      Rq[0] = c + c1 * (a0[0] * a0[0]); Rq[1] = c1 * (a0[0] * a0[1]) - s * a0[2]; Rq[2] = c1 * (a0[0] * a0[2]) + s * a0[1];
      Rq[3] = c1 * (a0[1] * a0[0]) + s * a0[2]; Rq[4] = c + c1 * (a0[1] * a0[1]); Rq[5] = c1 * (a0[1] * a0[2]) - s * a0[0];
      Rq[6] = c1 * (a0[2] * a0[0]) - s * a0[1]; Rq[7] = c1 * (a0[2] * a0[1]) + s * a0[0]; Rq[8] = c + c1 * (a0[2] * a0[2]);
      mm3(Ra, Rq, R0);
      for (int i = 0; i < 3; ++i) p0[i] = pa[i];
      T w[3], m[3];
      mv3(Ra, a0, w);                    // world axis; screw = (w, p x w)
      cross3(pa, w, m);
      for (int i = 0; i < 3; ++i) {
        sq[i] = w[i] * in.qd(qo); sq[3 + i] = m[i] * in.qd(qo);
        if (dyn) { sl[i] = w[i] * in.dl(qo); sl[3 + i] = m[i] * in.dl(qo); }
      }
    } else if (KT_FREE3D && jt == TS_JT_FREE2D) {
      // q = (x, y, theta): Q = [Rz(theta) (x, y, 0); 0 1]   (DH/Joint/JointFree2D.cpp).  The free3d-euler chart below
      // restricted to (p_x, p_y, r_3): twist xi = (theta_dot e_z, pdot + p x theta_dot e_z) about a's origin; the axis
      // does not move, so the velocity-product term is xi_dot = (0, pdot x theta_dot e_z).
      T s3, c3;
      dsincos(in.q(qo + 2), s3, c3);
      T Rq[9];
      Rq[0] = c3; Rq[1] = -s3; Rq[2] = 0.0;
      Rq[3] = s3; Rq[4] = c3; Rq[5] = 0.0;
      Rq[6] = 0.0; Rq[7] = 0.0; Rq[8] = 1.0;
      mm3(Ra, Rq, R0);
      T pq[3], t[3];
      pq[0] = in.q(qo); pq[1] = in.q(qo + 1); pq[2] = 0.0;
      mv3(Ra, pq, t);
      for (int i = 0; i < 3; ++i) p0[i] = pa[i] + t[i];
      for (int pass = 0; pass < (dyn ? 2 : 1); ++pass) {
        T pd[3], wa[3], va[3], c[3];
        pd[0] = pass ? in.dl(qo) : in.qd(qo); pd[1] = pass ? in.dl(qo + 1) : in.qd(qo + 1); pd[2] = 0.0;
        wa[0] = 0.0; wa[1] = 0.0; wa[2] = pass ? in.dl(qo + 2) : in.qd(qo + 2);
        cross3(pq, wa, c);
        for (int i = 0; i < 3; ++i) va[i] = pd[i] + c[i];
        T* so = pass ? sl : sq;
        T w[3], v[3], pw[3];
        mv3(Ra, wa, w);
        mv3(Ra, va, v);
        cross3(pa, w, pw);
        for (int i = 0; i < 3; ++i) { so[i] = w[i]; so[3 + i] = v[i] + pw[i]; }
      }
      if (dyn) {
        T pd[3], wa[3], vd[3], v[3];
        pd[0] = in.qd(qo); pd[1] = in.qd(qo + 1); pd[2] = 0.0;
        wa[0] = 0.0; wa[1] = 0.0; wa[2] = in.qd(qo + 2);
        cross3(pd, wa, vd);
        mv3(Ra, vd, v);
        for (int i = 0; i < 3; ++i) sl[3 + i] = sl[3 + i] + h2 * v[i];
      }
    } else if (KT_FREE3D && (jt == TS_JT_FREE3D_EULER || jt == TS_JT_SPHERICAL_EULER)) {
      // (spherical-euler, DH/Joint/JointSphericalEuler.cpp: the rotation alone, q = r, p = 0)
      const bool trans = jt == TS_JT_FREE3D_EULER;
      const int ro = trans ? qo + 3 : qo;
      // q = (p, r): Q = [R(r) p; 0 1], R = Rx(r1) Ry(r2) Rz(r3)   (DH/Joint/JointFree3DEuler.cpp:14-104,
      // JointSphericalEuler.cpp:15-60).  In the pre-motion frame a the child moves with the twist
      // xi = (G rdot, pdot + p x G rdot) about a's origin, G = [e_x, Rx e_y, Rx Ry e_z]; the rotation axes move
      // with r, which adds xi_dot = (Gdot rdot, pdot x G rdot + p x Gdot rdot) to the velocity-product term.
      T s1, c1, s2, c2, s3, c3;
      dsincos(in.q(ro), s1, c1);
      dsincos(in.q(ro + 1), s2, c2);
      dsincos(in.q(ro + 2), s3, c3);
      T Rq[9];
      Rq[0] = c2 * c3; Rq[1] = -(c2 * s3); Rq[2] = s2;
      Rq[3] = c1 * s3 + c3 * (s1 * s2); Rq[4] = c1 * c3 - (s1 * s2) * s3; Rq[5] = -(c2 * s1);
      Rq[6] = s1 * s3 - (c1 * c3) * s2; Rq[7] = c3 * s1 + (c1 * s2) * s3; Rq[8] = c1 * c2;
      mm3(Ra, Rq, R0);
      T pq[3], t[3];
      for (int i = 0; i < 3; ++i) pq[i] = trans ? in.q(qo + i) : T(0.0);
      mv3(Ra, pq, t);
      for (int i = 0; i < 3; ++i) p0[i] = pa[i] + t[i];
      T g2[3], g3[3];
      g2[0] = 0.0; g2[1] = c1; g2[2] = s1;
      g3[0] = s2; g3[1] = -(s1 * c2); g3[2] = c1 * c2;
      for (int pass = 0; pass < (dyn ? 2 : 1); ++pass) {
        T rd[3], pd[3];
        for (int i = 0; i < 3; ++i) {
          pd[i] = trans ? (pass ? in.dl(qo + i) : in.qd(qo + i)) : T(0.0);
          rd[i] = pass ? in.dl(ro + i) : in.qd(ro + i);
        }
        T wa[3], va[3], c[3];
        wa[0] = rd[0] + g3[0] * rd[2];
        wa[1] = g2[1] * rd[1] + g3[1] * rd[2];
        wa[2] = g2[2] * rd[1] + g3[2] * rd[2];
        cross3(pq, wa, c);
        for (int i = 0; i < 3; ++i) va[i] = pd[i] + c[i];
        T* so = pass ? sl : sq;
        T w[3], v[3], pw[3];
        mv3(Ra, wa, w);
        mv3(Ra, va, v);
        cross3(pa, w, pw);
        for (int i = 0; i < 3; ++i) { so[i] = w[i]; so[3 + i] = v[i] + pw[i]; }
      }
      if (dyn) {
        // h^2 Ad(E_0a) xi_dot, folded into sl (X_j = X_p + sl + h^2 ad(V_p) sq)
        T r1 = in.qd(ro), r2 = in.qd(ro + 1), r3 = in.qd(ro + 2);
        T wd[3], wa[3], vd[3], c0[3], c1v[3], pd[3];
        // Gdot rdot = r2 * d(g2)/dt + r3 * d(g3)/dt,  d(g2)/dt = r1 (0,-s1,c1),  d(g3)/dt = r1 (0,-c1 c2,-s1 c2) + r2 (c2, s1 s2, -c1 s2)
        wd[0] = r3 * (r2 * c2);
        wd[1] = r2 * (r1 * (-s1)) + r3 * (r1 * (-(c1 * c2)) + r2 * (s1 * s2));
        wd[2] = r2 * (r1 * c1) + r3 * (r1 * (-(s1 * c2)) + r2 * (-(c1 * s2)));
        wa[0] = r1 + g3[0] * r3;
        wa[1] = g2[1] * r2 + g3[1] * r3;
        wa[2] = g2[2] * r2 + g3[2] * r3;
        for (int i = 0; i < 3; ++i) pd[i] = trans ? in.qd(qo + i) : T(0.0);
        cross3(pd, wa, c0);
        cross3(pq, wd, c1v);
        for (int i = 0; i < 3; ++i) vd[i] = c0[i] + c1v[i];
        T w[3], v[3], pw[3];
        mv3(Ra, wd, w);
        mv3(Ra, vd, v);
        cross3(pa, w, pw);
        for (int i = 0; i < 3; ++i) { sl[i] = sl[i] + h2 * w[i]; sl[3 + i] = sl[3 + i] + h2 * (v[i] + pw[i]); }
      }
    } else if (KT_EXP3D && (jt == TS_JT_FREE3D_EXP || jt == TS_JT_SPHERICAL_EXP)) {
      const bool trans = jt == TS_JT_FREE3D_EXP;      // spherical-exp: the rotation alone, q = r, p = 0
      const int ro = trans ? qo + 3 : qo;
      // q = (p, r): Q = [exp([r]) p; 0 1]   (DH/Joint/JointFree3DExp.cpp:13-103, JointSphericalExp.cpp:22-245).
      // Same structure as the Euler chart above with G replaced by the left Jacobian JL(r) of SO(3):
      // xi = (JL rdot, pdot + p x JL rdot),  xi_dot = (JLdot rdot, pdot x JL rdot + p x JLdot rdot) with
      // JLdot rdot = Bdot (r x rdot) + Cdot r x (r x rdot) + C rdot x (r x rdot)   (B [rdot] rdot = 0).
      T rr[3], pq[3];
      for (int i = 0; i < 3; ++i) { pq[i] = trans ? in.q(qo + i) : T(0.0); rr[i] = in.q(ro + i); }
      T t2 = rr[0] * rr[0] + rr[1] * rr[1] + rr[2] * rr[2], cA, cB, cC, dB, dC;
      so3_coefs(t2, cA, cB, cC, dB, dC);
      T Rq[9], t[3];
      so3_exp(rr, cA, cB, t2, Rq);
      mm3(Ra, Rq, R0);
      mv3(Ra, pq, t);
      for (int i = 0; i < 3; ++i) p0[i] = pa[i] + t[i];
      for (int pass = 0; pass < (dyn ? 2 : 1); ++pass) {
        T rd[3], pd[3];
        for (int i = 0; i < 3; ++i) {
          pd[i] = trans ? (pass ? in.dl(qo + i) : in.qd(qo + i)) : T(0.0);
          rd[i] = pass ? in.dl(ro + i) : in.qd(ro + i);
        }
        T wa[3], va[3], c[3];
        so3_jac_mul(rr, cB, cC, rd, wa);
        cross3(pq, wa, c);
        for (int i = 0; i < 3; ++i) va[i] = pd[i] + c[i];
        T* so = pass ? sl : sq;
        T w[3], v[3], pw[3];
        mv3(Ra, wa, w);
        mv3(Ra, va, v);
        cross3(pa, w, pw);
        for (int i = 0; i < 3; ++i) { so[i] = w[i]; so[3 + i] = v[i] + pw[i]; }
      }
      if (dyn) {
        T rd[3], pd[3], wa[3], wd[3], c1[3], c2[3], c3[3], vd[3], c0[3], c1v[3];
        for (int i = 0; i < 3; ++i) { pd[i] = trans ? in.qd(qo + i) : T(0.0); rd[i] = in.qd(ro + i); }
        so3_jac_mul(rr, cB, cC, rd, wa);
        T rdr = 2.0 * (rr[0] * rd[0] + rr[1] * rd[1] + rr[2] * rd[2]);     // d|r|^2/dt
        cross3(rr, rd, c1);
        cross3(rr, c1, c2);
        cross3(rd, c1, c3);
        for (int i = 0; i < 3; ++i) wd[i] = (dB * rdr) * c1[i] + (dC * rdr) * c2[i] + cC * c3[i];
        cross3(pd, wa, c0);
        cross3(pq, wd, c1v);
        for (int i = 0; i < 3; ++i) vd[i] = c0[i] + c1v[i];
        T w[3], v[3], pw[3];
        mv3(Ra, wd, w);
        mv3(Ra, vd, v);
        cross3(pa, w, pw);
        for (int i = 0; i < 3; ++i) { sl[i] = sl[i] + h2 * w[i]; sl[3 + i] = sl[3 + i] + h2 * (v[i] + pw[i]); }
      }
    } else {
      for (int i = 0; i < 9; ++i) R0[i] = Ra[i];
      T pq[3], vq[3], vl[3];
      for (int i = 0; i < 3; ++i) { pq[i] = 0.0; vq[i] = 0.0; vl[i] = 0.0; }
      if (jt == TS_JT_PRISMATIC) {
        for (int i = 0; i < 3; ++i) { pq[i] = a0[i] * in.q(qo); vq[i] = a0[i] * in.qd(qo); if (dyn) vl[i] = a0[i] * in.dl(qo); }
      } else if (jt == TS_JT_PLANAR) {
        for (int i = 0; i < 3; ++i) {
          pq[i] = a0[i] * in.q(qo) + a1[i] * in.q(qo + 1);
          vq[i] = a0[i] * in.qd(qo) + a1[i] * in.qd(qo + 1);
          if (dyn) vl[i] = a0[i] * in.dl(qo) + a1[i] * in.dl(qo + 1);
        }
      } else if (jt == TS_JT_TRANSLATIONAL) {
        for (int i = 0; i < 3; ++i) { pq[i] = in.q(qo + i); vq[i] = in.qd(qo + i); if (dyn) vl[i] = in.dl(qo + i); }
      }
      T t[3];
      mv3(Ra, pq, t);
      for (int i = 0; i < 3; ++i) p0[i] = pa[i] + t[i];
      mv3(Ra, vq, sq + 3);               // pure translations: screw = (0, Ra axis)
      if (dyn) mv3(Ra, vl, sl + 3);
    }
    wk_st(W, j, WK_R0, 9, R0);
    wk_st(W, j, WK_P0, 3, p0);
    T Vl[6];
    for (int i = 0; i < 6; ++i) Vl[i] = Vp[i] + sq[i];
    wk_st(W, j, WK_V, 6, Vl);
    if (dyn) {
      // X_j = X_p + S dl + h^2 ad(V_p)(S qd): the joint axes are fixed in the parent body
      T c0[3], c1v[3], c2[3], Xl[6];
      cross3(Vp, sq, c0);                // w_p x s_w
      cross3(Vp, sq + 3, c1v);           // w_p x s_v
      cross3(Vp + 3, sq, c2);            // v_p x s_w
      for (int i = 0; i < 3; ++i) {
        Xl[i] = Xp[i] + sl[i] + h2 * c0[i];
        Xl[3 + i] = Xp[3 + i] + sl[3 + i] + h2 * (c1v[i] + c2[i]);
      }
      wk_st(W, j, WK_X, 6, Xl);
      // composite rigid body of this joint, while its frame is still in registers:
      // a_j = I chi_j - h^2 (coriolis + gravity) in the joint frame (per reference body:
      // DH/Body/Body.cpp:234-247; summed by linearity of the spatial inertia), pushed to the world frame
      if (ji[5]) {
        T ph[6], ch[6], hp[6], a[6];
        twist_to_frame(R0, p0, Vl, ph);
        twist_to_frame(R0, p0, Xl, ch);
        inertia_mul(jd, ph, hp);
        inertia_mul(jd, ch, a);
        // coriolis ad(phi)^T h = (h_ang x w + h_lin x v ; h_lin x w), gravity (mc x R^T g ; m R^T g)
        T d0[3], d1[3], d2[3], gb[3], gm[3];
        cross3(hp, ph, d0);
        cross3(hp + 3, ph + 3, d1);
        cross3(hp + 3, ph, d2);
        mtv3(R0, S.grav, gb);
        cross3(jd + KJ_MC, gb, gm);
        for (int i = 0; i < 3; ++i) {
          a[i] = a[i] - h2 * ((d0[i] + d1[i]) + gm[i]);
          a[3 + i] = a[3 + i] - h2 * (d2[i] + jd[KJ_MASS] * gb[i]);
        }
        T f[3], t[3], pf[3];
        mv3(R0, a + 3, f);
        mv3(R0, a, t);
        cross3(p0, f, pf);
        for (int i = 0; i < 3; ++i) { W.put(j, WK_WA + i, t[i] + pf[i]); W.put(j, WK_WA + 3 + i, f[i]); }
      } else {
        for (int i = 0; i < 6; ++i) W.put(j, WK_WA + i, T(0.0));
      }
    }
  }
}

// pose of body b (and its twist in body coordinates) from the work space
template <class WK, class T>
HD void body_frame(const SceneView& S, const WK& W, int b, T* R, T* p, T* ph) {
  const int j = S.ib[S.o_body + b * KB_ISTRIDE];
  const double* bd = S.db + S.d_body + b * KB_DSTRIDE;
  if (j < 0) {
    for (int i = 0; i < 9; ++i) R[i] = bd[KB_RMI + i];
    for (int i = 0; i < 3; ++i) p[i] = bd[KB_PMI + i];
    if (ph) for (int i = 0; i < 6; ++i) ph[i] = 0.0;
    return;
  }
  T R0[9], p0[3];
  wk_ld(W, j, WK_R0, 9, R0);
  wk_ld(W, j, WK_P0, 3, p0);
  mm3(R0, bd + KB_RMI, R);
  mv3(R0, bd + KB_PMI, p);
  for (int i = 0; i < 3; ++i) p[i] = p[i] + p0[i];
  if (ph) {
    T V[6];
    wk_ld(W, j, WK_V, 6, V);
    twist_to_frame(R, p, V, ph);
  }
}
// value-only variant (readouts run on the values of whatever scalar the work space holds)
template <class WK>
HD void body_frame_v(const SceneView& S, const WK& W, int b, double* R, double* p, double* ph) {
  const int j = S.ib[S.o_body + b * KB_ISTRIDE];
  const double* bd = S.db + S.d_body + b * KB_DSTRIDE;
  if (j < 0) {
    for (int i = 0; i < 9; ++i) R[i] = bd[KB_RMI + i];
    for (int i = 0; i < 3; ++i) p[i] = bd[KB_PMI + i];
    for (int i = 0; i < 6; ++i) ph[i] = 0.0;
    return;
  }
  double R0[9], p0[3], V[6];
  wk_ldv(W, j, WK_R0, 9, R0);
  wk_ldv(W, j, WK_P0, 3, p0);
  wk_ldv(W, j, WK_V, 6, V);
  mm3(R0, bd + KB_RMI, R);
  mv3(R0, bd + KB_PMI, p);
  for (int i = 0; i < 3; ++i) p[i] += p0[i];
  twist_to_frame(R, p, V, ph);
}

// body-frame wrench (moment; force) of the body at (R,p) -> world wrench about the origin, added
// (times scale) to the accumulator of moving joint j
template <class WK, class T>
HD void push_wrench(WK& W, int j, const T* R, const T* p, const T* wr, double scale) {
  if (j < 0) return;
  T f[3], t[3], pf[3];
  mv3(R, wr + 3, f);
  mv3(R, wr, t);
  cross3(p, f, pf);
  for (int i = 0; i < 3; ++i) {
    W.put(j, WK_WA + i, W.get(j, WK_WA + i) + scale * (t[i] + pf[i]));
    W.put(j, WK_WA + 3 + i, W.get(j, WK_WA + 3 + i) + scale * f[i]);
  }
}

// ------------------------------------------------------------------ cuboid SDF face pick
// (DH/Body/BodyCuboid.cpp:162-173: strict '>' scanned +x,-x,+y,-y,+z,-z)
HD double cuboid_face(const double* x, const double* hs, int& axis, double& sgn) {
  double d = -9999999.0;
  axis = 0; sgn = 1.0;
  for (int i = 0; i < 3; ++i) {
    if (x[i] - hs[i] > d) { d = x[i] - hs[i]; axis = i; sgn = 1.0; }
    if (-x[i] - hs[i] > d) { d = -x[i] - hs[i]; axis = i; sgn = -1.0; }
  }
  return d;
}
HD double cuboid_distance(const double* x, const double* hs) {
  double d = -99999999.0;
  for (int i = 0; i < 3; ++i) { d = fmax(d, fmax(x[i] - hs[i], -x[i] - hs[i])); }
  return d;
}
// distance(x) < 0 of DH/Body/BodyCuboid.cpp:135-144, without forming the distance:
// max_i(|x_i| - h_i) < 0  <=>  |x_i| < h_i for all i  (a - b < 0 <=> a < b in IEEE arithmetic)
HD bool cuboid_inside(const double* x, const double* hs) {
  return !(fabs(x[0]) >= hs[0]) && !(fabs(x[1]) >= hs[1]) && !(fabs(x[2]) >= hs[2]);
}
// Cheap conservative classification of a point given in the box frame through the PRECOMPOSED relative
// transform (x = R21 xi + r differs from the reference's evaluation order by rounding, ~1e-17 m):
// +1 surely inside, -1 surely outside, 0 too close to a face to tell -> caller runs the exact test.
#define TS_BAND 1e-12
HD int cuboid_classify(const double* R21, const double* r, const double* xi, const double* hs) {
  double x[3];
  mv3(R21, xi, x);
  const double m0 = hs[0] - fabs(x[0] + r[0]), m1 = hs[1] - fabs(x[1] + r[1]), m2 = hs[2] - fabs(x[2] + r[2]);
  if (m0 < -TS_BAND || m1 < -TS_BAND || m2 < -TS_BAND) return -1;
  if (m0 > TS_BAND && m1 > TS_BAND && m2 > TS_BAND) return 1;
  return 0;
}
// relative transform of frame 1 in frame 2: R21 = R2^T R1, r = R2^T (p1 - p2)
HD void rel_frame(const double* R1, const double* p1, const double* R2, const double* p2, double* R21, double* r) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R21[3 * i + j] = R2[i] * R1[j] + R2[3 + i] * R1[3 + j] + R2[6 + i] * R1[6 + j];
  double d[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
  mtv3(R2, d, r);
}
// Exact-safe cull of a whole point set: its bounding box (lo, hi in frame 1), mapped into the box frame
// by the precomposed relative transform, lies strictly beyond one face plane of the box (with the same
// slack band as cuboid_classify) => by convexity no point of the set is inside the box.
HD bool bbox_outside_box(const double* R21, const double* r, const double* lo, const double* hi, const double* hs) {
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int cidx = 0; cidx < 8; ++cidx) {
    const double c[3] = {(cidx & 1) ? hi[0] : lo[0], (cidx & 2) ? hi[1] : lo[1], (cidx & 4) ? hi[2] : lo[2]};
    double x[3];
    mv3(R21, c, x);
    for (int i = 0; i < 3; ++i) { const double v = x[i] + r[i]; mn[i] = fmin(mn[i], v); mx[i] = fmax(mx[i], v); }
  }
  for (int i = 0; i < 3; ++i)
    if (mn[i] - hs[i] > TS_BAND || -mx[i] - hs[i] > TS_BAND) return true;
  return false;
}
HD int ts_ffs(unsigned m) {   // index of the lowest set bit (m != 0)
#ifdef __CUDA_ARCH__
  return __ffs((int)m) - 1;
#else
  return __builtin_ctz(m);
#endif
}
template <class T> HD void vals3(const T* a, double* o) { o[0] = val(a[0]); o[1] = val(a[1]); o[2] = val(a[2]); }
template <class T> HD void vals9(const T* a, double* o) { for (int i = 0; i < 9; ++i) o[i] = val(a[i]); }

// ------------------------------------------------------------------ contact forces
// ground plane vs sampled body points: DH/Force/ForceGroundContact.cpp:105-147,
// detection d <= 0: DH/CollisionDetection/CollisionDetection.cpp:13-42
template <class WK>
HDN void ground_contacts(const SceneView& S, WK& W) {
  typedef typename WK::Scalar T;
  const double h2 = W.beta;
  for (int gi = 0; gi < S.nground; ++gi) {
    const int* r = S.ib + S.o_ground + gi * KG_ISTRIDE;
    const double* c = S.db + S.d_ground + gi * KG_DSTRIDE;
    const int b = r[0], po = r[1], pc = r[2];
    const int jb = S.ib[S.o_body + b * KB_ISTRIDE];
    if (jb < 0) continue;
    const double kn = c[0], kt = c[1], mu = c[2], damp = c[3];
    T R[9], p[3], ph[6];
    body_frame(S, W, b, R, p, ph);
    double Rv[9], pv[3];
    vals9(R, Rv);
    vals3(p, pv);
    T wr[6];
    for (int i = 0; i < 6; ++i) wr[i] = 0.0;
    bool any = false;
    // a sphere has ONE contact point, found from the values of its pose and then held fixed in the body frame
    // while the force is differentiated (DH/CollisionDetection/CollisionDetection.cpp:17-25)
    double xs[3] = {0.0, 0.0, 0.0};
    const bool sph = KT_SPHERE && pc < 0;
    if (sph) {
      const double rad = S.db[S.d_body + b * KB_DSTRIDE + KB_HALF];
      const double dc = (S.gn[0] * (pv[0] - S.gx[0]) + S.gn[1] * (pv[1] - S.gx[1]) + S.gn[2] * (pv[2] - S.gx[2])) - rad;
      if (!(dc <= 0.0)) continue;
      double xw[3], tp[3];
      for (int i = 0; i < 3; ++i) xw[i] = pv[i] - S.gn[i] * rad;
      mtv3(Rv, xw, xs);                     // xi = E_i0 xw = R^T xw + (-(R^T p))
      mtv3(Rv, pv, tp);
      for (int i = 0; i < 3; ++i) xs[i] = xs[i] + (-tp[i]);
    }
    for (int k = 0; k < (sph ? 1 : pc); ++k) {
      const double* xi = sph ? xs : S.db + S.d_points + 3 * (po + k);
      double xv[3];
      mv3(Rv, xi, xv);
      const double dv = (xv[0] + pv[0] - S.gx[0]) * S.gn[0] + (xv[1] + pv[1] - S.gx[1]) * S.gn[1] + (xv[2] + pv[2] - S.gx[2]) * S.gn[2];
      if (!sph && !(dv <= 0.0)) continue;
      any = true;
      T xw[3];
      mv3(R, xi, xw);
      T d = (xw[0] + p[0] - S.gx[0]) * S.gn[0] + (xw[1] + p[1] - S.gx[1]) * S.gn[1] + (xw[2] + p[2] - S.gx[2]) * S.gn[2];
      T w[3], vw[3];
      cross3(ph, xi, w);
      w[0] = w[0] + ph[3]; w[1] = w[1] + ph[4]; w[2] = w[2] + ph[5];
      mv3(R, w, vw);
      T vn = dot3(vw, S.gn);
      T F[3], a[3];
      for (int i = 0; i < 3; ++i) {
        F[i] = -kn * S.gn[i] * d - damp * S.gn[i] * vn;
        a[i] = vw[i] - S.gn[i] * vn;
      }
      if (!(mu < TS_EPS)) {
        double an = sqrt(val(a[0]) * val(a[0]) + val(a[1]) * val(a[1]) + val(a[2]) * val(a[2]));
        if (mu * fabs(kn * val(d)) >= kt * an - TS_EPS) {
          for (int i = 0; i < 3; ++i) F[i] = F[i] - kt * a[i];
        } else {
          T anT = dsqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
          T sc = (mu * kn) * d / anT;
          for (int i = 0; i < 3; ++i) F[i] = F[i] + sc * a[i];
        }
      }
      T Fb[3], tq[3];
      mtv3(R, F, Fb);
      cross3(xi, Fb, tq);
      for (int i = 0; i < 3; ++i) { wr[i] = wr[i] + tq[i]; wr[3 + i] = wr[3 + i] + Fb[i]; }
    }
    if (any) push_wrench(W, jb, R, p, wr, -h2);
  }
}

// Relative kinematics of the contact pair in the BOX frame (dual numbers, once per evaluation) plus
// the values the exact face pick needs: everything one contact point evaluation reads.
template <class T> struct GpPair {
  T Q[9], rr[3], w1b[3], v1b[3], ph2[6];    // R21 = R2^T R1, r = R2^T (p1 - p2), pad twist in box coordinates, box twist
  double R1v[9], p1v[3], R2v[9], p2v[3];    // values of the two body frames (reference evaluation order of the face pick)
};

// Penalty force of NP active sampled points (DH/Force/ForceGeneralPrimitiveContact.cpp:154-229,
// DH/Body/BodyCuboid.cpp:146-184), accumulated as wrenches on body 1 / body 2, both in box coordinates
// about the box origin.
// The force of one point is a chain of ~360 dependent fp64 instructions that a warp issues at ~0.1 IPC (7 warps per SM,
// nothing to switch to): NP points are evaluated side by side, stage by stage, so that the instruction scheduler
// interleaves NP independent chains.  The arithmetic of each point and the order in which the wrenches are summed are
// those of the one-point evaluation: results do not depend on NP.
template <int NP, class T>
HD void gp_point_terms(const GpPair<T>& P, const double* const* xi1, const double* hs, double kn, double kt, double mu,
                       double damp, T (*tq2)[3], T (*Fb)[3], T (*tq1)[3]) {
  // face pick on the reference's evaluation order (BodyCuboid.cpp:162-173), values only
  double sg[NP], hsa[NP], e[NP][3];
  bool a0[NP], a1[NP], a2[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    double xwv[3], yv[3], xv[3];
    mv3(P.R1v, xi1[q], xwv);
    for (int i = 0; i < 3; ++i) yv[i] = (xwv[i] + P.p1v[i]) - P.p2v[i];
    mtv3(P.R2v, yv, xv);
    int ax;
    cuboid_face(xv, hs, ax, sg[q]);
    // the face axis is data: select instead of indexing so that every vector stays in registers
    a0[q] = (ax == 0); a1[q] = (ax == 1); a2[q] = (ax == 2);
    e[q][0] = a0[q] ? sg[q] : 0.0; e[q][1] = a1[q] ? sg[q] : 0.0; e[q][2] = a2[q] ? sg[q] : 0.0;
    hsa[q] = a0[q] ? hs[0] : (a1[q] ? hs[1] : hs[2]);
  }
  T x[NP][3], d[NP], tb[NP][3], s[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    T ap[3];
    mv3(P.Q, xi1[q], ap);                                // pad point relative to the pad origin, box coordinates
    for (int i = 0; i < 3; ++i) x[q][i] = ap[i] + P.rr[i];
    d[q] = sg[q] * (a0[q] ? x[q][0] : (a1[q] ? x[q][1] : x[q][2])) - hsa[q];
    // relative velocity in the box frame: u = R2^T xw_dot - w2 x x - v2
    T u[3], t3[3];
    cross3(P.w1b, ap, u);
    cross3(P.ph2, x[q], t3);
    for (int i = 0; i < 3; ++i) u[i] = ((u[i] + P.v1b[i]) - t3[i]) - P.ph2[3 + i];
    T ddot = sg[q] * (a0[q] ? u[0] : (a1[q] ? u[1] : u[2]));
    // tangential velocity in the box frame: (I - e e^T)(u + d w2 x e)
    cross3(P.ph2, e[q], t3);
    for (int i = 0; i < 3; ++i) tb[q][i] = u[i] + d[q] * t3[i];
    {
      T c0 = tb[q][0] - sg[q] * (sg[q] * tb[q][0]), c1 = tb[q][1] - sg[q] * (sg[q] * tb[q][1]), c2 = tb[q][2] - sg[q] * (sg[q] * tb[q][2]);
      tb[q][0] = a0[q] ? c0 : tb[q][0]; tb[q][1] = a1[q] ? c1 : tb[q][1]; tb[q][2] = a2[q] ? c2 : tb[q][2];
    }
    s[q] = kn * d[q] - damp * ddot * d[q];
    for (int i = 0; i < 3; ++i) Fb[q][i] = -(s[q] * e[q][i]);
  }
  if (mu > TS_EPS) {
    // the reference uses the norm of the 6-vector wrench on body 1 (:208): n1 = R1^T R2 e = row `ax` of
    // R21 times sg, m1 = xi1 x n1
    T n1[NP][3], m1[NP][3];
    bool stat[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      for (int i = 0; i < 3; ++i) n1[q][i] = sg[q] * (a0[q] ? P.Q[i] : (a1[q] ? P.Q[3 + i] : P.Q[6 + i]));
      cross3(xi1[q], n1[q], m1[q]);
      double n6 = 0.0;
      for (int i = 0; i < 3; ++i) n6 += val(m1[q][i]) * val(m1[q][i]) + val(n1[q][i]) * val(n1[q][i]);
      double fcn = fabs(val(s[q])) * sqrt(n6);
      double tn = sqrt(val(tb[q][0]) * val(tb[q][0]) + val(tb[q][1]) * val(tb[q][1]) + val(tb[q][2]) * val(tb[q][2]));
      stat[q] = mu * fcn >= kt * tn - TS_EPS;
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      if (stat[q]) {
        for (int i = 0; i < 3; ++i) Fb[q][i] = Fb[q][i] - kt * tb[q][i];
      } else {
        T n6T = m1[q][0] * m1[q][0] + m1[q][1] * m1[q][1] + m1[q][2] * m1[q][2] + n1[q][0] * n1[q][0] + n1[q][1] * n1[q][1] + n1[q][2] * n1[q][2];
        T fcT = dabs(s[q]) * dsqrt(n6T);
        T tnT = dsqrt(tb[q][0] * tb[q][0] + tb[q][1] * tb[q][1] + tb[q][2] * tb[q][2]);
        T sc = mu * fcT / tnT;
        for (int i = 0; i < 3; ++i) Fb[q][i] = Fb[q][i] - sc * tb[q][i];
      }
    }
  }
  // body 2 gets -Fb at the surface point xi2 = x - d e, body 1 gets +Fb at the pad point x
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    T xi2[3];
    for (int i = 0; i < 3; ++i) xi2[i] = x[q][i] - d[q] * e[q][i];
    cross3(xi2, Fb[q], tq2[q]);
    cross3(x[q], Fb[q], tq1[q]);
  }
}
// the terms of one point added to the wrenches on body 1 / body 2 (the ONE place that defines the summation order)
template <class T>
HD void gp_point_accumulate(const T* tq2, const T* Fb, const T* tq1, T* w1, T* w2) {
  for (int i = 0; i < 3; ++i) { w2[i] = w2[i] - tq2[i]; w2[3 + i] = w2[3 + i] - Fb[i]; }
  for (int i = 0; i < 3; ++i) { w1[i] = w1[i] + tq1[i]; w1[3 + i] = w1[3 + i] + Fb[i]; }
}
template <int NP, class T>
HD void gp_point_force(const GpPair<T>& P, const double* const* xi1, const double* hs, double kn, double kt, double mu,
                       double damp, T* w1, T* w2) {
  T tq2[NP][3], Fb[NP][3], tq1[NP][3];
  gp_point_terms<NP>(P, xi1, hs, kn, kt, mu, damp, tq2, Fb, tq1);
#pragma unroll
  for (int q = 0; q < NP; ++q) gp_point_accumulate(tq2[q], Fb[q], tq1[q], w1, w2);    // point by point, in order
}

// ---- cylinder SDF (DH/Body/BodyCylinder.cpp:88-139: radial distance only, no contact with the caps)
// distance(xw) < 0 with the reference's evaluation order: x = R2^T xw + (-(R2^T p2))  (E_i0 = Einv(E_0i))
HD bool cylinder_inside_world(const double* R2, const double* p2, const double* xw, const double* rh) {
  double t[3], x[3];
  mtv3(R2, p2, t);
  mtv3(R2, xw, x);
  for (int i = 0; i < 3; ++i) x[i] = x[i] + (-t[i]);
  if (x[2] < -rh[1]) return false;
  if (x[2] > rh[1]) return false;
  return sqrt(x[0] * x[0] + x[1] * x[1]) - rh[0] < 0.0;
}

// capsule SDF (DH/Body/BodyCapsule.cpp:126-138): x = E_i0 xw, distance to the segment (0,0,[-s,s]) minus the radius
HD bool capsule_inside_world(const double* R2, const double* p2, const double* xw, const double* rh) {
  double t[3], x[3];
  mtv3(R2, p2, t);
  mtv3(R2, xw, x);
  for (int i = 0; i < 3; ++i) x[i] = x[i] + (-t[i]);
  const double s = rh[1];
  double d;
  if (x[2] < -s) d = sqrt(x[0] * x[0] + x[1] * x[1] + (x[2] - (-s)) * (x[2] - (-s))) - rh[0];
  else if (x[2] > s) d = sqrt(x[0] * x[0] + x[1] * x[1] + (x[2] - s) * (x[2] - s)) - rh[0];
  else d = sqrt(x[0] * x[0] + x[1] * x[1]) - rh[0];
  return d < 0.0;
}

// sphere SDF (DH/Body/BodySphere.cpp:79-83): distance(xw) = |xw - p2| - radius < 0
HD bool sphere_inside_world(const double* p2, const double* xw, double radius) {
  const double a = xw[0] - p2[0], b = xw[1] - p2[1], c = xw[2] - p2[2];
  return sqrt(a * a + b * b + c * c) - radius < 0.0;
}

// Penalty force of ONE active sampled point against a CYLINDER (DH/Force/ForceGeneralPrimitiveContact.cpp:154-229
// with DH/Body/BodyCylinder.cpp:105-139): same structure as gp_point_force, the normal e = x_r / |x_r| now
// depends on the point.  Wrenches in cylinder coordinates about the cylinder origin.
// sph: SPHERE (DH/Body/BodySphere.cpp:79-107): the same with the 3-d radial normal e = x / |x|.
// caps > 0: CAPSULE of half-length caps (DH/Body/BodyCapsule.cpp:140-172): radial from the nearest point of the segment
// (0,0,[-caps, caps]) -- the cylinder formula between the caps, the sphere formula about (0,0,+-caps) beyond them.
template <class T>
HD void gp_point_force_cyl(const GpPair<T>& P, const double* xi1, const double* rh, double kn, double kt, double mu,
                           double damp, T* w1, T* w2, bool sph = false, double caps = -1.0) {
  T ap[3], x[3];
  mv3(P.Q, xi1, ap);
  for (int i = 0; i < 3; ++i) x[i] = ap[i] + P.rr[i];
  T u[3], t3[3];
  cross3(P.w1b, ap, u);
  cross3(P.ph2, x, t3);
  for (int i = 0; i < 3; ++i) u[i] = ((u[i] + P.v1b[i]) - t3[i]) - P.ph2[3 + i];
  T vz = 0.0;                            // axial component of the radial vector
  if (KT_SPHERE && sph) vz = x[2];
  if (KT_SPHERE && caps > 0.0) {
    if (val(x[2]) < -caps) vz = x[2] + caps;
    else if (val(x[2]) > caps) vz = x[2] - caps;
  }
  const bool zfree = KT_SPHERE && (sph || (caps > 0.0 && (val(x[2]) < -caps || val(x[2]) > caps)));
  T r = zfree ? dsqrt(x[0] * x[0] + x[1] * x[1] + vz * vz) : dsqrt(x[0] * x[0] + x[1] * x[1]);
  T e[3];
  e[0] = x[0] / r; e[1] = x[1] / r; e[2] = 0.0;
  if (zfree) e[2] = vz / r;
  T d = r - rh[0];
  T ddot = dot3(e, u);
  T tb[3];
  cross3(P.ph2, e, t3);
  for (int i = 0; i < 3; ++i) tb[i] = u[i] + d * t3[i];
  T et = dot3(e, tb);
  for (int i = 0; i < 3; ++i) tb[i] = tb[i] - e[i] * et;
  T s = kn * d - damp * ddot * d;
  T Fb[3];
  for (int i = 0; i < 3; ++i) Fb[i] = -(s * e[i]);
  if (mu > TS_EPS) {
    T n1[3], m1[3];
    mtv3(P.Q, e, n1);                       // normal in body-1 coordinates: R1^T R2 e
    cross3(xi1, n1, m1);
    double n6 = 0.0;
    for (int i = 0; i < 3; ++i) n6 += val(m1[i]) * val(m1[i]) + val(n1[i]) * val(n1[i]);
    double fcn = fabs(val(s)) * sqrt(n6);
    double tn = sqrt(val(tb[0]) * val(tb[0]) + val(tb[1]) * val(tb[1]) + val(tb[2]) * val(tb[2]));
    if (mu * fcn >= kt * tn - TS_EPS) {
      for (int i = 0; i < 3; ++i) Fb[i] = Fb[i] - kt * tb[i];
    } else {
      T n6T = m1[0] * m1[0] + m1[1] * m1[1] + m1[2] * m1[2] + n1[0] * n1[0] + n1[1] * n1[1] + n1[2] * n1[2];
      T fcT = dabs(s) * dsqrt(n6T);
      T tnT = dsqrt(tb[0] * tb[0] + tb[1] * tb[1] + tb[2] * tb[2]);
      T sc = mu * fcT / tnT;
      for (int i = 0; i < 3; ++i) Fb[i] = Fb[i] - sc * tb[i];
    }
  }
  T xi2[3], tq[3];
  for (int i = 0; i < 3; ++i) xi2[i] = x[i] - d * e[i];
  cross3(xi2, Fb, tq);
  for (int i = 0; i < 3; ++i) { w2[i] = w2[i] - tq[i]; w2[3 + i] = w2[3 + i] - Fb[i]; }
  cross3(x, Fb, tq);
  for (int i = 0; i < 3; ++i) { w1[i] = w1[i] + tq[i]; w1[3 + i] = w1[3 + i] + Fb[i]; }
}

// ---- block-cooperative evaluation of the active cuboid contact points (fwd_kernel of the variants that enable it)
// In a lock-step block every round costs what its slowest warp costs, and a tile in contact runs the point force
// ~9 times in a row while the other 27 tiles of the block wait at the next vote.  Here the tiles in contact PUBLISH
// their pair kinematics (GpPair: 24 dual numbers + 24 frame values) and their list of active points in shared memory,
// the active points of the whole block form one item list (tile by tile, points ascending), every tile of the block
// evaluates the items dealt to it (item i of a batch -> tile i) with the publisher's kinematics and writes the
// per-point terms (tq2, Fb, tq1) back, and each publisher sums the terms of its own points IN POINT ORDER with the
// arithmetic of the serial evaluation -- same numbers, bit for bit, whoever computed them.
// Block-wide barriers inside: every thread of the block calls this the same number of times (idle tiles with mine = false).
template <bool COOP> struct CoopTag {};
// TS_GP_ILP: active cuboid points evaluated side by side by one lane (gp_point_terms<NP>).  MEASURED on B200 (forward
// call, B=4096, T=200): 1 -> 103.0 ms, 2 -> 131.2, 3 -> 142.4, 4 -> 153.0: the second chain does not fit in the 255
// registers next to the pair kinematics and spills; one point at a time it is.
#ifndef TS_GP_ILP
#define TS_GP_ILP 1
#endif
#if defined(TS_COOP_GP) && defined(__CUDACC__)
// Shared-memory budget: the step loop's block holds the scene tables, 28 tile regions (96 KB) and this area; with at
// most TS_COOP_SLOTS = 22 publishers per phase (of 28 tiles; on average 8 are in contact) and byte-sized point lists the
// block stays below the 164 KB shared-memory configuration, which leaves 92 KB of L1 for the per-lane tangents in local
// memory instead of 60 KB (forcing the L1 down to 28 KB cost 7 %, profiles/r02_experiments.md).  A tile beyond the
// slots (never seen in the bench workload) takes part in the barriers and in the work and sums its own points serially.
#ifndef TS_COOP_SLOTS
#define TS_COOP_SLOTS 22
#endif
template <int LPE> struct CoopArea {
  static const int NT = TS_BLOCK / LPE;           // tiles per block = items per batch
  static const int NS = TS_COOP_SLOTS < NT ? TS_COOP_SLOTS : NT;   // publishers per phase
  static const int MAXP = KT_MAXPW * 32;          // sampled points of a general body (fits a byte: static_assert below)
  int cnt[NT];                                   // active points of each tile for the force in flight
  int soff[NS];                                  // first item of each publisher (slot)
  unsigned char owner[NS * MAXP];                // slot of each item
  unsigned char pts[NS][MAXP];                   // active point indices of each slot, ascending
  double Pv[NS][48];                             // values: Q rr w1b v1b ph2 (24), then R1v p1v R2v p2v (24)
  double Pt[NS][24][LPE];                        // tangents of the first 24, by component and lane
  double Rv[NT][9];                              // per-item terms tq2, Fb, tq1: values ...
  double Rt[NT][9][LPE];                         // ... and tangents by lane
};
template <class Tile, class T>
__device__ __forceinline__ bool gp_points_coop(const Tile& tl, CoopTag<true>, const SceneView& S, GpPair<T>& P, const unsigned* act,
                                               bool mine, int po, const double* hs, double kn, double kt, double mu,
                                               double damp, T* w1, T* w2) {
  typedef CoopArea<Tile::LPE> CA;
  static_assert(CA::MAXP <= 256, "point indices of the cooperative lists are bytes");
  CA& C = *(CA*)tl.coop;
  const int my = tl.tile_id, lane = tl.lane;
  int cnt = 0;
  if (mine) for (int w = 0; w < KT_MAXPW; ++w) cnt += __popc(act[w]);
  TS_CPT0();
  // (the counts are published before the vote: its barrier orders them; nobody reads cnt[] of a phase after that
  // phase's publish barrier, so the next phase may overwrite it)
  if (lane == 0) C.cnt[my] = cnt;
  if (!tl.cta_or_unaligned(cnt > 0)) return false;              // nobody in this block touches: one barrier
  // publishers in tile order take the slots; off / total count the items of the publishers only
  int off = 0, total = 0, slot = -1, nsl = 0;
#pragma unroll
  for (int t = 0; t < CA::NT; ++t) {
    const int c = C.cnt[t];
    const bool pub = c > 0 && nsl < CA::NS;
    if (t == my) { off = total; slot = pub ? nsl : -1; }
    if (pub) { total += c; ++nsl; }
  }
  if (slot >= 0) {
    if (lane == 0) {
      C.soff[slot] = off;
      int r = 0;
      for (int w = 0; w < KT_MAXPW; ++w) {
        unsigned m = act[w];
        while (m) { C.owner[off + r] = (unsigned char)slot; C.pts[slot][r++] = (unsigned char)(32 * w + ts_ffs(m)); m &= m - 1; }
      }
    }
    const T* pt = P.Q;                            // Q rr w1b v1b ph2: 24 contiguous scalars
    const double* pv = P.R1v;                     // R1v p1v R2v p2v: 24 contiguous doubles
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      C.Pt[slot][i][lane] = tan_of(pt[i]);
      if (lane == (i % Tile::LPE)) { C.Pv[slot][i] = val(pt[i]); C.Pv[slot][24 + i] = pv[i]; }
    }
  }
  tl.cta_sync_unaligned();
  TS_CPT(tl, 11);
  TS_CPN(tl, 8, total); TS_CPN(tl, 9, 1);
  for (int base = 0; base < total; base += CA::NT) {
    TS_CPN(tl, 10, 1);
    const int item = base + my;
    if (item < total) {
      const int src = C.owner[item];              // publisher (slot) of the item
      const int k = C.pts[src][item - C.soff[src]];
      GpPair<T> Q;
      T* qt = Q.Q;
      double* qv = Q.R1v;
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        qt[i] = Lift<T>::mk(C.Pv[src][i], C.Pt[src][i][lane]);
        qv[i] = C.Pv[src][24 + i];
      }
      const double* xi1 = S.db + S.d_points + 3 * (po + k);
      T tq2[1][3], Fb[1][3], tq1[1][3];
      gp_point_terms<1>(Q, &xi1, hs, kn, kt, mu, damp, tq2, Fb, tq1);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        C.Rt[my][i][lane] = tan_of(tq2[0][i]); C.Rt[my][3 + i][lane] = tan_of(Fb[0][i]); C.Rt[my][6 + i][lane] = tan_of(tq1[0][i]);
        if (lane == 0) { C.Rv[my][i] = val(tq2[0][i]); C.Rv[my][3 + i] = val(Fb[0][i]); C.Rv[my][6 + i] = val(tq1[0][i]); }
      }
    }
    TS_CPT(tl, 12);
    tl.cta_sync_unaligned();
    TS_CPT(tl, 13);
    if (slot >= 0) {
      const int lo = off > base ? off : base, hi = (off + cnt < base + CA::NT) ? off + cnt : base + CA::NT;
      // (two items per pass: the loads of the second overlap the sums of the first; the order of the sums is unchanged)
      int it = lo;
      for (; it + 1 < hi; it += 2) {
        const int r = it - base;
        T a[9], b2[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { a[i] = Lift<T>::mk(C.Rv[r][i], C.Rt[r][i][lane]); b2[i] = Lift<T>::mk(C.Rv[r + 1][i], C.Rt[r + 1][i][lane]); }
        gp_point_accumulate(a, a + 3, a + 6, w1, w2);
        gp_point_accumulate(b2, b2 + 3, b2 + 6, w1, w2);
      }
      if (it < hi) {
        const int r = it - base;
        T a[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) a[i] = Lift<T>::mk(C.Rv[r][i], C.Rt[r][i][lane]);
        gp_point_accumulate(a, a + 3, a + 6, w1, w2);
      }
    }
    TS_CPT(tl, 14);
    tl.cta_sync_unaligned();
    TS_CPT(tl, 13);
  }
  return slot >= 0;                               // false: no points, or no slot left -> the caller's serial loop
}
#endif
template <class Tile, class T>
HD bool gp_points_coop(const Tile&, CoopTag<false>, const SceneView&, GpPair<T>&, const unsigned*, bool, int, const double*,
                       double, double, double, double, T*, T*) { return false; }

// sampled points of a general body vs a primitive SDF (cuboid or cylinder):
// DH/Force/ForceGeneralPrimitiveContact.cpp:154-229, DH/Body/BodyCuboid.cpp:146-184, BodyCylinder.cpp:105-139,
// detection d < 0: CollisionDetection.cpp:66-83.
// Detection is dealt to the lanes of the tile and gathered by ballot.  (A warp-cooperative evaluation of the
// active points -- pair kinematics by shuffles, points dealt to the tiles of the warp, xor-tree reduction -- was
// measured slower on B200 twice: 144 vs 125-140 ms, the pair structs end up in local memory; kept as
// tools/experiments/warp_cooperative_contacts_v2.patch.)
// idle: this tile has no evaluation of its own in flight; it only takes part in the cooperative point evaluation of its block
template <class Tile, class WK>
HDN void gp_contacts(const Tile& tl, const SceneView& S, WK& W, bool idle = false) {
  typedef typename WK::Scalar T;
  const int L = Tile::LPE;
  const double h2 = W.beta;
  for (int fi = 0; fi < S.ngp; ++fi) {
    const int* r = S.ib + S.o_gp + fi * KP_ISTRIDE;
    const double* c = S.db + S.d_gp + fi * KP_DSTRIDE;
    const int b1 = r[0], b2 = r[1], po = r[2], pc = r[3];
    const bool sph = KT_SPHERE && r[5] == TS_SH_SPHERE;
    const bool cap = KT_SPHERE && r[5] == TS_SH_CAPSULE;
    const bool cyl = (KT_CYLINDER && r[5] == TS_SH_CYLINDER) || sph || cap;      // point-dependent normal
    const int j1 = S.ib[S.o_body + b1 * KB_ISTRIDE], j2 = S.ib[S.o_body + b2 * KB_ISTRIDE];
    const double kn = c[0], kt = c[1], mu = c[2], damp = c[3];
    const double* bd2 = S.db + S.d_body + b2 * KB_DSTRIDE;
    const double* hs = bd2 + KB_HALF;
    TS_GPT0();
    GpPair<T> P;
    double phv[6];
    body_frame_v(S, W, b1, P.R1v, P.p1v, phv);
    body_frame_v(S, W, b2, P.R2v, P.p2v, phv);
    unsigned act[KT_MAXPW];
    for (int i = 0; i < KT_MAXPW; ++i) act[i] = 0u;
    if (!idle) {
      // exact-safe culls on values: bounding spheres, then (cuboid) the bounding box of the point set against
      // the face planes of the box
      const double rr = c[4] + bd2[KB_RBOUND] + TS_CULL_MARGIN;
      const double dx = P.p1v[0] - P.p2v[0], dy = P.p1v[1] - P.p2v[1], dz = P.p1v[2] - P.p2v[2];
      double R21[9], r21[3];
      bool maybe = !(dx * dx + dy * dy + dz * dz > rr * rr);
      if (maybe && !cyl) {
        rel_frame(P.R1v, P.p1v, P.R2v, P.p2v, R21, r21);
        maybe = !bbox_outside_box(R21, r21, c + KP_BBOX, c + KP_BBOX + 3, hs);
      }
      // detection (values only), points dealt to the lanes of the tile, results gathered by ballot
      if (KT_SPHERE && maybe && sph) {
        // dense point sets against a sphere, word by word: the centre in the body-1 frame against the box of each
        // 32 consecutive points (exact-safe with the cull margin); only words the sphere can reach are tested
        const double* wbox = S.db + r[KP_WBOX];
        double dc[3], c1[3];
        for (int i = 0; i < 3; ++i) dc[i] = P.p2v[i] - P.p1v[i];
        mtv3(P.R1v, dc, c1);
        const double rs = hs[0] + TS_CULL_MARGIN;
        for (int w0 = 0; w0 < pc; w0 += 32) {
          const double* bx = wbox + 6 * (w0 >> 5);
          double d2 = 0.0;
          for (int i = 0; i < 3; ++i) {
            const double lo = bx[i] - c1[i], hi = c1[i] - bx[3 + i];
            const double e = lo > 0.0 ? lo : (hi > 0.0 ? hi : 0.0);
            d2 += e * e;
          }
          if (d2 > rs * rs) continue;
          for (int base = w0; base < pc && base < w0 + 32; base += L) {
            const int k = base + tl.lane;
            bool in = false;
            if (k < pc && k < w0 + 32) {
              const double* xi1 = S.db + S.d_points + 3 * (po + k);
              double xwv[3];
              mv3(P.R1v, xi1, xwv);
              for (int i = 0; i < 3; ++i) xwv[i] = xwv[i] + P.p1v[i];
              in = sphere_inside_world(P.p2v, xwv, hs[0]);
            }
            const unsigned bits = tl.ballot(in);
            act[base >> 5] |= bits << (base & 31);
          }
        }
      } else if (maybe) {
        for (int base = 0; base < pc; base += L) {
          const int k = base + tl.lane;
          bool in = false;
          if (k < pc) {
            const double* xi1 = S.db + S.d_points + 3 * (po + k);
            if (cyl) {
              double xwv[3];
              mv3(P.R1v, xi1, xwv);
              for (int i = 0; i < 3; ++i) xwv[i] = xwv[i] + P.p1v[i];
              in = sph ? sphere_inside_world(P.p2v, xwv, hs[0])
                       : (cap ? capsule_inside_world(P.R2v, P.p2v, xwv, hs) : cylinder_inside_world(P.R2v, P.p2v, xwv, hs));
            } else {
              const int cls = cuboid_classify(R21, r21, xi1, hs);
              if (cls > 0) in = true;
              else if (cls == 0) {                 // reference evaluation order (CollisionDetection.cpp:73-79)
                double xwv[3], yv[3], xv[3];
                mv3(P.R1v, xi1, xwv);
                for (int i = 0; i < 3; ++i) yv[i] = (xwv[i] + P.p1v[i]) - P.p2v[i];
                mtv3(P.R2v, yv, xv);
                in = cuboid_inside(xv, hs);
              }
            }
          }
          const unsigned bits = tl.ballot(in);
          act[base >> 5] |= bits << (base & 31);
        }
      }
    }
    unsigned any = 0u;
    for (int i = 0; i < KT_MAXPW; ++i) any |= act[i];
    TS_GPT(tl, 0);
    const bool coop = Tile::COOP && !cyl;          // (block-uniform: a property of the force)
    if (!any && !coop) continue;
    // the tile's own relative kinematics in the frame of body 2 (dual numbers)
    T R2[9], p2[3];
    if (any) {
      W.gp_any = 1;
      T R1[9], p1[3], ph1[6], dp[3];
      body_frame(S, W, b1, R1, p1, ph1);
      body_frame(S, W, b2, R2, p2, P.ph2);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) P.Q[3 * i + j] = R2[i] * R1[j] + R2[3 + i] * R1[3 + j] + R2[6 + i] * R1[6 + j];
      for (int i = 0; i < 3; ++i) dp[i] = p1[i] - p2[i];
      mtv3(R2, dp, P.rr);
      mv3(P.Q, ph1, P.w1b);
      mv3(P.Q, ph1 + 3, P.v1b);
    }
    TS_GPT(tl, 1);
    T w1[6], w2[6];          // wrenches on body 1 / body 2, both in body-2 coordinates about its origin
    for (int i = 0; i < 6; ++i) { w1[i] = 0.0; w2[i] = 0.0; }
    if (coop && gp_points_coop(tl, CoopTag<Tile::COOP>(), S, P, act, any != 0u, po, hs, kn, kt, mu, damp, w1, w2)) {
      push_wrench(W, j1, R2, p2, w1, -h2);
      push_wrench(W, j2, R2, p2, w2, -h2);
      continue;
    }
    if (!any) continue;
    const double* pend[TS_GP_ILP];           // active cuboid points waiting to be evaluated side by side
    int npend = 0;
    for (int wd = 0; wd < KT_MAXPW; ++wd) {
      unsigned m = act[wd];
      while (m) {
        const int k = 32 * wd + ts_ffs(m);
        m &= m - 1;
        const double* xi1 = S.db + S.d_points + 3 * (po + k);
        if (cyl) gp_point_force_cyl(P, xi1, hs, kn, kt, mu, damp, w1, w2, sph, cap ? hs[1] : -1.0);
        else {
          pend[npend++] = xi1;
          if (npend == TS_GP_ILP) { gp_point_force<TS_GP_ILP>(P, pend, hs, kn, kt, mu, damp, w1, w2); npend = 0; }
        }
      }
    }
#if TS_GP_ILP > 1
    for (int i = 0; i < npend; ++i) gp_point_force<1>(P, pend + i, hs, kn, kt, mu, damp, w1, w2);
#endif
    push_wrench(W, j1, R2, p2, w1, -h2);
    push_wrench(W, j2, R2, p2, w2, -h2);
    TS_GPT(tl, 3);
  }
}

// World axes of the six coordinates of a free3d-euler joint whose frame is (R0, .): translations along
// Ra e_i = R0 (row i of R(r)), rotations about R0 T_i with T = R^T G = [[c2c3, s3, 0], [-c2s3, c3, 0], [s2, 0, 1]]
// (the S_j of DH/Joint/JointSphericalEuler.cpp:62-66).
template <class T>
HD void euler_world_axes(const T* R0, T r1, T r2, T r3, T (*tax)[3], T (*rax)[3]) {
  T s1, c1, s2, c2, s3, c3;
  dsincos(r1, s1, c1);
  dsincos(r2, s2, c2);
  dsincos(r3, s3, c3);
  T rows[3][3], tc[3][3];
  rows[0][0] = c2 * c3; rows[0][1] = -(c2 * s3); rows[0][2] = s2;
  rows[1][0] = c1 * s3 + c3 * (s1 * s2); rows[1][1] = c1 * c3 - (s1 * s2) * s3; rows[1][2] = -(c2 * s1);
  rows[2][0] = s1 * s3 - (c1 * c3) * s2; rows[2][1] = c3 * s1 + (c1 * s2) * s3; rows[2][2] = c1 * c2;
  tc[0][0] = c2 * c3; tc[0][1] = -(c2 * s3); tc[0][2] = s2;
  tc[1][0] = s3; tc[1][1] = c3; tc[1][2] = 0.0;
  tc[2][0] = 0.0; tc[2][1] = 0.0; tc[2][2] = 1.0;
  for (int i = 0; i < 3; ++i) { mv3(R0, rows[i], tax[i]); mv3(R0, tc[i], rax[i]); }
}

// position-controlled motor on one coordinate: PD on the state at the START of the step, clamped
// (DH/Actuator/ActuatorMotor.cpp:37-41); dfdu / dfdq0 / dfdqd0 are P / -P / -D while unclamped (:56-73)
template <class T>
HD T pos_motor_force(double u, T q0, T qd0, double P, double D, double cmin, double cmax) {
  T f = P * (u - q0) + D * (-qd0);
  if (val(f) > cmax) return T(cmax);
  if (val(f) < cmin) return T(cmin);
  return f;
}

// FORCE motor: clamp, affine map to ctrl_range, clamp  (DH/Actuator/ActuatorMotor.cpp:31-36, Utils.h:384-386)
HD double motor_force(double u, double cmin, double cmax) {
  double uc = fmax(fmin(u, 1.0), -1.0);
  double f = (uc - (-1.0)) * (1.0 / (1.0 - (-1.0))) * (cmax - cmin) + cmin;
  return fmax(fmin(f, cmax), cmin);
}

// Inward sweep: g = S^T (subtree wrench) - h^2 (joint damping + limit springs + motors)
template <class WK, class In>
HDN void inward(const SceneView& S, WK& W, const In& in, const double* u, typename WK::Scalar* g) {
  typedef typename WK::Scalar T;
  const double h2 = W.beta;
  for (int j = S.nj - 1; j >= 0; --j) {
    const int* ji = S.ib + S.o_joint + j * KJ_ISTRIDE;
    const double* jd = S.db + S.d_joint + j * KJ_DSTRIDE;
    const int jt = ji[0], par = ji[1], qo = ji[2], nd = ji[3];
    T A[6], R0[9], p0[3];
    wk_ld(W, j, WK_WA, 6, A);
    wk_ld(W, j, WK_R0, 9, R0);
    wk_ld(W, j, WK_P0, 3, p0);
    const double* a0 = jd + KJ_AX0;
    const double* a1 = jd + KJ_AX1;
    T fj[3];
    mtv3(R0, A + 3, fj);                 // force in joint coordinates
    if (jt == TS_JT_REVOLUTE) {
      T pf[3], t[3], nj3[3];
      cross3(p0, A + 3, pf);             // moment about the joint origin
      for (int i = 0; i < 3; ++i) t[i] = A[i] - pf[i];
      mtv3(R0, t, nj3);
      g[qo] = dot3(a0, nj3);
    } else if (jt == TS_JT_PRISMATIC) g[qo] = dot3(a0, fj);
    else if (jt == TS_JT_PLANAR) { g[qo] = dot3(a0, fj); g[qo + 1] = dot3(a1, fj); }
    else if (jt == TS_JT_TRANSLATIONAL) { g[qo] = fj[0]; g[qo + 1] = fj[1]; g[qo + 2] = fj[2]; }
    else if ((KT_FREE3D && (jt == TS_JT_FREE3D_EULER || jt == TS_JT_SPHERICAL_EULER)) ||
             (KT_EXP3D && (jt == TS_JT_FREE3D_EXP || jt == TS_JT_SPHERICAL_EXP))) {
      const int nt = (jt == TS_JT_FREE3D_EULER || jt == TS_JT_FREE3D_EXP) ? 3 : 0;
      const int ro = qo + nt;
      T tax[3][3], rax[3][3], pf[3], t[3];
      if (KT_EXP3D && (jt == TS_JT_FREE3D_EXP || jt == TS_JT_SPHERICAL_EXP)) exp_world_axes(R0, in.q(ro), in.q(ro + 1), in.q(ro + 2), tax, rax);
      else euler_world_axes(R0, in.q(ro), in.q(ro + 1), in.q(ro + 2), tax, rax);
      cross3(p0, A + 3, pf);             // moment about the joint origin
      for (int i = 0; i < 3; ++i) t[i] = A[i] - pf[i];
      for (int i = 0; i < 3; ++i) { if (nt) g[qo + i] = dot3(tax[i], A + 3); g[ro + i] = dot3(rax[i], t); }
    } else if (KT_FREE3D && jt == TS_JT_FREE2D) {
      // translations along Ra e_x, Ra e_y = R0 (rows 0, 1 of Rz(theta)); rotation about Ra e_z = R0 e_z through p0
      T s3, c3, pf[3], t[3];
      dsincos(in.q(qo + 2), s3, c3);
      T r0[3] = {c3, -s3, T(0.0)}, r1[3] = {s3, c3, T(0.0)}, ax0[3], ax1[3];
      mv3(R0, r0, ax0);
      mv3(R0, r1, ax1);
      cross3(p0, A + 3, pf);
      for (int i = 0; i < 3; ++i) t[i] = A[i] - pf[i];
      g[qo] = dot3(ax0, A + 3);
      g[qo + 1] = dot3(ax1, A + 3);
      g[qo + 2] = R0[2] * t[0] + R0[5] * t[1] + R0[8] * t[2];
    }
    // joint damping and one-sided limit springs                (DH/Joint/Joint.cpp:251-263)
    const double damp = jd[KJ_DAMP], lo = jd[KJ_LIMLO], hi = jd[KJ_LIMHI], lk = jd[KJ_LIMK];
    for (int i = 0; i < nd; ++i) {
      T fr = -(damp * in.qd(qo + i));
      const T qi = in.q(qo + i);
      if (val(qi) < lo) fr = fr + lk * (lo - qi);
      if (val(qi) > hi) fr = fr + lk * (hi - qi);
      g[qo + i] = g[qo + i] - h2 * fr;
    }
    if (par >= 0) for (int i = 0; i < 6; ++i) W.put(par, WK_WA + i, W.get(par, WK_WA + i) + A[i]);
  }
  for (int ai = 0; ai < S.nact; ++ai) {
    const int* r = S.ib + S.o_act + ai * KA_ISTRIDE;
    const double* c = S.db + S.d_act + ai * KA_DSTRIDE;
    const int qo = S.ib[S.o_joint + r[0] * KJ_ISTRIDE + 2];
    if (KT_POS_MOTOR && r[1] == TS_ACT_POS) {
      for (int i = 0; i < r[3]; ++i)
        g[qo + i] = g[qo + i] - h2 * pos_motor_force(u[r[2] + i], in.q0(qo + i), in.qd0(qo + i), c[6 + i], c[9 + i], c[i], c[3 + i]);
    } else {
      for (int i = 0; i < r[3]; ++i) g[qo + i] = g[qo + i] - h2 * motor_force(u[r[2] + i], c[i], c[3 + i]);
    }
  }
}

// residual g = M(q) dl - beta f(q, qd) of one implicit stage.  BDF1: qd = (q - q0) / h, dl = q - q0 - h qd0,
// beta = h^2 (DH/Simulation.cpp:1227-1235); the other integrators only change (qd, dl, beta): stage_inputs().
template <class Tile, class WK, class In>
HDN void eval_g(const Tile& tl, const SceneView& S, const In& in, const double* u, WK& W, typename WK::Scalar* g,
                double beta, bool idle = false) {
  // (idle: Tile::COOP only -- nothing of its own to evaluate, the tile helps the block with its contact points; the
  // general-primitive phase has ONE call site so that the lanes of a warp run their shares of the points together)
  const bool own = !(Tile::COOP && idle);
  if (own) {
    W.beta = beta;
    W.gp_any = 0;
  }
#ifdef TS_PROFILE_GP      // slots 0, 1, 3 are re-used for the phases INSIDE gp_contacts (detect / pair setup / point forces)
  if (own) { kinematics(S, in, W, true); ground_contacts(S, W); }
  { TS_TIC(tl); gp_contacts(tl, S, W, !own); TS_TOC(tl, 2); }
  if (own) inward(S, W, in, u, g);
#else
  if (own) { TS_TIC(tl); kinematics(S, in, W, true); TS_TOC(tl, 0); }
  if (own) { TS_TIC(tl); ground_contacts(S, W); TS_TOC(tl, 1); }
  { TS_TIC(tl); gp_contacts(tl, S, W, !own); TS_TOC(tl, 2); }
  if (own) { TS_TIC(tl); inward(S, W, in, u, g); TS_TOC(tl, 3); }
#endif
}

// ------------------------------------------------------------------ tile policies
struct HostTile {
  static const int LPE = 1;
  static const bool COOP = false;
  static const bool LS_SERVICE = false;
  int lane;
#ifdef TS_PROFILE
  mutable long long acc[16];
#endif
  HD HostTile() : lane(0) {}
  HD double bcast(double v, int) const { return v; }
  HD int bcasti(int v, int) const { return v; }
  HD double sum(double v) const { return v; }
  HD unsigned ballot(bool p) const { return p ? 1u : 0u; }
  HD void cta_sync() const {}
  HD void tile_sync() const {}
  HD bool cta_any(bool p) const { return p; }
  HD bool warp_all(bool p) const { return p; }
  HD bool warp_any(bool p) const { return p; }
  template <int SUBL> HD HostTile sub() const { return *this; }
  // whole-warp helpers (one tile per "warp" on the host)
  static const int TPW = 1;
  HD int tile_in_warp() const { return 0; }
  HD unsigned tiles_ballot(bool p) const { return p ? 1u : 0u; }
  HD double warp_shfl(double v, int) const { return v; }
  HD unsigned warp_shfl_u(unsigned v, int) const { return v; }
  HD double sum_tiles(double v) const { return v; }
};

#define TS_NC(LPE) ((TS_MAXN + (LPE)-1) / (LPE))

// Partial-pivot LU solve of the n x n system whose columns are dealt round-robin to the lanes
// (column k lives in lane k % LPE at slot k / LPE); rhs is replicated and overwritten with the
// solution.  Pivot = first row of maximal |a| (Eigen partialPivLu, DH/Simulation.cpp:1178).
// The matrix is first all-gathered (one shuffle per entry), then every lane factors its own copy
// in registers with compile-time indices: no cross-lane traffic inside the elimination, and the
// solution comes out replicated, which is how the Newton update needs it.
// PIVOT=false is the same elimination without row exchanges; it reports whether partial pivoting
// would have exchanged any row (then the caller redoes the solve with PIVOT=true).
// The GPU kernels do NOT run this any more (tiles with one lane per row: lu_rows_solve_pivot below); it serves tiles
// with fewer lanes than dofs -- the host harness of the CPU tests -- and the A/B builds with -DTS_NO_LU_PIVOT.
#if TS_MAXN > 8
// 16-dof variant: plain loops, the matrix lives in local memory
template <bool PIVOT>
HDN bool lu_factor_solve(double (*A)[TS_MAXN], double* b) {
  bool exchanged = false;
  for (int j = 0; j < TS_MAXN; ++j) {
    int p = j;
    double best = fabs(A[j][j]);
    for (int i = j + 1; i < TS_MAXN; ++i) {
      const double v = fabs(A[i][j]);
      if (v > best) { best = v; p = i; }
    }
    if (p != j) {
      if (!PIVOT) exchanged = true;
      else {
        for (int c = 0; c < TS_MAXN; ++c) { const double t = A[j][c]; A[j][c] = A[p][c]; A[p][c] = t; }
        const double t = b[j]; b[j] = b[p]; b[p] = t;
      }
    }
    const double piv = A[j][j];
    for (int i = j + 1; i < TS_MAXN; ++i) {
      const double l = A[i][j] / piv;
      for (int c = j + 1; c < TS_MAXN; ++c) A[i][c] -= l * A[j][c];
      b[i] -= l * b[j];
    }
  }
  for (int k = TS_MAXN - 1; k >= 0; --k) {
    const double xk = b[k] / A[k][k];
    b[k] = xk;
    for (int i = 0; i < k; ++i) b[i] -= A[i][k] * xk;
  }
  return exchanged;
}
#else
template <bool PIVOT>
HD bool lu_factor_solve(double (*A)[TS_MAXN], double* b) {
  bool exchanged = false;
#pragma unroll
  for (int j = 0; j < TS_MAXN; ++j) {
    int p = j;
    double best = fabs(A[j][j]);
#pragma unroll
    for (int i = j + 1; i < TS_MAXN; ++i) {
      const double v = fabs(A[i][j]);
      if (v > best) { best = v; p = i; }
    }
    if (PIVOT) {
#pragma unroll
      for (int i = j + 1; i < TS_MAXN; ++i) {
        const bool sw = (p == i);
#pragma unroll
        for (int c = j; c < TS_MAXN; ++c) {
          const double t = A[j][c], s2 = A[i][c];
          A[j][c] = sw ? s2 : t;
          A[i][c] = sw ? t : s2;
        }
        const double t = b[j], s2 = b[i];
        b[j] = sw ? s2 : t;
        b[i] = sw ? t : s2;
      }
    } else if (p != j) {
      exchanged = true;
    }
    const double piv = A[j][j];
#pragma unroll
    for (int i = j + 1; i < TS_MAXN; ++i) {
      const double l = A[i][j] / piv;
#pragma unroll
      for (int c = j + 1; c < TS_MAXN; ++c) A[i][c] -= l * A[j][c];
      b[i] -= l * b[j];
    }
  }
#pragma unroll
  for (int k = TS_MAXN - 1; k >= 0; --k) {
    const double xk = b[k] / A[k][k];
    b[k] = xk;
#pragma unroll
    for (int i = 0; i < k; ++i) b[i] -= A[i][k] * xk;
  }
  return exchanged;
}
#endif

template <class Tile>
HDN void lu_solve(const Tile& tl, double (*col)[TS_MAXN], double* rhs, int n) {
  const int L = Tile::LPE;
  double A[TS_MAXN][TS_MAXN];           // A[i][c], every index below is a compile-time constant
  double b[TS_MAXN];
#pragma unroll
  for (int c = 0; c < TS_MAXN; ++c)
#pragma unroll
    for (int i = 0; i < TS_MAXN; ++i) {
      double v = tl.bcast(col[c / L][i], c % L);
      A[i][c] = (i < n && c < n) ? v : ((i == c) ? 1.0 : 0.0);   // identity padding keeps the 8x8 factorisation regular
    }
#pragma unroll
  for (int i = 0; i < TS_MAXN; ++i) b[i] = (i < n) ? rhs[i] : 0.0;
  if (lu_factor_solve<false>(A, b)) {
    // rare: a row exchange is needed -- gather again and run the pivoting elimination
#pragma unroll
    for (int c = 0; c < TS_MAXN; ++c)
#pragma unroll
      for (int i = 0; i < TS_MAXN; ++i) {
        double v = tl.bcast(col[c / L][i], c % L);
        A[i][c] = (i < n && c < n) ? v : ((i == c) ? 1.0 : 0.0);
      }
#pragma unroll
    for (int i = 0; i < TS_MAXN; ++i) b[i] = (i < n) ? rhs[i] : 0.0;
    lu_factor_solve<true>(A, b);
  }
#pragma unroll
  for (int i = 0; i < TS_MAXN; ++i) rhs[i] = b[i];
}

// Row-owner elimination for tiles with one lane per row (LPE >= TS_MAXN): lane i holds row i (a) and
// b_i; pivot rows travel by shuffles, the solution comes out replicated in x.  Same operations on the
// same numbers as lu_factor_solve<false> (no row exchanges), at 1/8 of the per-lane work and without
// the 8x8 register copy.  Returns true (tile-wide) when partial pivoting would have exchanged a row:
// the caller then runs the replicated pivoting solve instead.  (Round-1 / early round-2 path, kept for the A/B builds
// -DTS_NO_LU_PIVOT [-DTS_LU_SMEM=0]: the kernels run lu_rows_solve_pivot.)
template <class Tile>
HD bool lu_rows_solve(const Tile& tl, double* a, double b, double* x) {
  bool exch = false;
#pragma unroll
  for (int j = 0; j < TS_MAXN; ++j) {
    const double pjj = tl.bcast(a[j], j);
    const bool below = tl.lane > j;
    exch = exch || (below && fabs(a[j]) > fabs(pjj));
    const double l = a[j] / pjj;
#pragma unroll
    for (int c = j + 1; c < TS_MAXN; ++c) {
      const double pjc = tl.bcast(a[c], j);
      if (below) a[c] -= l * pjc;
    }
    const double bj = tl.bcast(b, j);
    if (below) b -= l * bj;
  }
#pragma unroll
  for (int k = TS_MAXN - 1; k >= 0; --k) {
    const double xk = tl.bcast(b / a[k], k);
    x[k] = xk;
    if (tl.lane < k) b -= a[k] * xk;
  }
  return tl.ballot(exch) != 0u;
}

// The same elimination with the pivot rows travelling through the tile's shared scratch instead of shuffles (a double
// shuffle is two 32-bit shuffles plus packing: ~5 instructions against one shared-memory load): scr holds two pivot-row
// buffers of TS_MAXN + 1 doubles (row + right-hand side, alternating between steps) and TS_MAXN solution slots.
// Same operations on the same numbers as lu_rows_solve: bit-identical results.
template <class Tile>
HD bool lu_rows_solve_smem(const Tile& tl, double* a, double b, double* x, double* scr) {
  bool exch = false;
  double* xs = scr + 2 * (TS_MAXN + 1);
#pragma unroll
  for (int j = 0; j < TS_MAXN; ++j) {
    double* pr = scr + (j & 1) * (TS_MAXN + 1);
    if (tl.lane == j) {
#pragma unroll
      for (int c = j; c < TS_MAXN; ++c) pr[c] = a[c];
      pr[TS_MAXN] = b;
    }
    tl.tile_sync();
    const double pjj = pr[j];
    const bool below = tl.lane > j;
    exch = exch || (below && fabs(a[j]) > fabs(pjj));
    const double l = a[j] / pjj;
#pragma unroll
    for (int c = j + 1; c < TS_MAXN; ++c) {
      const double pjc = pr[c];
      if (below) a[c] -= l * pjc;
    }
    const double bj = pr[TS_MAXN];
    if (below) b -= l * bj;
  }
#pragma unroll
  for (int k = TS_MAXN - 1; k >= 0; --k) {
    if (tl.lane == k) xs[k] = b / a[k];
    tl.tile_sync();
    const double xk = xs[k];
    x[k] = xk;
    if (tl.lane < k) b -= a[k] * xk;
  }
  return tl.ballot(exch) != 0u;
}

// Row-owner elimination WITH partial pivoting (what the kernels run): the rows stay in their lanes and carry their current
// position; step j takes the row of maximal |a_j| among the positions >= j (first position on ties: Eigen partialPivLu,
// DH/Simulation.cpp:1178, and lu_factor_solve<true>), exchanges the two positions and eliminates as above.  Same
// operations on the same numbers as the replicated pivoting solve.  The exchange-free elimination above hands every
// system that needs a row exchange to the replicated solve -- 16 x 16 in local memory, ~50 K cycles -- and the Newton
// matrices of the struggling steps (hundreds of iterations at the cap, the environments a kernel waits for) need one
// in every iteration: 38 % of the time of the slowest tile of TactileInsertion, 22 % of DClaw (profiles/r02_experiments.md).
// TactilePush needs an exchange in fewer solves, but in a lock-step block one tile in the replicated 8 x 8 solve holds up
// the round of all 28: forward call 75.6 -> 68.0 ms, adjoint call 11.6 -> 8.7 ms at B = 4096, T = 200 (exchange-free
// attempt first, pivoting on demand: 69.2 / 9.6 ms).
// scr: two pivot-row buffers of TS_MAXN + 1, TS_MAXN solution slots, two magnitude buffers of TS_MAXN.
template <class Tile>
HD void lu_rows_solve_pivot(const Tile& tl, double* a, double b, double* x, double* scr) {
  double* xs = scr + 2 * (TS_MAXN + 1);
  double* mag = xs + TS_MAXN;
  const bool row = tl.lane < TS_MAXN;
  int pos = tl.lane;
#pragma unroll
  for (int j = 0; j < TS_MAXN; ++j) {
    double* pr = scr + (j & 1) * (TS_MAXN + 1);
    double* mg = mag + (j & 1) * TS_MAXN;
    if (row && pos >= j) mg[pos] = fabs(a[j]);
    tl.tile_sync();
    int p = j;
    double best = mg[j];
#pragma unroll
    for (int i = j + 1; i < TS_MAXN; ++i) {
      const double v = mg[i];
      if (v > best) { best = v; p = i; }
    }
    if (pos == p) pos = j;
    else if (pos == j) pos = p;
    if (row && pos == j) {
#pragma unroll
      for (int c = j; c < TS_MAXN; ++c) pr[c] = a[c];
      pr[TS_MAXN] = b;
    }
    tl.tile_sync();
    const double pjj = pr[j];
    const bool below = pos > j;
    const double l = a[j] / pjj;
#pragma unroll
    for (int c = j + 1; c < TS_MAXN; ++c) {
      const double pjc = pr[c];
      if (below) a[c] -= l * pjc;
    }
    const double bj = pr[TS_MAXN];
    if (below) b -= l * bj;
  }
#pragma unroll
  for (int k = TS_MAXN - 1; k >= 0; --k) {
    if (row && pos == k) xs[k] = b / a[k];
    tl.tile_sync();
    const double xk = xs[k];
    x[k] = xk;
    if (pos < k) b -= a[k] * xk;
  }
}
#ifndef TS_LU_SMEM
#define TS_LU_SMEM 1
#endif

HD double norm_n(const double* v, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return sqrt(s);
}

// One Dual evaluation per owned column at the point x (an array of the tile state): leaves g in
// ts.g and returns the owned columns of dg/d(seed).  seed: 0 = q1, 1 = q0, 2 = qd0.  The step state
// machine calls it from exactly ONE place, so the residual code exists once per kernel.
template <class Tile, class WK>
HD void eval_columns(const Tile& tl, const SceneView& S, TileState& ts, const double* x, int seed, int mode, WK& W,
                     double (*col)[TS_MAXN], bool idle = false) {
  const int L = Tile::LPE;
  const int n = S.n;
  SeedIn in;
  in.xq = ts.xq; in.xv = ts.xv; in.xl = ts.xl;
  // tangents of (q1, qd1, dl) per unit of the seeded variable:  q1: (1, 1/h, 1)   q0: (0, -1/h, -1)   qd0: (0, 0, -h)
  in.tq = (seed == 0) ? 1.0 : 0.0;
  double stv, sbeta;
  stage_coef(S, mode, stv, sbeta);
  in.tv = (seed == 0) ? stv : ((seed == 1) ? (0.0 - 1.0) / S.h : 0.0);
  in.tl = (seed == 0) ? 1.0 : ((seed == 1) ? -1.0 : -S.h);
  in.q0v = ts.q; in.qd0v = ts.qd;
  in.tq0 = (seed == 1) ? 1.0 : 0.0;
  in.tqd0 = (seed == 2) ? 1.0 : 0.0;
  tl.tile_sync();          // every lane of the tile is done with the previous evaluation and its bookkeeping
  if (!(Tile::COOP && idle)) {
    if (L >= TS_MAXN) {
      // one coordinate per lane (the stage formulas hold fp64 divisions: 8 of them in every lane cost 6 % of the
      // instructions of the step loop); the values are the same, written once instead of LPE times
      const int i = tl.lane;
      if (i < TS_MAXN) {
        const double xi = (i < n) ? x[i] : 0.0;
        double xv = 0.0, xl = 0.0;
        if (i < n) stage_inputs(S, ts, mode, i, xi, xv, xl);
        ts.xq[i] = xi;
        ts.xv[i] = xv;
        ts.xl[i] = xl;
      }
      tl.tile_sync();
    } else {
#pragma unroll
      for (int i = 0; i < TS_MAXN; ++i) {
        const double xi = (i < n) ? x[i] : 0.0;
        double xv = 0.0, xl = 0.0;
        if (i < n) stage_inputs(S, ts, mode, i, xi, xv, xl);
        ts.xq[i] = xi;
        ts.xv[i] = xv;
        ts.xl[i] = xl;
      }
    }
  }
  for (int c = 0; c < TS_NC(L); ++c) {
    in.k = tl.lane + c * L;
    Dual gD[TS_MAXN];
    eval_g(tl, S, in, ts.u, W, gD, sbeta, idle);
    if (Tile::COOP && idle) continue;
#pragma unroll
    for (int i = 0; i < TS_MAXN; ++i) {
      ts.g[i] = (i < n) ? gD[i].v : 0.0;
      col[c][i] = (i < n && in.k < n) ? gD[i].d : 0.0;
    }
  }
}

// Column k of the mass matrix M = J^T Mm J at the poses held in W (values), value arithmetic only:
// psi = S_k on every joint of the subtree of dof k's joint, a_i = I psi_i, projected inward.
template <class WK>
HDN void mass_column(const SceneView& S, const WK& W, const double* qv, int k, double* Mcol) {
  for (int i = 0; i < TS_MAXN; ++i) Mcol[i] = 0.0;
  if (k >= S.n) return;
  // joint and world screw of dof k
  int jk = -1;
  double Sk[6] = {0, 0, 0, 0, 0, 0};
  for (int j = 0; j < S.nj; ++j) {
    const int* ji = S.ib + S.o_joint + j * KJ_ISTRIDE;
    if (k < ji[2] || k >= ji[2] + ji[3]) continue;
    jk = j;
    const double* jd = S.db + S.d_joint + j * KJ_DSTRIDE;
    double R0[9], p0[3];
    wk_ldv(W, j, WK_R0, 9, R0);
    wk_ldv(W, j, WK_P0, 3, p0);
    const int jt = ji[0], loc = k - ji[2];
    if (jt == TS_JT_REVOLUTE) {
      mv3(R0, jd + KJ_AX0, Sk);            // R0 a = Ra a for a rotation about a
      cross3(p0, Sk, Sk + 3);
    } else if (jt == TS_JT_PRISMATIC) mv3(R0, jd + KJ_AX0, Sk + 3);
    else if (jt == TS_JT_PLANAR) mv3(R0, loc == 0 ? jd + KJ_AX0 : jd + KJ_AX1, Sk + 3);
    else if ((KT_FREE3D && (jt == TS_JT_FREE3D_EULER || jt == TS_JT_SPHERICAL_EULER)) ||
             (KT_EXP3D && (jt == TS_JT_FREE3D_EXP || jt == TS_JT_SPHERICAL_EXP))) {
      const int nt = (jt == TS_JT_FREE3D_EULER || jt == TS_JT_FREE3D_EXP) ? 3 : 0;     // translational coordinates first
      const int ro = ji[2] + nt;
      double tax[3][3], rax[3][3];
      if (KT_EXP3D && (jt == TS_JT_FREE3D_EXP || jt == TS_JT_SPHERICAL_EXP)) exp_world_axes(R0, qv[ro], qv[ro + 1], qv[ro + 2], tax, rax);
      else euler_world_axes(R0, qv[ro], qv[ro + 1], qv[ro + 2], tax, rax);
      if (loc < nt) { for (int i = 0; i < 3; ++i) Sk[3 + i] = tax[loc][i]; }
      else {
        for (int i = 0; i < 3; ++i) Sk[i] = rax[loc - nt][i];
        cross3(p0, Sk, Sk + 3);
      }
    } else if (KT_FREE3D && jt == TS_JT_FREE2D) {
      const double s3 = sin(qv[ji[2] + 2]), c3 = cos(qv[ji[2] + 2]);
      if (loc == 0) { const double r0[3] = {c3, -s3, 0.0}; mv3(R0, r0, Sk + 3); }
      else if (loc == 1) { const double r1[3] = {s3, c3, 0.0}; mv3(R0, r1, Sk + 3); }
      else {
        Sk[0] = R0[2]; Sk[1] = R0[5]; Sk[2] = R0[8];
        cross3(p0, Sk, Sk + 3);
      }
    } else { for (int i = 0; i < 3; ++i) Sk[3 + i] = R0[3 * i + loc]; }
  }
  if (jk < 0) return;
  double Wm[TS_MAXJ][6];
  for (int j = 0; j < TS_MAXJ; ++j) for (int i = 0; i < 6; ++i) Wm[j][i] = 0.0;
  for (int j = 0; j < S.nj; ++j) {
    const int* ji = S.ib + S.o_joint + j * KJ_ISTRIDE;
    if (!ji[5] || !((ji[4] >> jk) & 1)) continue;      // massless, or not in the subtree of jk
    const double* jd = S.db + S.d_joint + j * KJ_DSTRIDE;
    double R[9], p[3], ps[6], a[6];
    wk_ldv(W, j, WK_R0, 9, R);
    wk_ldv(W, j, WK_P0, 3, p);
    twist_to_frame(R, p, Sk, ps);
    inertia_mul(jd, ps, a);
    double f[3], t[3], pf[3];
    mv3(R, a + 3, f);
    mv3(R, a, t);
    cross3(p, f, pf);
    for (int i = 0; i < 3; ++i) { Wm[j][i] += t[i] + pf[i]; Wm[j][3 + i] += f[i]; }
  }
  for (int j = S.nj - 1; j >= 0; --j) {
    const int* ji = S.ib + S.o_joint + j * KJ_ISTRIDE;
    const double* jd = S.db + S.d_joint + j * KJ_DSTRIDE;
    const int jt = ji[0], par = ji[1], qo = ji[2];
    double R0[9], p0[3], fj[3];
    wk_ldv(W, j, WK_R0, 9, R0);
    wk_ldv(W, j, WK_P0, 3, p0);
    const double* A = Wm[j];
    mtv3(R0, A + 3, fj);
    if (jt == TS_JT_REVOLUTE) {
      double pf[3], t[3], nj3[3];
      cross3(p0, A + 3, pf);
      for (int i = 0; i < 3; ++i) t[i] = A[i] - pf[i];
      mtv3(R0, t, nj3);
      Mcol[qo] = dot3(jd + KJ_AX0, nj3);
    } else if (jt == TS_JT_PRISMATIC) Mcol[qo] = dot3(jd + KJ_AX0, fj);
    else if (jt == TS_JT_PLANAR) { Mcol[qo] = dot3(jd + KJ_AX0, fj); Mcol[qo + 1] = dot3(jd + KJ_AX1, fj); }
    else if (jt == TS_JT_TRANSLATIONAL) { Mcol[qo] = fj[0]; Mcol[qo + 1] = fj[1]; Mcol[qo + 2] = fj[2]; }
    else if ((KT_FREE3D && (jt == TS_JT_FREE3D_EULER || jt == TS_JT_SPHERICAL_EULER)) ||
             (KT_EXP3D && (jt == TS_JT_FREE3D_EXP || jt == TS_JT_SPHERICAL_EXP))) {
      const int nt = (jt == TS_JT_FREE3D_EULER || jt == TS_JT_FREE3D_EXP) ? 3 : 0;
      const int ro = qo + nt;
      double tax[3][3], rax[3][3], pf[3], t[3];
      if (KT_EXP3D && (jt == TS_JT_FREE3D_EXP || jt == TS_JT_SPHERICAL_EXP)) exp_world_axes(R0, qv[ro], qv[ro + 1], qv[ro + 2], tax, rax);
      else euler_world_axes(R0, qv[ro], qv[ro + 1], qv[ro + 2], tax, rax);
      cross3(p0, A + 3, pf);
      for (int i = 0; i < 3; ++i) t[i] = A[i] - pf[i];
      for (int i = 0; i < 3; ++i) { if (nt) Mcol[qo + i] = dot3(tax[i], A + 3); Mcol[ro + i] = dot3(rax[i], t); }
    } else if (KT_FREE3D && jt == TS_JT_FREE2D) {
      const double s3 = sin(qv[qo + 2]), c3 = cos(qv[qo + 2]);
      const double r0[3] = {c3, -s3, 0.0}, r1[3] = {s3, c3, 0.0};
      double ax0[3], ax1[3], pf[3], t[3];
      mv3(R0, r0, ax0);
      mv3(R0, r1, ax1);
      cross3(p0, A + 3, pf);
      for (int i = 0; i < 3; ++i) t[i] = A[i] - pf[i];
      Mcol[qo] = dot3(ax0, A + 3);
      Mcol[qo + 1] = dot3(ax1, A + 3);
      Mcol[qo + 2] = R0[2] * t[0] + R0[5] * t[1] + R0[8] * t[2];
    }
    if (par >= 0) for (int i = 0; i < 6; ++i) Wm[par][i] += A[i];
  }
}

// ||g(x + alpha dx)|| by a value-only evaluation (plain per-lane work space) that a SUB-TILE of the tile's lanes runs
// for itself: the sub-tiles of a tile evaluate different step lengths of a struggling line search at the same time.
// TS_LS_SUB = 1 (default): every lane alone (one-lane policy).  TS_LS_SUB = 4: four lanes per step length repeat the same
// value arithmetic and share the contact DETECTION of the evaluation (dealt to the lanes, gathered by ballot) -- the
// scenes of the 16-lane variants test hundreds of sampled points per force (DClaw: 3 x 217).  MEASURED SLOWER on B200
// (forward call: DClaw B=2048 T=200 691 -> 798 ms, TactileInsertion B=1024 T=45 675 -> 904 ms, StableGrasp 719 -> 981 ms):
// a struggling search accepts a step length many halvings down, and four step lengths per pass instead of sixteen
// means more passes.
template <class SubTile>
HDN double trial_norm(const SubTile& solo, const SceneView& S, const TileState& ts, double alpha, int mode) {
  const int n = S.n;
  double xq[TS_MAXN], xv[TS_MAXN], xl[TS_MAXN], g[TS_MAXN];
  for (int i = 0; i < TS_MAXN; ++i) {
    const double xi = (i < n) ? ts.x[i] + alpha * ts.dx[i] : 0.0;
    xq[i] = xi;
    xv[i] = 0.0;
    xl[i] = 0.0;
    if (i < n) stage_inputs(S, ts, mode, i, xi, xv[i], xl[i]);
    g[i] = 0.0;
  }
  ArrIn<double> in;
  in.q_ = xq; in.qd_ = xv; in.dl_ = xl; in.q0_ = ts.q; in.qd0_ = ts.qd;
  WorkRec<double> Wv;
  double stv, sbeta;
  stage_coef(S, mode, stv, sbeta);
  eval_g(solo, S, in, ts.u, Wv, g, sbeta);
  return norm_n(g, n);
}

// status word per env-step: newton iterations | line-search evaluations << 8 | flags << 16
#define TS_STAT_NOT_CONVERGED (1 << 16)
#define TS_STAT_NAN (1 << 17)

// State of one implicit step in flight (DH/Simulation.cpp:1325-1351 with the Newton of :1150-1225).
// phase 0: (g, H) at x, then iterate | 1: line-search trial at xn | 2: (H) at the final x | 3: G0 at the final x
struct StepVars {
  double alpha, gnorm;
  int phase, fail_strike, iters, ls, trial;
  bool converged, batch_ls;
  int cap_newton;
  int mode;                 // TS_ST_*: implicit stage in flight
  bool defer_g0;            // the G0 / G1 blocks of the tape are written by the pass of their own (env_tape), not by phase 3
  bool ls_pending, ls_fresh; // a request to the line-search service is out (step_post resumes behind it); `fresh` across it
};

HD void step_begin(const SceneView& S, StepVars& v, TileState& ts) {
  // initial guesses: q0 + h qd0 (BDF1, with h := alpha h in the first SDIRK2 stage), q1 + h qd1 (BDF2),
  // q_alpha + (1 - alpha) h qd_alpha (second SDIRK2 stage)
  for (int i = 0; i < TS_MAXN; ++i) {
    double x0 = (i < S.n) ? ts.q[i] + S.h * ts.qd[i] : 0.0;
#if KT_MULTISTEP
    if (i < S.n && v.mode == TS_ST_SDIRK_A) x0 = ts.q[i] + (TS_SDIRK_ALPHA * S.h) * ts.qd[i];
    if (i < S.n && v.mode == TS_ST_SDIRK_B) x0 = ts.pq[i] + (1 - TS_SDIRK_ALPHA) * S.h * ts.pqd[i];
#endif
    ts.x[i] = x0; ts.dx[i] = 0.0; ts.xn[i] = 0.0;
  }
  v.phase = 0; v.fail_strike = 0; v.iters = 0; v.ls = 0; v.trial = 0;
  v.alpha = 1.0; v.gnorm = 0.0; v.converged = false;
  v.ls_pending = false; v.ls_fresh = false;
}

// One round = ONE residual evaluation (with Jacobian columns) plus the bookkeeping that follows it.
// Returns true when the step is complete: v.x is the new q, the tape (grad mode: H, G0, G1 as
// [3][n][n] row-major, columns written by their owner lanes) is written and the work space holds the
// kinematics (values) of the new state.
// Every line-search trial carries its Jacobian, so an accepted trial is at once the next iterate's
// (g, H) and -- when converged -- the tape's H.
// step_eval is the evaluation (every tile of a warp runs it together, also tiles whose step is already
// complete: the residual code votes and shuffles across the warp); step_post is the bookkeeping.
template <class Tile, class WK>
HD void step_eval(const Tile& tl, const SceneView& S, const StepVars& v, WK& WD, double (*cole)[TS_MAXN], bool idle = false) {
  TileState& ts = WD.state();
  eval_columns(tl, S, ts, (v.phase == 1) ? ts.xn : ts.x, v.phase == 3 ? 1 : 0, v.mode, WD, cole, idle);
}

template <class Tile, class WK>
HD bool step_post(const Tile& tl, const SceneView& S, StepVars& v, double* tape, WK& WD, double (*cole)[TS_MAXN]) {
  const int L = Tile::LPE;
  const int n = S.n;
  TileState& ts = WD.state();
  const double* ge = ts.g;
  const int ref_newton = 20 * n > S.max_iter ? 20 * n : S.max_iter;      // DH/Simulation.cpp:1155
  const int max_newton = (v.cap_newton > 0 && v.cap_newton < ref_newton) ? v.cap_newton : ref_newton;
  if (v.phase == 3) {
    // G0 = dg/dq0 from this evaluation, G1 = dg/dqdot0 = -h M from mass-matrix columns
    for (int c = 0; c < TS_NC(L); ++c) {
      const int k = tl.lane + c * L;
      if (k < n) {
        double Mc[TS_MAXN];
        mass_column(S, WD, ts.xq, k, Mc);
        for (int i = 0; i < n; ++i) {
          st_stream(tape + n * n + i * n + k, cole[c][i]);
          tape[2 * n * n + i * n + k] = -S.h * Mc[i];      // (read-modify-written below by position motors)
        }
      }
    }
    // d f_r / d u per control (the gain the reverse sweep needs: DH/Actuator/ActuatorMotor.cpp:48-62), and the
    // velocity feedback of position-controlled motors, which enters G1 = dg/dqdot0 = -h M + h^2 diag(D)
    tl.tile_sync();
    if (tl.lane == 0) {
      for (int ai = 0; ai < S.nact; ++ai) {
        const int* r = S.ib + S.o_act + ai * KA_ISTRIDE;
        const double* cdat = S.db + S.d_act + ai * KA_DSTRIDE;
        const int qo = S.ib[S.o_joint + r[0] * KJ_ISTRIDE + 2];
        for (int i = 0; i < r[3]; ++i) {
          const double uu = ts.u[r[2] + i];
          double gain;
          if (KT_POS_MOTOR && r[1] == TS_ACT_POS) {
            const double f = cdat[6 + i] * (uu - ts.q[qo + i]) + cdat[9 + i] * (-ts.qd[qo + i]);
            const bool open = f >= cdat[i] && f <= cdat[3 + i];
            gain = open ? cdat[6 + i] : 0.0;
            if (open) tape[2 * n * n + (qo + i) * n + (qo + i)] += S.h * S.h * cdat[9 + i];
          } else {
            gain = (uu >= -1.0 && uu <= 1.0) ? (cdat[3 + i] - cdat[i]) / 2.0 : 0.0;
          }
          tape[3 * n * n + r[2] + i] = gain;
        }
      }
    }
    return true;
  }
  bool finished = false, fresh = (v.phase == 2);   // fresh: (ge, cole) belong to the final x
  // eager: the Newton iteration below has just set up a search and this step is struggling (TS_LS_EAGER iterations or
  // more): the search starts with the batched evaluation of the step lengths instead of two sequential trials -- the
  // same accepted step length and evaluation counts, one or two rounds fewer per iteration of the steps that run into
  // hundreds of iterations (and decide the duration of a kernel whose blocks end with their slowest environment).
  // Threshold, measured on B200 with the pivoting row-owner LU in place (forward call): TactilePush (variant 8) 3 / 4 / 6 /
  // 8 / 12 / 16 / 24 / 40 iterations: 78.8 / 73.0 / 67.9 / 65.7 / 65.5 / 65.9 / 65.9 / 65.5 ms; the 16-dof scenes are
  // best at 6 (3 / 4 / 6 / 10: DClaw 557 / 543 / 535 / 548 ms, TactileInsertion 385 / 389 / 394 / 400 ms, StableGrasp 859 /
  // 856 / 862 / 875 ms).
#ifndef TS_LS_EAGER
#define TS_LS_EAGER (TS_MAXN > 8 ? 6 : 12)
#endif
  bool eager = false;
#if TS_LS_SERVICE
  // second call of a round: the answers of the line-search service are in the tile state, the search resumes at its
  // batched evaluation ((ge, cole) are still those of the first call: same round, same scope)
  const bool resume = Tile::LS_SERVICE && v.ls_pending;
  if (resume) { eager = true; fresh = v.ls_fresh; v.ls_pending = false; }
#endif
  for (;;) {
  if (v.phase == 1) {
    double gnn = 0.0;
    bool accepted = false;
    if (!eager) {
      ++v.ls;
      gnn = norm_n(ge, n);
      accepted = gnn < v.gnorm;
    }
    if (accepted) {                                // trial accepted: it is the next iterate
      for (int i = 0; i < n; ++i) ts.x[i] = ts.xn[i];
      v.fail_strike = 0;
      fresh = true;
      if (gnn < S.tol) { v.converged = true; finished = true; }
    } else {
      if (!eager) {
        ++v.trial;
        v.alpha *= 0.5;
        // (sequential trials before the batch, measured on B200 with the pivoting row-owner LU in place, forward call:
        // TactilePush 1 / 2 / 3 / 4 / 5 / 8 / 20 trials: 70.2 / 65.1 / 63.0 / 62.4 / 62.3 / 62.6 / 62.3 ms -- a value-only
        // pass in one tile holds up the round of its lock-step block; the 16-dof scenes stay at 2: DClaw 542 / 536 / 536 ms,
        // TactileInsertion 387 / 392 / 394 ms, StableGrasp 861 / 863 / 868 ms for 1 / 2 / 3)
#ifndef TS_LS_BATCH_AFTER
#define TS_LS_BATCH_AFTER (TS_MAXN > 8 ? 2 : 4)
#endif
        if (v.trial < S.max_ls && !(L >= TS_MAXN && v.batch_ls && v.trial >= TS_LS_BATCH_AFTER)) {
          for (int i = 0; i < n; ++i) ts.xn[i] = ts.x[i] + v.alpha * ts.dx[i];
          return false;
        }
      }
      if (v.trial < S.max_ls) {
        // TS_LS_BATCH_AFTER trials already failed: a struggling line search (up to max_ls = 20 trials per iteration,
        // DH/Simulation.cpp:1186-1200) would serialise the whole block behind this tile.  The lanes of the
        // tile evaluate the next LPE step lengths at once (value only); the first one, in the reference's
        // order, that reduces ||g|| is then evaluated with its Jacobian by the normal path, which also
        // re-checks the acceptance.  Same accepted step length as the sequential search.
        bool found = false;
        TS_CPT0();
#ifndef TS_LS_SUB
#define TS_LS_SUB 1
#endif
        const int NBMAX = L / TS_LS_SUB > 0 ? L / TS_LS_SUB : 1;     // step lengths evaluated at once
#if TS_LS_SERVICE
        if (Tile::LS_SERVICE && (resume || S.max_ls - v.trial <= TS_LS_CAP)) {
          const int nal = S.max_ls - v.trial;
          if (!resume) {
            // hand the step lengths to the block (env_forward runs ls_service behind this call, then calls again)
            for (int i = tl.lane; i < nal; i += L) ts.lsn[i] = HUGE_VAL;
            if (tl.lane == 0) {
              ts.ls_alpha0 = v.alpha; ts.ls_gnorm = v.gnorm; ts.ls_db = S.db;
              ts.ls_n = nal; ts.ls_mode = v.mode; ts.ls_found = 0; ts.ls_req = 1;
            }
            v.ls_pending = true; v.ls_fresh = fresh;
            return false;
          }
          int k = -1;
          for (int i = nal - 1; i >= 0; --i) if (ts.lsn[i] < v.gnorm) k = i;      // first accepted, in the reference's order
          const int adv = k >= 0 ? k : nal;
          v.trial += adv;
          v.ls += adv;
          for (int i = 0; i < adv; ++i) v.alpha *= 0.5;
          found = k >= 0;
          if (!found) gnn = ts.lsn[nal - 1];
          tl.tile_sync();
          if (tl.lane == 0) ts.ls_req = 0;
        } else
#endif
        while (v.trial < S.max_ls) {
          const int left = S.max_ls - v.trial;
          const int nb = left < NBMAX ? left : NBMAX;
          const int mine = tl.lane / TS_LS_SUB;      // the step length this lane works on
          double my_alpha = v.alpha;
          for (int i = 0; i < mine; ++i) my_alpha *= 0.5;
          double my_norm = 0.0;
          if (mine < nb) my_norm = trial_norm(tl.template sub<TS_LS_SUB>(), S, ts, my_alpha, v.mode);
          const unsigned okbits = tl.ballot(mine < nb && my_norm < v.gnorm);
          if (okbits) {
            const int k = ts_ffs(okbits) / TS_LS_SUB;
            v.trial += k;
            v.ls += k;                           // the rejected trials the sequential search would have evaluated
            for (int i = 0; i < k; ++i) v.alpha *= 0.5;
            found = true;
            break;
          }
          gnn = tl.bcast(my_norm, (nb - 1) * TS_LS_SUB);   // ||g|| of the last trial evaluated (used by the exhausted path)
          v.trial += nb;
          v.ls += nb;
          for (int i = 0; i < nb; ++i) v.alpha *= 0.5;
        }
        TS_CPT(tl, 15);
        if (found) {
          tl.tile_sync();
          for (int i = 0; i < n; ++i) ts.xn[i] = ts.x[i] + v.alpha * ts.dx[i];
          return false;
        }
      }
      // line search exhausted (DH/Simulation.cpp:1201-1214): strike, else step with the last alpha
      ++v.fail_strike;
      if (v.fail_strike >= 10) finished = true;
      else {
        double xs[TS_MAXN];                    // read-modify-write of shared tile state: reads, sync, writes
        for (int i = 0; i < TS_MAXN; ++i) xs[i] = (i < n) ? ts.x[i] + v.alpha * ts.dx[i] : 0.0;
        tl.tile_sync();
        for (int i = 0; i < TS_MAXN; ++i) ts.x[i] = xs[i];
        fresh = false;                         // x moved: (ge, cole) are no longer those of x
        if (gnn < S.tol) { v.converged = true; finished = true; }
        else if (v.iters >= max_newton) finished = true;
        else { v.phase = 0; return false; }
      }
    }
    if (!finished && v.iters >= max_newton) finished = true;
  }
  if (finished || v.phase == 2) {
    if (!fresh) { v.phase = 2; return false; }     // (H) and the work space must be at the final x
    if (!tape) return true;
    for (int c = 0; c < TS_NC(L); ++c) {
      const int k = tl.lane + c * L;
      if (k < n) for (int i = 0; i < n; ++i) st_stream(tape + i * n + k, cole[c][i]);
    }
    if (v.defer_g0) return true;
    v.phase = 3;
    return false;
  }
  // Newton iteration from (x, ge, cole): dx = -H^-1 g, then line search from alpha = 1
  ++v.iters;
  v.gnorm = norm_n(ge, n);
  double dx[TS_MAXN];
  bool solved = false;
  if (L >= TS_MAXN) {
    // lane k holds column k of H: transpose through the tile's scratch so that lane i holds row i
    double* Hs = WD.scratch();
    double a[TS_MAXN];
    tl.tile_sync();
    if (tl.lane < TS_MAXN)
      for (int i = 0; i < TS_MAXN; ++i) Hs[i * TS_MAXN + tl.lane] = (i < n && tl.lane < n) ? cole[0][i] : ((i == tl.lane) ? 1.0 : 0.0);
    tl.tile_sync();
    for (int c = 0; c < TS_MAXN; ++c) a[c] = (tl.lane < TS_MAXN) ? Hs[tl.lane * TS_MAXN + c] : 0.0;
    double bsel = 0.0;
    for (int i = 0; i < TS_MAXN; ++i) if (i == tl.lane && i < n) bsel = -ge[i];
#if TS_LU_SMEM && !defined(TS_NO_LU_PIVOT)
    tl.tile_sync();                      // every lane has its row: the scratch is free for the pivot rows
    lu_rows_solve_pivot(tl, a, bsel, dx, Hs);
    solved = true;
#elif TS_LU_SMEM
    tl.tile_sync();                      // every lane has its row: the scratch is free for the pivot rows
    solved = !lu_rows_solve_smem(tl, a, bsel, dx, Hs);
#else
    solved = !lu_rows_solve(tl, a, bsel, dx);
#endif
    tl.tile_sync();
  }
  if (!solved) {
    for (int i = 0; i < TS_MAXN; ++i) dx[i] = (i < n) ? -ge[i] : 0.0;
    lu_solve(tl, cole, dx, n);
  }
  v.alpha = 1.0;
  v.trial = 0;
  for (int i = 0; i < TS_MAXN; ++i) { ts.dx[i] = dx[i]; ts.xn[i] = (i < n) ? ts.x[i] + dx[i] : 0.0; }
  v.phase = 1;
  if (!(TS_LS_EAGER > 0 && L >= TS_MAXN && v.batch_ls && v.iters >= TS_LS_EAGER)) return false;
  tl.tile_sync();                      // dx is in place for the value-only trials of every lane
  eager = true;
  }   // for (;;): second pass = the search of a struggling step, batched from its first trial
}


// ---- block-wide line-search service (TS_LS_SERVICE, the kernels whose tile policy sets LS_SERVICE)
// A struggling line search asks for ||g|| at up to max_ls = 20 step lengths per Newton iteration, hundreds of iterations
// per step, and the kernels of the 16-dof scenes end with ONE such environment (profiles/r02_experiments.md).  The
// in-tile batch (step_post) gives every step length to one lane of the searching tile: sixteen neighbouring lanes of one
// warp, each with its own contact set, walking through the value-only evaluation under divergence.  Here the requests
// of a round are served by the block instead: item (step length i, request j) = number k = i R + j (R requests, lower
// step-length indices first) goes to lane k / NW of warp k % NW (NW warps) -- still one lane per step length, with its
// joint records in local memory (16 to 20 evaluating lanes: the footprint that fits L1), but two or three of them per
// warp instead of sixteen.  The norms come back through the requester's tile state; step_post picks the first
// accepted one in the reference's order, so the accepted step length and the evaluation counts are those of the
// sequential search (DH/Simulation.cpp:1186-1200).  Forward call, B200: DClaw 534 -> 500 ms, TactileInsertion 406 ->
// 371 ms, StableGrasp 866 -> 805 ms.  (Sub-tiles of 8 lanes sharing the contact detection of a step length, with
// their joint records in local or in shared memory, lost: tools/experiments/line_search_service.patch.)
template <bool ON> struct LsTag {};
template <class Tile>
HD bool ls_service_round(const Tile&, const SceneView&, bool, LsTag<false>) { return false; }
template <class Tile>
HDN bool ls_service_round(const Tile& tl, const SceneView& S, bool pending, LsTag<true>) {
#if TS_LS_SERVICE && defined(__CUDACC__)          // (device tiles only: the host harness never instantiates it)
  if (!tl.cta_or_unaligned(pending)) return false;
  const int NT = Tile::NTILES, NW = TS_BLOCK / 32;
  unsigned reqmask = 0, found = 0;
  int R = 0, maxn = 0;
  for (int r = 0; r < NT; ++r) {
    const TileState* rts = tl.peer_state(r);
    if (rts->ls_req) { reqmask |= 1u << r; ++R; maxn = rts->ls_n > maxn ? rts->ls_n : maxn; }
  }
  const int items = maxn * R;
  // item k goes to lane k / NW of warp k % NW: the step lengths of a search are evaluated by ONE lane each, as in the
  // in-tile batch, but spread over the warps of the block instead of sixteen neighbouring lanes of one warp
  const int mine = (int)(threadIdx.x & 31) * NW + (int)(threadIdx.x >> 5);
  for (int base = 0; base < items; base += TS_BLOCK) {
    const int k = base + mine;
    if (k < items) {
      const int i = k / R;
      int j = k % R, r = 0;
      for (unsigned m = reqmask;; m &= m - 1, --j) if (j == 0) { r = ts_ffs(m); break; }
      TileState* rts = tl.peer_state(r);
      if (i < rts->ls_n && !((found >> r) & 1u)) {
        double al = rts->ls_alpha0;
        for (int q = 0; q < i; ++q) al *= 0.5;
        SceneView Sl = S;
        Sl.db = rts->ls_db;
        const double nrm = trial_norm(HostTile(), Sl, *rts, al, rts->ls_mode);
        rts->lsn[i] = nrm;
        if (nrm < rts->ls_gnorm) rts->ls_found = 1;
      }
    }
    tl.cta_sync_unaligned();           // the answers of this pass are in place
    for (int r = 0; r < NT; ++r) if (((reqmask >> r) & 1u) && tl.peer_state(r)->ls_found) found |= 1u << r;
    if (found == reqmask) break;       // (block-uniform) every search has its step length
    if (base + TS_BLOCK < items) tl.cta_sync_unaligned();   // flags read before the next pass writes them
  }
  return true;
#else
  return false;
#endif
}


// ------------------------------------------------------------------ readouts at a state
// end-effector positions (DH/EndEffector/EndEffector.cpp:31-36)
template <class WK, class T>
HD void variable_of(const SceneView& S, const WK& W, int e, T* out) {
  const int j = S.ib[S.o_ee + e * KE_ISTRIDE];
  const double* pos = S.db + S.d_ee + e * KE_DSTRIDE;
  if (j < 0) { for (int i = 0; i < 3; ++i) out[i] = pos[i]; return; }
  T R0[9];
  wk_ld(W, j, WK_R0, 9, R0);
  mv3(R0, pos, out);
  for (int i = 0; i < 3; ++i) out[i] = out[i] + W.get(j, WK_P0 + i);
}

template <class WK>
HDN void sensor_frames(const SceneView& S, const WK& W, const int* sr, const double* sd, Frames& F) {
  const int nc = sr[3];
  body_frame_v(S, W, sr[0], F.R[0], F.p[0], F.ph[0]);
  F.near[0] = true;
  for (int c = 0; c < nc; ++c) {
    const int b2 = sr[4 + c];
    const int sh2 = S.ib[S.o_body + b2 * KB_ISTRIDE + 1];
    const bool cyl = (KT_CYLINDER && sh2 == TS_SH_CYLINDER) || (KT_SPHERE && (sh2 == TS_SH_SPHERE || sh2 == TS_SH_CAPSULE));
    body_frame_v(S, W, b2, F.R[1 + c], F.p[1 + c], F.ph[1 + c]);
    const double rr = sd[KS_RMARK] + S.db[S.d_body + b2 * KB_DSTRIDE + KB_RBOUND] + TS_CULL_MARGIN;
    const double dx = F.p[0][0] - F.p[1 + c][0], dy = F.p[0][1] - F.p[1 + c][1], dz = F.p[0][2] - F.p[1 + c][2];
    F.near[1 + c] = !(dx * dx + dy * dy + dz * dz > rr * rr);
    if (F.near[1 + c] && !cyl) {
      rel_frame(F.R[0], F.p[0], F.R[1 + c], F.p[1 + c], F.R21[c], F.r21[c]);
      if (bbox_outside_box(F.R21[c], F.r21[c], sd + KS_BBOX, sd + KS_BBOX + 3, S.db + S.d_body + b2 * KB_DSTRIDE + KB_HALF))
        F.near[1 + c] = false;
    }
  }
}

// per-marker intermediate of the tactile force, shared by the value pass and its adjoint
struct MarkerHit {
  int cand;           // index of the contacted candidate (last candidate with d < 0), -1 if none
  bool cyl;           // the candidate is a cylinder or a sphere (normal depends on the point), else a cuboid (face normal)
  bool sph;           // ... a sphere (or the cap of a capsule): the normal has an axial component
  double vz;          // axial component of the radial vector
  double e[3];        // contact normal in the candidate's frame
  double x[3], u[3], d, ddot, tb[3], s, tn, rad;
  bool dynamic;
};

// Evaluate one marker of sensor record (sr, sd).  DH/Sensor/TactileSensor.cpp:29-87 with the SDFs of
// DH/Body/BodyCuboid.cpp:135-184 and DH/Body/BodyCylinder.cpp:88-139.
HD void marker_force(const SceneView& S, const Frames& F, const int* sr, const double* sd, const double* xi1,
                     MarkerHit& H, double* F1 /* force in the pad frame */) {
  const int nc = sr[3];
  const double kn = sd[0], kt = sd[1], mu = sd[2], damp = sd[3];
  const double* R1 = F.R[0]; const double* p1 = F.p[0]; const double* ph1 = F.ph[0];
  H.cand = -1;
  H.cyl = false;
  H.sph = false;
  for (int c = 0; c < nc; ++c) {
    if (!F.near[1 + c]) continue;
    const double* hs = S.db + S.d_body + sr[4 + c] * KB_DSTRIDE + KB_HALF;
    double xw[3], y[3], x[3];
    if (KT_SPHERE && S.ib[S.o_body + sr[4 + c] * KB_ISTRIDE + 1] == TS_SH_CAPSULE) {
      // TactileSensor.cpp:44-47 with BodyCapsule::distance; between the caps the cylinder formula, beyond them the
      // sphere formula about the end of the axis
      mv3(R1, xi1, xw);
      for (int i = 0; i < 3; ++i) xw[i] = xw[i] + p1[i];
      if (!capsule_inside_world(F.R[1 + c], F.p[1 + c], xw, hs)) continue;
      for (int i = 0; i < 3; ++i) y[i] = xw[i] - F.p[1 + c][i];
      mtv3(F.R[1 + c], y, x);
      H.cand = c; H.cyl = true; H.x[0] = x[0]; H.x[1] = x[1]; H.x[2] = x[2];
      H.sph = x[2] < -hs[1] || x[2] > hs[1];
      H.vz = x[2] < -hs[1] ? x[2] + hs[1] : (x[2] > hs[1] ? x[2] - hs[1] : 0.0);
      continue;
    }
    if (KT_SPHERE && S.ib[S.o_body + sr[4 + c] * KB_ISTRIDE + 1] == TS_SH_SPHERE) {
      // TactileSensor.cpp:44-47 with BodySphere::distance; the force is evaluated at x = R2^T (xw - p2)
      mv3(R1, xi1, xw);
      for (int i = 0; i < 3; ++i) xw[i] = xw[i] + p1[i];
      if (!sphere_inside_world(F.p[1 + c], xw, hs[0])) continue;
      for (int i = 0; i < 3; ++i) y[i] = xw[i] - F.p[1 + c][i];
      mtv3(F.R[1 + c], y, x);
      H.cand = c; H.cyl = true; H.sph = true; H.x[0] = x[0]; H.x[1] = x[1]; H.x[2] = x[2]; H.vz = x[2];
      continue;
    }
    if (KT_CYLINDER && S.ib[S.o_body + sr[4 + c] * KB_ISTRIDE + 1] == TS_SH_CYLINDER) {
      mv3(R1, xi1, xw);
      for (int i = 0; i < 3; ++i) xw[i] = xw[i] + p1[i];
      if (!cylinder_inside_world(F.R[1 + c], F.p[1 + c], xw, hs)) continue;
      // the force is evaluated at x = R2^T (xw - p2)  (BodyCylinder.cpp:118)
      for (int i = 0; i < 3; ++i) y[i] = xw[i] - F.p[1 + c][i];
      mtv3(F.R[1 + c], y, x);
      H.cand = c; H.cyl = true; H.sph = false; H.x[0] = x[0]; H.x[1] = x[1]; H.x[2] = x[2];
      continue;
    }
    if (cuboid_classify(F.R21[c], F.r21[c], xi1, hs) < 0) continue;   // surely outside
    // reference evaluation order (TactileSensor.cpp:44-47); also yields the exact box-frame point
    mv3(R1, xi1, xw);
    for (int i = 0; i < 3; ++i) y[i] = (xw[i] + p1[i]) - F.p[1 + c][i];
    mtv3(F.R[1 + c], y, x);
    if (cuboid_inside(x, hs)) { H.cand = c; H.cyl = false; H.sph = false; H.x[0] = x[0]; H.x[1] = x[1]; H.x[2] = x[2]; }
  }
  F1[0] = F1[1] = F1[2] = 0.0;
  if (H.cand < 0) return;
  const double* hs = S.db + S.d_body + sr[4 + H.cand] * KB_DSTRIDE + KB_HALF;
  const double* R2 = F.R[1 + H.cand]; const double* ph2 = F.ph[1 + H.cand];
  if (KT_SPHERE && H.sph) {
    H.rad = sqrt(H.x[0] * H.x[0] + H.x[1] * H.x[1] + H.vz * H.vz);
    H.d = H.rad - hs[0];
    H.e[0] = H.x[0] / H.rad; H.e[1] = H.x[1] / H.rad; H.e[2] = H.vz / H.rad;
  } else if (KT_CYLINDER && H.cyl) {
    H.rad = sqrt(H.x[0] * H.x[0] + H.x[1] * H.x[1]);
    H.d = H.rad - hs[0];
    H.e[0] = H.x[0] / H.rad; H.e[1] = H.x[1] / H.rad; H.e[2] = 0.0;
  } else {
    int ax; double sg;
    H.d = cuboid_face(H.x, hs, ax, sg);
    H.e[0] = H.e[1] = H.e[2] = 0.0;
    H.e[ax] = sg;
    H.rad = 0.0;
  }
  double v1[3], xwd[3], t3[3];
  cross3(ph1, xi1, v1);
  for (int i = 0; i < 3; ++i) v1[i] += ph1[3 + i];
  mv3(R1, v1, xwd);
  mtv3(R2, xwd, H.u);
  cross3(ph2, H.x, t3);
  for (int i = 0; i < 3; ++i) H.u[i] = H.u[i] - t3[i] - ph2[3 + i];
  H.ddot = dot3(H.e, H.u);
  cross3(ph2, H.e, t3);
  for (int i = 0; i < 3; ++i) H.tb[i] = H.u[i] + H.d * t3[i];
  const double et = dot3(H.e, H.tb);
  for (int i = 0; i < 3; ++i) H.tb[i] -= H.e[i] * et;
  H.s = kn * H.d - damp * H.ddot * H.d;
  double Fb[3];
  for (int i = 0; i < 3; ++i) Fb[i] = -(H.s * H.e[i]);
  H.dynamic = false;
  H.tn = sqrt(H.tb[0] * H.tb[0] + H.tb[1] * H.tb[1] + H.tb[2] * H.tb[2]);
  if (mu > TS_EPS) {
    const double fcn = fabs(H.s);   // |fc| of the 3-vector -s R1^T n
    if (mu * fcn >= kt * H.tn - TS_EPS) {
      for (int i = 0; i < 3; ++i) Fb[i] -= kt * H.tb[i];
    } else {
      H.dynamic = true;
      const double sc = mu * fcn / H.tn;
      for (int i = 0; i < 3; ++i) Fb[i] -= sc * H.tb[i];
    }
  }
  double Fw[3];
  mv3(R2, Fb, Fw);
  mtv3(R1, Fw, F1);
}

// tactile values of this env, markers strided over the tile's lanes.
// out: [M][3] = (shear . axis0, shear . axis1, normal)      (TactileSensor.cpp:74-81)
// prezeroed: `out` is already all zeros (tac_kernel after a memset): sensors that nothing can reach are skipped.
template <class Tile, class WK>
HDN void tactile_values(const Tile& tl, const SceneView& S, WK& W, double* out, int* body_out, bool prezeroed = false) {
  for (int si = 0; si < S.nsens; ++si) {
    const int* sr = S.ib + S.o_sensor + si * KS_ISTRIDE;
    const double* sd = S.db + S.d_sensor + si * KS_DSTRIDE;
    const int mo = sr[1], mc = sr[2];
    Frames& F = W.frames();
    tl.tile_sync();
    sensor_frames(S, W, sr, sd, F);
    tl.tile_sync();
    bool anynear = false;
    for (int c = 0; c < sr[3]; ++c) anynear = anynear || F.near[1 + c];
    if (!anynear) {
      // no candidate body can reach the pad: zero field, written with unit stride across the lanes
      if (!prezeroed) for (int i = tl.lane; i < 3 * mc; i += Tile::LPE) st_stream(out + 3 * mo + i, (i % 3 == 2) ? -0.0 : 0.0);
      if (body_out) for (int m = tl.lane; m < mc; m += Tile::LPE) st_stream(body_out + mo + m, -1);
      continue;
    }
    for (int m = tl.lane; m < mc; m += Tile::LPE) {
      double* o = out + 3 * (mo + m);
      const double* mk = S.mk + KM_STRIDE * (mo + m);     // position, axis0, axis1, normal
      MarkerHit H;
      double F1[3];
      marker_force(S, F, sr, sd, mk, H, F1);
      st_stream(o + 0, dot3(F1, mk + 3));
      st_stream(o + 1, dot3(F1, mk + 6));
      st_stream(o + 2, -dot3(F1, mk + 9));
      if (body_out) st_stream(body_out + mo + m, H.cand < 0 ? -1 : sr[4 + H.cand]);
    }
  }
}

// cotangent accumulators of the tactile adjoint for one candidate body:
// R21 = R2^T R1 (9), r = R2^T (p1 - p2) (3), phi1 (6), phi2 (6)
struct TacAcc {
  double v[24];
};

// Reverse-mode through marker_force for every marker of sensor si in this lane's stride.  Returns
// (tile-wide) whether any marker is in contact; acc is only meaningful then.
template <class Tile, class WK>
HDN bool tactile_vjp(const Tile& tl, const SceneView& S, WK& W, int si, const double* wbar, TacAcc* acc) {
  for (int c = 0; c < TS_MAXCAND; ++c) for (int i = 0; i < 24; ++i) acc[c].v[i] = 0.0;
  double hits = 0.0;
  {
    const int* sr = S.ib + S.o_sensor + si * KS_ISTRIDE;
    const double* sd = S.db + S.d_sensor + si * KS_DSTRIDE;
    const int mo = sr[1], mc = sr[2];
    const double kn = sd[0], kt = sd[1], mu = sd[2], damp = sd[3];
    Frames& F = W.frames();
    tl.tile_sync();
    sensor_frames(S, W, sr, sd, F);
    tl.tile_sync();
    bool anynear = false;
    for (int c = 0; c < sr[3]; ++c) anynear = anynear || F.near[1 + c];
    if (!anynear) return false;
    const double* R1 = F.R[0]; const double* ph1 = F.ph[0];
    for (int m = tl.lane; m < mc; m += Tile::LPE) {
      const double* mk = S.mk + KM_STRIDE * (mo + m);
      const double* xi1 = mk;
      const double* wb = wbar + 3 * (mo + m);
      MarkerHit H;
      double F1[3];
      marker_force(S, F, sr, sd, xi1, H, F1);
      if (H.cand < 0) continue;
      hits += 1.0;
      double* A = acc[H.cand].v;
      const double* R2 = F.R[1 + H.cand]; const double* ph2 = F.ph[1 + H.cand];
      const double* e = H.e;
      double R21[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R21[3 * i + j] = R2[i] * R1[j] + R2[3 + i] * R1[3 + j] + R2[6 + i] * R1[6 + j];
      // F1 = R21^T Fb ; tau = P^T F1 with P = [axis0, axis1, -normal] of this marker
      double F1b[3], Fbb[3], Fb[3];
      for (int i = 0; i < 3; ++i) F1b[i] = mk[3 + i] * wb[0] + mk[6 + i] * wb[1] - mk[9 + i] * wb[2];
      mv3(R21, F1b, Fbb);
      // recompute Fb (candidate frame) for the R21 cotangent
      for (int i = 0; i < 3; ++i) Fb[i] = -(H.s * e[i]);
      double sbar, tbar[3];
      if (!(mu > TS_EPS)) {
        sbar = -dot3(e, Fbb);
        tbar[0] = tbar[1] = tbar[2] = 0.0;
      } else if (!H.dynamic) {
        for (int i = 0; i < 3; ++i) Fb[i] -= kt * H.tb[i];
        sbar = -dot3(e, Fbb);
        for (int i = 0; i < 3; ++i) tbar[i] = -kt * Fbb[i];
      } else {
        const double as = fabs(H.s), sgn = H.s < 0.0 ? -1.0 : 1.0;
        const double sc = mu * as / H.tn;
        for (int i = 0; i < 3; ++i) Fb[i] -= sc * H.tb[i];
        const double tF = dot3(H.tb, Fbb);
        sbar = -dot3(e, Fbb) - mu * sgn * tF / H.tn;
        for (int i = 0; i < 3; ++i) tbar[i] = -mu * as * (Fbb[i] / H.tn - H.tb[i] * tF / (H.tn * H.tn * H.tn));
      }
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[3 * i + j] += Fb[i] * F1b[j];
      double dbar = sbar * (kn - damp * H.ddot);
      double ddbar = -sbar * damp * H.d;
      // tb = (I - e e^T) u + d (w2 x e)
      double ubar[3], w2e[3], t3[3];
      const double etb = dot3(e, tbar);
      for (int i = 0; i < 3; ++i) ubar[i] = tbar[i] - e[i] * etb;
      cross3(ph2, e, w2e);
      dbar += dot3(tbar, w2e);
      cross3(e, tbar, t3);
      double w2bar[3], v2bar[3], xbar[3];
      for (int i = 0; i < 3; ++i) w2bar[i] = H.d * t3[i];
      if (KT_CYLINDER && H.cyl) {
        // the normal e = x_r / |x_r| moves with the point: Fb = -s e - ..., ddot = e.u, tb = u - e (e.u) + d (w2 x e)
        double ebar[3], tw[3];
        const double eu = dot3(e, H.u);
        cross3(tbar, ph2, tw);
        for (int i = 0; i < 3; ++i) ebar[i] = -H.s * Fbb[i] + ddbar * H.u[i] - eu * tbar[i] - etb * H.u[i] + H.d * tw[i];
        if (!(KT_SPHERE && H.sph)) ebar[2] = 0.0;        // cylinder: e is radial in the x-y plane; sphere: in all three
        const double ee = dot3(e, ebar);
        for (int i = 0; i < 3; ++i) xbar[i] = dbar * e[i] + (ebar[i] - e[i] * ee) / H.rad;
      } else {
        for (int i = 0; i < 3; ++i) xbar[i] = dbar * e[i];
      }
      for (int i = 0; i < 3; ++i) ubar[i] += ddbar * e[i];
      // u = R21 w - w2 x x - v2 ,  w = w1 x xi1 + v1
      double w[3], wbar_[3];
      cross3(ph1, xi1, w);
      for (int i = 0; i < 3; ++i) w[i] += ph1[3 + i];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[3 * i + j] += ubar[i] * w[j];
      mtv3(R21, ubar, wbar_);
      cross3(ubar, H.x, t3);
      for (int i = 0; i < 3; ++i) w2bar[i] += t3[i];
      cross3(ph2, ubar, t3);
      for (int i = 0; i < 3; ++i) { xbar[i] += t3[i]; v2bar[i] = -ubar[i]; }
      cross3(xi1, wbar_, t3);
      for (int i = 0; i < 3; ++i) { A[12 + i] += t3[i]; A[15 + i] += wbar_[i]; }
      for (int i = 0; i < 3; ++i) { A[18 + i] += w2bar[i]; A[21 + i] += v2bar[i]; }
      // x = R21 xi1 + r
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[3 * i + j] += xbar[i] * xi1[j];
      for (int i = 0; i < 3; ++i) A[9 + i] += xbar[i];
    }
  }
  hits = tl.sum(hits);
  if (hits == 0.0) return false;
  for (int c = 0; c < TS_MAXCAND; ++c) for (int i = 0; i < 24; ++i) acc[c].v[i] = tl.sum(acc[c].v[i]);
  return true;
}

// body-frame Jacobian column from a Dual pose: w = axial(R^T dR), v = R^T dp
HD void jac_col(const Dual* R, const Dual* p, double* J) {
  J[0] = R[2].v * R[1].d + R[5].v * R[4].d + R[8].v * R[7].d;   // (R^T dR)[2][1]
  J[1] = R[0].v * R[2].d + R[3].v * R[5].d + R[6].v * R[8].d;   // (R^T dR)[0][2]
  J[2] = R[1].v * R[0].d + R[4].v * R[3].d + R[7].v * R[6].d;   // (R^T dR)[1][0]
  J[3] = R[0].v * p[0].d + R[3].v * p[1].d + R[6].v * p[2].d;
  J[4] = R[1].v * p[0].d + R[4].v * p[1].d + R[7].v * p[2].d;
  J[5] = R[2].v * p[0].d + R[5].v * p[1].d + R[8].v * p[2].d;
}

// One reverse step of the BDF1 adjoint (DH/Simulation.cpp:1619-1713 / :1921-1971) in
// "pending contribution" form: pendA = everything later steps add to y_k, pendB = what step
// k+1 adds to y_{k-1}.  Both are distributed (slot c of lane l holds dof l + c*LPE).
//   y_k = df_dq_k + dvar_dq^T b + dtac_dq^T w + (1/h) dtac_dqdot^T w - pendA
//   z_k = H_k^-T y_k ;  df_du_k = h^2 dfr_du^T z_k
//   pendA' = pendB + (G0 + G1/h)^T z_k + (1/h) dtac_dqdot^T w ;  pendB' = -(G1/h)^T z_k
// out_g0z / out_g1z (optional, distributed) expose G0^T z (+ pending terms) and G1^T z for the
// q0 / qdot0 gradients of the first step.
// Cotangent pull-back of the readouts of ONE state (the tile state holds q, qd): per owned dof k,
//   yk = d(var)/dq_k . w_var + d(tactile)/dq_k . w_tac   and   ck = d(tactile)/dqdot_k . w_tac
// (the rows dvar_dq^T w, dtactile_dq^T w, dtactile_dqdot^T w of DH/Simulation.cpp:1640-1650).  It depends on the state
// and the cotangents only, not on the adjoint recursion: tsim_backward evaluates it for all env-steps in a balanced
// pass of its own (vjp_kernel) and the reverse sweep reads the two n-vectors.
template <class Tile, class WK>
// light_only: when a candidate body can reach a pad (the expensive case: marker loop + reverse mode) nothing is
// evaluated and true is returned -- the caller defers the env-step to a pass that holds only such env-steps.
HDN bool vjp_terms(const Tile& tl, const SceneView& S, const double* dvar_cot, const double* dtac_cot, WK& WD,
                   double* yk_out, double* ck_out, bool light_only = false) {
  const int L = Tile::LPE;
  for (int c = 0; c < TS_NC(L); ++c) { yk_out[c] = 0.0; ck_out[c] = 0.0; }
  const bool have_var = dvar_cot != 0 && S.nee > 0;
  bool have_tac = dtac_cot != 0 && S.nmark > 0;
  if (have_var || have_tac) {
    TacAcc acc[TS_MAXCAND];
    // unit tangent on q_k; the values come from the tile state (written by the caller)
    TileState& ts = WD.state();
    SeedIn in;
    in.xq = ts.q; in.xv = ts.qd; in.xl = ts.qd;      // dl is not read by the kinematics-only pass
    in.tq = 1.0; in.tv = 0.0; in.tl = 0.0;
    in.q0v = ts.q; in.qd0v = ts.qd; in.tq0 = 0.0; in.tqd0 = 0.0;
    for (int c = 0; c < TS_NC(L); ++c) {
      const int k = tl.lane + c * L;
      in.k = k;
      tl.tile_sync();
      kinematics(S, in, WD, false);
      if (light_only && have_tac && c == 0) {
        bool anynear = false;
        for (int si = 0; si < S.nsens; ++si) {
          const int* sr = S.ib + S.o_sensor + si * KS_ISTRIDE;
          const double* sd = S.db + S.d_sensor + si * KS_DSTRIDE;
          Frames& F = WD.frames();
          tl.tile_sync();
          sensor_frames(S, WD, sr, sd, F);
          tl.tile_sync();
          for (int cc = 0; cc < sr[3]; ++cc) anynear = anynear || F.near[1 + cc];
        }
        if (anynear) return true;
      }
      double yk = 0.0, ck = 0.0;
      if (have_var) {
        for (int e = 0; e < S.nee; ++e) {
          Dual v[3];
          variable_of(S, WD, e, v);
          yk += v[0].d * dvar_cot[3 * e] + v[1].d * dvar_cot[3 * e + 1] + v[2].d * dvar_cot[3 * e + 2];
        }
      }
      if (have_tac) {
        for (int si = 0; si < S.nsens; ++si) {
          // the values of the Dual work space are the kinematics of the state: the marker pass reads them
          if (!tactile_vjp(tl, S, WD, si, dtac_cot, acc)) continue;
          const int* sr = S.ib + S.o_sensor + si * KS_ISTRIDE;
          const int b1 = sr[0], nc = sr[3];
          Dual R1[9], p1[3], ph1[6];
          body_frame(S, WD, b1, R1, p1, ph1);
          double J1[6];
          jac_col(R1, p1, J1);
          for (int ci = 0; ci < nc; ++ci) {
            const double* A = acc[ci].v;
            double nz = 0.0;
            for (int i = 0; i < 24; ++i) nz += fabs(A[i]);
            if (nz == 0.0) continue;
            const int b2 = sr[4 + ci];
            Dual R2[9], p2[3], ph2[6];
            body_frame(S, WD, b2, R2, p2, ph2);
            // R21 = R2^T R1, r = R2^T (p1 - p2) on Duals; take the tangents
            for (int i = 0; i < 3; ++i)
              for (int j = 0; j < 3; ++j) {
                Dual r = R2[i] * R1[j] + R2[3 + i] * R1[3 + j] + R2[6 + i] * R1[6 + j];
                yk += A[3 * i + j] * r.d;
              }
            Dual dp[3], rr[3];
            for (int i = 0; i < 3; ++i) dp[i] = p1[i] - p2[i];
            mtv3(R2, dp, rr);
            double J2[6];
            jac_col(R2, p2, J2);
            for (int i = 0; i < 3; ++i) yk += A[9 + i] * rr[i].d;
            for (int i = 0; i < 6; ++i) {
              yk += A[12 + i] * ph1[i].d + A[18 + i] * ph2[i].d;
              ck += A[12 + i] * J1[i] + A[18 + i] * J2[i];
            }
          }
        }
      }
      yk_out[c] = yk;
      ck_out[c] = ck;
    }
  }
  return false;
}

template <class Tile, class WK>
HDN void step_backward(const Tile& tl, const SceneView& S, const double* uk, const double* tape, const double* dq_cot, const double* dvar_cot, const double* dtac_cot,
                       double* pendA, double* pendB, double* du_out, double* out_g0z, double* out_g1z,
                       WK& WD, const double* pre_y = 0, const double* pre_c = 0) {
  const int L = Tile::LPE;
  const int n = S.n;
  double y[TS_NC(L)], cterm[TS_NC(L)];
  for (int c = 0; c < TS_NC(L); ++c) {
    const int k = tl.lane + c * L;
    y[c] = (k < n && dq_cot) ? dq_cot[k] : 0.0;
    if (k < n) y[c] -= pendA[c];
    cterm[c] = 0.0;
  }
  // readout pull-back: precomputed by the balanced pass (pre_y, pre_c: n doubles each), or evaluated here
  {
    double yk[TS_NC(L)], ck[TS_NC(L)];
    if (pre_y) {
      for (int c = 0; c < TS_NC(L); ++c) {
        const int k = tl.lane + c * L;
        yk[c] = (k < n) ? pre_y[k] : 0.0;
        ck[c] = (k < n) ? pre_c[k] : 0.0;
      }
    } else {
      vjp_terms(tl, S, dvar_cot, dtac_cot, WD, yk, ck);
    }
    if ((dvar_cot != 0 && S.nee > 0) || (dtac_cot != 0 && S.nmark > 0)) {
      for (int c = 0; c < TS_NC(L); ++c) {
        const int k = tl.lane + c * L;
        if (k < n) { cterm[c] = ck[c] / S.h; y[c] += yk[c] + cterm[c]; }
      }
    }
  }
  // solve H^T z = y
  double z[TS_MAXN];
  bool solved = false;
  if (L >= TS_MAXN) {
    // row-owner elimination: lane i holds row i of H^T = column i of H, and y_i
    double a[TS_MAXN];
    for (int c = 0; c < TS_MAXN; ++c)
      a[c] = (tl.lane < n && c < n) ? tape[c * n + tl.lane] : ((c == tl.lane) ? 1.0 : 0.0);
#if TS_LU_SMEM && !defined(TS_NO_LU_PIVOT)
    tl.tile_sync();
    lu_rows_solve_pivot(tl, a, (tl.lane < n) ? y[0] : 0.0, z, WD.scratch());
    solved = true;
    tl.tile_sync();
#elif TS_LU_SMEM
    tl.tile_sync();
    solved = !lu_rows_solve_smem(tl, a, (tl.lane < n) ? y[0] : 0.0, z, WD.scratch());
    tl.tile_sync();
#else
    solved = !lu_rows_solve(tl, a, (tl.lane < n) ? y[0] : 0.0, z);
#endif
  }
  if (!solved) {
    // replicated pivoting solve: gather y; the lane owning dof k loads row k of H = column k of H^T
    double col[TS_NC(L)][TS_MAXN];
    for (int i = 0; i < TS_MAXN; ++i) z[i] = 0.0;
    for (int k = 0; k < n; ++k) {
      double v = 0.0;
      for (int c = 0; c < TS_NC(L); ++c) if (tl.lane + c * L == k) v = y[c];
      v = tl.bcast(v, k % L);
      for (int i = 0; i < TS_MAXN; ++i) if (i == k) z[i] = v;
    }
    for (int c = 0; c < TS_NC(L); ++c) {
      const int k = tl.lane + c * L;
      for (int i = 0; i < TS_MAXN; ++i) col[c][i] = (k < n && i < n) ? tape[k * n + i] : 0.0;
    }
    lu_solve(tl, col, z, n);
  }
  // controls: dg/du = -h^2 dfr/du; dfr/du was taped by the forward pass (DH/Actuator/ActuatorMotor.cpp:48-62)
  if (du_out && tl.lane == 0) {
    for (int ai = 0; ai < S.nact; ++ai) {
      const int* r = S.ib + S.o_act + ai * KA_ISTRIDE;
      const int qo = S.ib[S.o_joint + r[0] * KJ_ISTRIDE + 2];
      for (int i = 0; i < r[3]; ++i) {
        const double gain = tape[3 * n * n + r[2] + i];
        double zz = 0.0;
        for (int m = 0; m < TS_MAXN; ++m) if (m == qo + i) zz = z[m];
        du_out[r[2] + i] = S.h * S.h * gain * zz;
      }
    }
  }
  // pending contributions for the earlier steps
  const double* G0 = tape + n * n;
  const double* G1 = tape + 2 * n * n;
  for (int c = 0; c < TS_NC(L); ++c) {
    const int k = tl.lane + c * L;
    if (k >= n) continue;
    double g0z = 0.0, g1z = 0.0;
    for (int i = 0; i < n; ++i) { g0z += G0[i * n + k] * z[i]; g1z += G1[i * n + k] * z[i]; }
    const double oldB = pendB[c];
    pendA[c] = oldB + g0z + g1z / S.h + cterm[c];
    pendB[c] = -g1z / S.h;
    if (out_g0z) out_g0z[c] = g0z + oldB + cterm[c];   // what the step adds to dL/dq0 (to be subtracted)
    if (out_g1z) out_g1z[c] = g1z;
  }
}

// ------------------------------------------------------------------ contact index sets (diagnostic outputs)
// S.cmw words per env-step, force by force in scene order (ground forces first, then the general-primitive
// forces, ceil(points / 32) words each): bit k of a force = its sampled point k is active.
// TactilePush: word 0 = ground force, words 1..3 = pad-box force.
template <class WK>
HDN void contact_sets(const SceneView& S, const WK& W, unsigned* mw) {
  for (int i = 0; i < S.cmw; ++i) mw[i] = 0u;
  double R1[9], p1[3], R2[9], p2[3], ph[6];
  for (int gi = 0; gi < S.nground; ++gi) {
    const int* r = S.ib + S.o_ground + gi * KG_ISTRIDE;
    const int b = r[0], po = r[1], pc = r[2];
    body_frame_v(S, W, b, R1, p1, ph);
    if (KT_SPHERE && pc < 0) {               // sphere: one contact, bit 0
      const double rad = S.db[S.d_body + b * KB_DSTRIDE + KB_HALF];
      const double dc = (S.gn[0] * (p1[0] - S.gx[0]) + S.gn[1] * (p1[1] - S.gx[1]) + S.gn[2] * (p1[2] - S.gx[2])) - rad;
      if (dc <= 0.0) mw[r[3]] |= 1u;
      continue;
    }
    for (int k = 0; k < pc; ++k) {
      const double* xi = S.db + S.d_points + 3 * (po + k);
      double xw[3];
      mv3(R1, xi, xw);
      double d = (xw[0] + p1[0] - S.gx[0]) * S.gn[0] + (xw[1] + p1[1] - S.gx[1]) * S.gn[1] +
                 (xw[2] + p1[2] - S.gx[2]) * S.gn[2];
      if (d <= 0.0) mw[r[3] + (k >> 5)] |= (1u << (k & 31));
    }
  }
  for (int fi = 0; fi < S.ngp; ++fi) {
    const int* r = S.ib + S.o_gp + fi * KP_ISTRIDE;
    const int b1 = r[0], b2 = r[1], po = r[2], pc = r[3];
    const double* hs = S.db + S.d_body + b2 * KB_DSTRIDE + KB_HALF;
    body_frame_v(S, W, b1, R1, p1, ph);
    body_frame_v(S, W, b2, R2, p2, ph);
    for (int k = 0; k < pc; ++k) {
      const double* xi = S.db + S.d_points + 3 * (po + k);
      double xw[3], y[3], x[3];
      mv3(R1, xi, xw);
      bool in;
      if (KT_SPHERE && r[5] == TS_SH_CAPSULE) {
        for (int i = 0; i < 3; ++i) xw[i] = xw[i] + p1[i];
        in = capsule_inside_world(R2, p2, xw, hs);
      } else if (KT_SPHERE && r[5] == TS_SH_SPHERE) {
        for (int i = 0; i < 3; ++i) xw[i] = xw[i] + p1[i];
        in = sphere_inside_world(p2, xw, hs[0]);
      } else if (KT_CYLINDER && r[5] == TS_SH_CYLINDER) {
        for (int i = 0; i < 3; ++i) xw[i] = xw[i] + p1[i];
        in = cylinder_inside_world(R2, p2, xw, hs);
      } else {
        for (int i = 0; i < 3; ++i) y[i] = (xw[i] + p1[i]) - p2[i];
        mtv3(R2, y, x);
        in = cuboid_distance(x, hs) < 0.0;
      }
      if (in) mw[r[4] + (k >> 5)] |= (1u << (k & 31));
    }
  }
}

struct FwdArgs {
  int B, T;
  double* q; double* qd;              // [B,n] state, in/out
  const double* u; long long u_stride; // u[t*u_stride + env*nu + i]
  double* q_traj; double* qd_traj;    // [T,B,n] or null
  double* var_out; const int* var_row; // [rows,B,nvar]; row of step t (null map = t), <0 = skip
  double* tac_out; const int* tac_row; // [rows,B,3M]
  double* tape;                       // [T,B,ntape] or null (ntape = 3 n^2 + nu: H, G0, G1, dfr/du)
  int* status;                        // [T,B] or null
  unsigned* cmask;                    // [T,B,cmw] or null
  int* marker_body;                   // [rows,B,M] (rows as tac_out) or null
  int ls_batch;                       // TSIM_OPT_LS_BATCH
  int max_newton;                     // TSIM_OPT_MAX_NEWTON (0 = the reference's rule)
  // two-state integrators (BDF2): state one step back [B,n], in/out, or null; steps already taken since reset
  double* q_prev; double* qd_prev;
  int steps_done;
  int defer_g0;                       // G0, G1, dfr/du of the tape are written by the pass of their own (env_tape)
  const double* q_start; const double* qd_start;   // [B,n] copy of the state at the start of the call (env_tape, step 0)
  int* tape_order;                    // [T*B] env-steps in the order env_tape takes them: steps in contact first (appended
                                      // from the front), the others from the back; counters work_counter[2], [3]
  int defer_tac;                      // the tactile field is read out by the pass of its own (env_tactile), not by the step loop
  int tac_prezeroed;                  // ... into a buffer that tsim_forward has already set to zero
  unsigned* work_counter;             // dynamic distribution of the env-steps of that pass
  const double* env_db; long long env_stride;   // per-environment lowered double tables [B][env_stride], or null
};

// readouts from a work space that holds the kinematics of the state: variables, tactile field, contact sets
template <class Tile, class WK>
HDN void readout_from_work(const Tile& tl, const SceneView& S, WK& W, double* var_o, double* tac_o,
                           int* mb_o, unsigned* cm_o) {
  typedef typename WK::Scalar T;
  if (var_o && tl.lane == 0)
    for (int e = 0; e < S.nee; ++e) {
      T v[3];
      variable_of(S, W, e, v);
      for (int i = 0; i < 3; ++i) var_o[3 * e + i] = val(v[i]);
    }
  if (tac_o) tactile_values(tl, S, W, tac_o, mb_o);
  if (cm_o && tl.lane == 0) contact_sets(S, W, cm_o);
}

// readouts of a given state (q,qd)
template <class Tile, class WK>
HDN void env_readout(const Tile& tl, const SceneView& S, const double* q, const double* qd, double* var_o,
                     double* tac_o, int* mb_o, unsigned* cm_o, WK& WS) {
  ArrIn<double> in;
  in.q_ = q; in.qd_ = qd; in.dl_ = qd; in.q0_ = q; in.qd0_ = qd;
  kinematics(S, in, WS, false);
  readout_from_work(tl, S, WS, var_o, tac_o, mb_o, cm_o);
}

// T steps of one environment.  Scheduling: the tiles of a warp advance step by step together; the
// warps of a block advance ROUND by round together (one block-wide vote per residual evaluation), so
// the whole block streams through the residual code at the same time -- the kernel is bound by
// instruction fetch otherwise -- while a warp whose environments need extra Newton rounds delays
// only itself, not the block.
template <class Tile, class WK>
HDN void env_forward(const Tile& tl, const SceneView& S, const FwdArgs& a, int env_, WK& WD) {
  const int n = S.n, nu = S.nu, B = a.B;
  const bool active = env_ < B;          // surplus tiles of the last block only keep the votes balanced
  const int env = active ? env_ : B - 1;
  TileState& ts = WD.state();            // tile-uniform step state (shared memory on the GPU)
  for (int i = 0; i < TS_MAXN; ++i) { ts.q[i] = (i < n) ? a.q[(long long)env * n + i] : 0.0; ts.qd[i] = (i < n) ? a.qd[(long long)env * n + i] : 0.0; }
  StepVars v;
#if TS_LS_SERVICE
  if (tl.lane == 0) { ts.ls_req = 0; ts.ls_found = 0; }      // (the first vote of the round loop orders it for the block)
#endif
  v.ls_pending = false; v.ls_fresh = false;
  v.batch_ls = a.ls_batch != 0;
  v.cap_newton = a.max_newton;
  v.mode = TS_ST_BDF1;
  v.defer_g0 = a.defer_g0 != 0;
  int t = 0;                             // step in flight: per tile (TS_TILE_STEPS) or warp-uniform
  bool tile_done = !active;
#if KT_MULTISTEP
  const int integ = S.ib[KI_INTEGRATOR];
  // BDF2 starts with one SDIRK2 step (DH/Simulation.cpp:1079-1086); SDIRK2 runs its two stages every step
  if (integ == TS_INT_SDIRK2 || (integ == TS_INT_BDF2 && a.steps_done == 0)) v.mode = TS_ST_SDIRK_A;
  else if (integ == TS_INT_BDF2) v.mode = TS_ST_BDF2;
  for (int i = 0; i < TS_MAXN; ++i) {
    const bool have = i < n && a.q_prev && a.steps_done > 0;
    ts.pq[i] = have ? a.q_prev[(long long)env * n + i] : 0.0;
    ts.pqd[i] = have ? a.qd_prev[(long long)env * n + i] : 0.0;
  }
#endif
  if (a.T > 0) {
    for (int i = 0; i < TS_MAXU; ++i) ts.u[i] = (i < nu) ? a.u[(long long)env * nu + i] : 0.0;
    tl.tile_sync();
    step_begin(S, v, ts);
  }
  // Scheduling of the rounds (bit-identical results, only the pacing differs):
  //  * TS_ROUNDS_PER_VOTE (default 1 = strict lock-step: one block-wide vote per evaluation round).  With K > 1 a warp
  //    runs up to K rounds between two votes and stops after a round in which one of its tiles had an active
  //    general-primitive contact point (a round in contact costs ~3x a contact-free one), so that contact-free warps
  //    run on while the warps in contact are in their point loop.  MEASURED SLOWER on B200 (forward call, B=4096,
  //    T=200: K=1 105.6 ms, K=2 172, K=3 141, K=4 143, K=6 164): as soon as the warps of an SM are at different places
  //    of the 110 KB residual code, instruction fetch dominates -- the lock-step stays strict.
  //  * TS_TILE_STEPS: the tiles of a warp advance through the time steps independently (t is per tile): a tile whose
  //    step is complete starts its next step in the next round instead of idling until the slowest tile of the warp is done.
#ifndef TS_ROUNDS_PER_VOTE
#define TS_ROUNDS_PER_VOTE 1
#endif
#ifndef TS_TILE_STEPS
#define TS_TILE_STEPS 1
#endif
  for (;;) {
    { TS_TIC(tl); const bool go = tl.cta_any(t < a.T); TS_TOC(tl, 4); if (!go) break; }
#pragma unroll 1
    for (int rep = 0; rep < TS_ROUNDS_PER_VOTE; ++rep) {      // (not unrolled: ONE copy of the residual code)
      bool heavy = false;
      {
        // ONE call site of the evaluation.  Block-cooperative kernels (Tile::COOP) run it in every tile of the block,
        // every round: a tile without an evaluation of its own (idle) still joins the barriers of the cooperative
        // contact-point phase and takes its share of the block's active points.
        TS_TIC2(tl);
        const bool run = t < a.T && !tile_done;
        double cole[TS_NC(Tile::LPE)][TS_MAXN];
        if (run || Tile::COOP) step_eval(tl, S, v, WD, cole, !run);
        if (run) {
          const long long es0 = (long long)t * B + env;
          tile_done = step_post(tl, S, v, a.tape ? a.tape + es0 * S.ntape : (double*)0, WD, cole);
          heavy = WD.gp_any != 0;
        }
        if (Tile::LS_SERVICE) {
          // requests of this round to the line-search service: the whole block serves them, the requesters resume
          TS_CPT0();
          const bool pending = run && v.ls_pending;
          if (ls_service_round(tl, S, pending, LsTag<Tile::LS_SERVICE>()) && pending) {
            const long long es0 = (long long)t * B + env;
            tile_done = step_post(tl, S, v, a.tape ? a.tape + es0 * S.ntape : (double*)0, WD, cole);
          }
          TS_CPT(tl, 14);
        }
        TS_TOC2(tl, 5);
      }
      if (t < a.T) {                       // per tile with TS_TILE_STEPS, warp-uniform otherwise
        const long long es = (long long)t * B + env;
#if TS_TILE_STEPS
        const bool finish = tile_done;
#else
        const bool finish = tl.warp_all(tile_done);
#endif
        if (finish) {
          TS_TIC(tl);
          bool next_stage = false;
#if KT_MULTISTEP
          if (v.mode == TS_ST_SDIRK_A) {
            // first SDIRK2 stage solved: (q_alpha, qd_alpha) become the second state, the second stage starts
            double qa[TS_MAXN], qda[TS_MAXN];
            for (int i = 0; i < TS_MAXN; ++i) {
              qa[i] = (i < n) ? ts.x[i] : 0.0;
              qda[i] = (i < n) ? (ts.x[i] - ts.q[i]) / (TS_SDIRK_ALPHA * S.h) : 0.0;
            }
            tl.tile_sync();
            for (int i = 0; i < TS_MAXN; ++i) { ts.pq[i] = qa[i]; ts.pqd[i] = qda[i]; }
            v.mode = TS_ST_SDIRK_B;
            tl.tile_sync();
            step_begin(S, v, ts);
            tile_done = !active;
            next_stage = true;
          }
#endif
          if (!next_stage) {
            // ---- the step is complete (for this tile / for every tile of this warp)
            if (active) {
              int stat = (v.iters & 0xff) | ((v.ls & 0xff) << 8) | (v.converged ? 0 : TS_STAT_NOT_CONVERGED);
              double qn[TS_MAXN], qdn[TS_MAXN];        // read-modify-write of shared tile state: reads, sync, writes
#if KT_MULTISTEP
              double qo[TS_MAXN], qdo[TS_MAXN];
#endif
              for (int i = 0; i < TS_MAXN; ++i) {
                const double q1 = ts.x[i];
                double xv = 0.0, xl = 0.0;
                if (i < n) stage_inputs(S, ts, v.mode, i, q1, xv, xl);
                qdn[i] = xv;
                qn[i] = (i < n) ? q1 : 0.0;
                if (i < n && !(q1 == q1)) stat |= TS_STAT_NAN;
#if KT_MULTISTEP
                qo[i] = ts.q[i]; qdo[i] = ts.qd[i];
#endif
              }
              tl.tile_sync();
              for (int i = 0; i < TS_MAXN; ++i) {
                ts.q[i] = qn[i]; ts.qd[i] = qdn[i];
#if KT_MULTISTEP
                ts.pq[i] = qo[i]; ts.pqd[i] = qdo[i];      // state one step back (BDF2)
#endif
              }
              if (tl.lane == 0) {
#ifdef __CUDA_ARCH__
                if (a.tape_order) {
                  // contact and contact-free evaluations cost 3:1; the tape pass takes its env-steps grouped by kind so that
                  // the tiles of a warp and the warps of a block run evaluations of the same cost together
                  const long long items = (long long)a.T * B;
                  const unsigned pos = WD.gp_any ? atomicAdd(a.work_counter + 2, 1u)
                                                 : (unsigned)(items - 1) - atomicAdd(a.work_counter + 3, 1u);
                  a.tape_order[pos] = (int)es;
                }
#endif
                if (a.status) a.status[es] = stat;
                if (a.q_traj) for (int i = 0; i < n; ++i) st_stream(a.q_traj + es * n + i, qn[i]);
                if (a.qd_traj) for (int i = 0; i < n; ++i) st_stream(a.qd_traj + es * n + i, qdn[i]);
              }
              const int vr = a.var_out ? (a.var_row ? a.var_row[t] : t) : -1;
              const int tr = (a.tac_out && !a.defer_tac) ? (a.tac_row ? a.tac_row[t] : t) : -1;
              if (vr >= 0 || tr >= 0 || a.cmask) {
                // the work space already holds the kinematics of the new state (last residual evaluation)
                readout_from_work(tl, S, WD,
                                  vr >= 0 ? a.var_out + ((long long)vr * B + env) * 3 * S.nee : (double*)0,
                                  tr >= 0 ? a.tac_out + ((long long)tr * B + env) * 3 * S.nmark : (double*)0,
                                  (tr >= 0 && a.marker_body) ? a.marker_body + ((long long)tr * B + env) * S.nmark : (int*)0,
                                  a.cmask ? a.cmask + es * S.cmw : (unsigned*)0);
              }
            }
            ++t;
#if KT_MULTISTEP
            v.mode = (integ == TS_INT_SDIRK2) ? TS_ST_SDIRK_A : ((integ == TS_INT_BDF2) ? TS_ST_BDF2 : TS_ST_BDF1);
#endif
            if (t < a.T) {
              tl.tile_sync();                    // readouts of this step are done in every lane of the tile
              for (int i = 0; i < TS_MAXU; ++i) ts.u[i] = (i < nu) ? a.u[t * a.u_stride + (long long)env * nu + i] : 0.0;
              step_begin(S, v, ts);
              tile_done = !active;
            }
          }
          TS_TOC(tl, 6);
        }
      }
      // (whole-warp votes, outside of the per-tile control flow)
      if (tl.warp_any(heavy) || !tl.warp_any(t < a.T)) break;
    }
  }
#if KT_MULTISTEP
  tl.tile_sync();
  if (active && tl.lane == 0 && a.q_prev)
    for (int i = 0; i < n; ++i) { a.q_prev[(long long)env * n + i] = ts.pq[i]; a.qd_prev[(long long)env * n + i] = ts.pqd[i]; }
#endif
  tl.tile_sync();
  if (active && tl.lane == 0)
    for (int i = 0; i < n; ++i) { a.q[(long long)env * n + i] = ts.q[i]; a.qd[(long long)env * n + i] = ts.qd[i]; }
}

struct BwdArgs {
  int B, T;
  const double* q_traj; const double* qd_traj;   // [T,B,n] states AFTER each step
  const double* u; long long u_stride;
  const double* tape;                            // [T,B,ntape]
  const double* df_dq; const int* dq_row;        // cotangents [rows,B,*]; row maps as in FwdArgs
  const double* df_dvar; const int* dvar_row;
  const double* df_dtac; const int* dtac_row;
  double* carry;                                 // [B,2,n] pending vectors, in/out
  double* df_du;                                 // [T,B,nu] or null
  double* df_dq0; double* df_dqdot0;             // [B,n] or null: MINUS the adjoint terms of step 0
  double* vjp_y; double* vjp_c;                  // [T,B,n] readout pull-backs (vjp_terms): written by env_vjp, read by the sweep; or null
  unsigned* work_counter;                        // [4] dynamic distribution of the env-steps of the vjp passes
  int* vjp_list;                                 // [T*B] env-steps deferred by the first vjp pass (count: work_counter[2])
  const double* env_db; long long env_stride;    // per-environment lowered double tables [B][env_stride], or null
};

template <class Tile, class WK>
HDN void env_backward(const Tile& tl, const SceneView& S, const BwdArgs& a, int env, WK& WD) {
  const int L = Tile::LPE;
  const int n = S.n, nu = S.nu, B = a.B;
  double pA[TS_NC(L)], pB[TS_NC(L)], g0z[TS_NC(L)], g1z[TS_NC(L)];
  for (int c = 0; c < TS_NC(L); ++c) {
    const int k = tl.lane + c * L;
    pA[c] = (k < n) ? a.carry[((long long)env * 2 + 0) * n + k] : 0.0;
    pB[c] = (k < n) ? a.carry[((long long)env * 2 + 1) * n + k] : 0.0;
    g0z[c] = g1z[c] = 0.0;
  }
  for (int t = a.T - 1; t >= 0; --t) {
    const long long es = (long long)t * B + env;
    TileState& ts = WD.state();
    double u[TS_MAXU];
    tl.tile_sync();                      // the previous (later-in-time) step is done reading the tile state
    for (int i = 0; i < TS_MAXN; ++i) { ts.q[i] = (i < n) ? a.q_traj[es * n + i] : 0.0; ts.qd[i] = (i < n) ? a.qd_traj[es * n + i] : 0.0; }
    for (int i = 0; i < TS_MAXU; ++i) u[i] = (i < nu) ? a.u[t * a.u_stride + (long long)env * nu + i] : 0.0;
    const int r0 = a.df_dq ? (a.dq_row ? a.dq_row[t] : t) : -1;
    const int r1 = a.df_dvar ? (a.dvar_row ? a.dvar_row[t] : t) : -1;
    const int r2 = a.df_dtac ? (a.dtac_row ? a.dtac_row[t] : t) : -1;
    step_backward(tl, S, u, a.tape + es * S.ntape,
                  r0 >= 0 ? a.df_dq + ((long long)r0 * B + env) * n : (const double*)0,
                  r1 >= 0 ? a.df_dvar + ((long long)r1 * B + env) * 3 * S.nee : (const double*)0,
                  r2 >= 0 ? a.df_dtac + ((long long)r2 * B + env) * 3 * S.nmark : (const double*)0,
                  pA, pB, a.df_du ? a.df_du + es * nu : (double*)0, g0z, g1z, WD,
                  a.vjp_y ? a.vjp_y + es * n : (const double*)0, a.vjp_c ? a.vjp_c + es * n : (const double*)0);
  }
  for (int c = 0; c < TS_NC(L); ++c) {
    const int k = tl.lane + c * L;
    if (k >= n) continue;
    a.carry[((long long)env * 2 + 0) * n + k] = pA[c];
    a.carry[((long long)env * 2 + 1) * n + k] = pB[c];
    if (a.df_dq0) a.df_dq0[(long long)env * n + k] = -g0z[c];
    if (a.df_dqdot0) a.df_dqdot0[(long long)env * n + k] = -g1z[c];
  }
}

// Readout pull-back of ONE env-step (item = t * B + env) into vjp_y / vjp_c: the balanced pass of tsim_backward.
template <class Tile, class WK>
HDN bool env_vjp(const Tile& tl, const SceneView& S, const BwdArgs& a, long long item, WK& WD, bool light_only = false) {
  const int L = Tile::LPE;
  const int n = S.n, B = a.B;
  const int t = (int)(item / B);
  TileState& ts = WD.state();
  tl.tile_sync();                        // the previous item of this tile is done reading the tile state
  for (int i = 0; i < TS_MAXN; ++i) { ts.q[i] = (i < n) ? a.q_traj[item * n + i] : 0.0; ts.qd[i] = (i < n) ? a.qd_traj[item * n + i] : 0.0; }
  const int r1 = a.df_dvar ? (a.dvar_row ? a.dvar_row[t] : t) : -1;
  const int r2 = a.df_dtac ? (a.dtac_row ? a.dtac_row[t] : t) : -1;
  const long long env = item - (long long)t * B;
  double yk[TS_NC(L)], ck[TS_NC(L)];
  if (vjp_terms(tl, S, r1 >= 0 ? a.df_dvar + ((long long)r1 * B + env) * 3 * S.nee : (const double*)0,
                r2 >= 0 ? a.df_dtac + ((long long)r2 * B + env) * 3 * S.nmark : (const double*)0, WD, yk, ck, light_only))
    return true;                         // deferred to the pass of the env-steps whose pads can be reached
  for (int c = 0; c < TS_NC(L); ++c) {
    const int k = tl.lane + c * L;
    if (k < n) { a.vjp_y[item * n + k] = yk[c]; a.vjp_c[item * n + k] = ck[c]; }
  }
  return false;
}

// kinematics of the state held in (ts.xq, ts.xv) + tactile field.  Generic work space: the evaluation's own scalar type
// with zero tangents (host harness); the GPU's split work space: VALUES ONLY in the tile's shared region (WorkSharedV).
template <class Tile, class WK>
HD void tactile_of_state(const Tile& tl, const SceneView& S, TileState& ts, WK& WD, double* tac_o, int* mb_o, bool prezeroed) {
  SeedIn in;
  in.xq = ts.xq; in.xv = ts.xv; in.xl = ts.xv;     // dl is not read by the kinematics-only pass
  in.k = -1; in.tq = 0.0; in.tv = 0.0; in.tl = 0.0;
  in.q0v = ts.xq; in.qd0v = ts.xv; in.tq0 = 0.0; in.tqd0 = 0.0;
  kinematics(S, in, WD, false);
  tactile_values(tl, S, WD, tac_o, mb_o, prezeroed);
}
template <class Tile>
HD void tactile_of_state(const Tile& tl, const SceneView& S, TileState& ts, WorkSplit& WD, double* tac_o, int* mb_o, bool prezeroed) {
  WorkSharedV WV;
  WV.sv = WD.sv; WV.fr = WD.fr; WV.ts = WD.ts; WV.beta = 0.0; WV.gp_any = 0;
  ArrIn<double> in;
  in.q_ = ts.xq; in.qd_ = ts.xv; in.dl_ = ts.xv;
  in.q0_ = ts.xq; in.qd0_ = ts.xv;
  kinematics(S, in, WV, false);
  tactile_values(tl, S, WV, tac_o, mb_o, prezeroed);
}

// Tactile field of ONE env-step (item = t * B + env) from the recorded trajectory: the readout pass of tsim_forward.
// Same code and work space as the readout at the end of a step (kinematics of the state, then tactile_values), so the
// field is the same value for value; it only runs where the whole GPU can share it instead of inside the step loop,
// where one warp's readout made the other warps of its block wait.
template <class Tile, class WK>
HDN void env_tactile(const Tile& tl, const SceneView& S, const FwdArgs& a, long long item, WK& WD) {
  const int n = S.n, B = a.B;
  const int t = (int)(item / B);
  const int tr = a.tac_row ? a.tac_row[t] : t;
  if (tr < 0) return;
  const long long env = item - (long long)t * B;
  TileState& ts = WD.state();
  tl.tile_sync();
  for (int i = 0; i < TS_MAXN; ++i) { ts.xq[i] = (i < n) ? a.q_traj[item * n + i] : 0.0; ts.xv[i] = (i < n) ? a.qd_traj[item * n + i] : 0.0; }
  tl.tile_sync();
  tactile_of_state(tl, S, ts, WD, a.tac_out + ((long long)tr * B + env) * 3 * S.nmark,
                   a.marker_body ? a.marker_body + ((long long)tr * B + env) * S.nmark : (int*)0, a.tac_prezeroed != 0);
}

// Adjoint blocks G0 = dg/dq0, G1 = dg/dqdot0 and the control gains of ONE env-step (item = t * B + env) from the recorded
// trajectory: what phase 3 of the step state machine evaluates at the converged point.  They depend on the states
// before and after the step and on the controls only, so tsim_forward evaluates them for all env-steps in a pass of
// its own spread over the whole GPU; the step loop -- sequential per environment, its blocks in lock-step -- loses one
// residual evaluation in about five.
template <class Tile, class WK>
HDN void env_tape(const Tile& tl, const SceneView& S, const FwdArgs& a, long long item, WK& WD) {
  const int n = S.n, nu = S.nu, B = a.B;
  const int t = (int)(item / B);
  const long long env = item - (long long)t * B;
  TileState& ts = WD.state();
  tl.tile_sync();
  const double* qs = t ? a.q_traj + (item - B) * n : a.q_start + env * n;
  const double* qds = t ? a.qd_traj + (item - B) * n : a.qd_start + env * n;
  for (int i = 0; i < TS_MAXN; ++i) {
    ts.q[i] = (i < n) ? qs[i] : 0.0;
    ts.qd[i] = (i < n) ? qds[i] : 0.0;
    ts.x[i] = (i < n) ? a.q_traj[item * n + i] : 0.0;
  }
  for (int i = 0; i < TS_MAXU; ++i) ts.u[i] = (i < nu) ? a.u[t * a.u_stride + env * nu + i] : 0.0;
  StepVars v;
  v.phase = 3; v.mode = TS_ST_BDF1; v.defer_g0 = false; v.batch_ls = false; v.cap_newton = 0;
  v.iters = 0; v.ls = 0; v.trial = 0; v.fail_strike = 0; v.alpha = 1.0; v.gnorm = 0.0; v.converged = true;
  double cole[TS_NC(Tile::LPE)][TS_MAXN];
  tl.tile_sync();
  step_eval(tl, S, v, WD, cole);
  step_post(tl, S, v, a.tape + item * S.ntape, WD, cole);
}
