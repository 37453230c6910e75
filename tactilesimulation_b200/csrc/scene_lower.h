// Host-side lowering of the portable scene blob (scene_layout.h, written by
// tactilesimulation_b200/layout.py) into the kernel tables of kernel_layout.h.
// Plain C++ (no CUDA): used by tsim_scene_create and by the test-only host harness.
#pragma once
#include <math.h>
#include <string>
#include <vector>

#include "kernel_layout.h"
#include "scene_layout.h"

struct KernelTables {
  std::vector<int> ib;
  std::vector<double> db;
  int nj_ref;          // joints (= bodies) of the reference topology
};

namespace tsim_lower {

struct Xf { double R[9]; double p[3]; };

inline Xf xf_identity() {
  Xf e;
  for (int i = 0; i < 9; ++i) e.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  e.p[0] = e.p[1] = e.p[2] = 0.0;
  return e;
}
inline Xf xf_mul(const Xf& a, const Xf& b) {   // a * b
  Xf o;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o.R[3 * i + j] = a.R[3 * i] * b.R[j] + a.R[3 * i + 1] * b.R[3 + j] + a.R[3 * i + 2] * b.R[6 + j];
  for (int i = 0; i < 3; ++i) o.p[i] = a.R[3 * i] * b.p[0] + a.R[3 * i + 1] * b.p[1] + a.R[3 * i + 2] * b.p[2] + a.p[i];
  return o;
}
inline Xf xf_load(const double* R, const double* p) {
  Xf e;
  for (int i = 0; i < 9; ++i) e.R[i] = R[i];
  for (int i = 0; i < 3; ++i) e.p[i] = p[i];
  return e;
}
inline double norm3(const double* v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

}  // namespace tsim_lower

// Returns "" on success, otherwise an error message.
inline std::string lower_scene(const int* ib, long long ni, const double* db, long long nd, KernelTables& out) {
  using namespace tsim_lower;
  if (ni < TS_I_HEADER || ib[TS_I_MAGIC] != TS_MAGIC || ib[TS_I_VERSION] < TS_VERSION_MIN || ib[TS_I_VERSION] > TS_VERSION)
    return "not a scene blob of this version";
  const int si_stride = ib[TS_I_VERSION] >= 5 ? TS_SI_STRIDE : (ib[TS_I_VERSION] == 4 ? TS_SI_STRIDE_V4 : TS_SI_STRIDE_V3);
  const int nj = ib[TS_I_NJ], n = ib[TS_I_NDOF_R], nu = ib[TS_I_NDOF_U], nee = ib[TS_I_NEE], nmark = ib[TS_I_NMARKERS];
  const int nground = ib[TS_I_NGROUND], ngp = ib[TS_I_NGP], nact = ib[TS_I_NACT], nsens = ib[TS_I_NSENSORS];
  const int npoints = ib[TS_I_NPOINTS];
  const int integrator = ib[TS_I_INTEGRATOR];
  if (integrator < TS_INT_BDF1 || integrator > TS_INT_SDIRK2) return "unknown integrator";
  // The blob comes through the public C ABI: every count and every section [offset, offset + count * stride) is checked
  // against the two buffers before anything is indexed (a truncated or stale blob must fail here, not read out of bounds).
  {
    if (nj < 0 || n < 0 || nu < 0 || nee < 0 || nmark < 0 || nground < 0 || ngp < 0 || nact < 0 || nsens < 0 || npoints < 0)
      return "malformed scene blob (negative count)";
    struct Sec { int off_slot; long long count, stride; bool dbl; const char* what; };
    const Sec secs[] = {
      {TS_I_OFF_JOINT, nj, TS_JI_STRIDE, false, "joint records"}, {TS_I_OFF_GROUND, nground, TS_GI_STRIDE, false, "ground contact records"},
      {TS_I_OFF_GP, ngp, TS_PI_STRIDE, false, "general-primitive contact records"}, {TS_I_OFF_ACT, nact, TS_AI_STRIDE, false, "actuator records"},
      {TS_I_OFF_EE, nee, TS_EI_STRIDE, false, "end-effector records"}, {TS_I_OFF_SENSOR, nsens, si_stride, false, "sensor records"},
      {TS_I_DOFF_JOINT, nj, TS_JD_STRIDE, true, "joint data"}, {TS_I_DOFF_GROUND, nground, TS_CD_STRIDE, true, "ground contact data"},
      {TS_I_DOFF_GP, ngp, TS_CD_STRIDE, true, "general-primitive contact data"}, {TS_I_DOFF_ACT, nact, TS_AD_STRIDE, true, "actuator data"},
      {TS_I_DOFF_EE, nee, TS_ED_STRIDE, true, "end-effector data"}, {TS_I_DOFF_SENSOR, nsens, TS_SD_STRIDE, true, "sensor data"},
      {TS_I_DOFF_POINTS, npoints, 3, true, "contact points"}, {TS_I_DOFF_MARKERS, nmark, 3, true, "marker positions"}};
    for (const Sec& c : secs) {
      const long long off = ib[c.off_slot], lim = c.dbl ? nd : ni;
      if (c.count == 0) continue;
      if (off < (c.dbl ? (long long)TS_D_HEADER : (long long)TS_I_HEADER) || off + c.count * c.stride > lim)
        return std::string("malformed scene blob (") + c.what + " out of range)";
    }
    if (ib[TS_I_DOFF_MARKER_AXES] != 0 && (ib[TS_I_DOFF_MARKER_AXES] < TS_D_HEADER || (long long)ib[TS_I_DOFF_MARKER_AXES] + 9ll * nmark > nd))
      return "malformed scene blob (marker axes out of range)";
    if (nd < TS_D_HEADER) return "malformed scene blob (double header)";
  }
  if (integrator != TS_INT_BDF1 && !KT_MULTISTEP) return "scene exceeds the compiled capacity (BDF2 / SDIRK2 integrators)";
  if (nj > KT_MAXB) return "scene exceeds the compiled capacity (bodies)";
  if (n > KT_MAXN || nu > KT_MAXU) return "scene exceeds the compiled capacities (dofs/controls)";
  const int* J = ib + ib[TS_I_OFF_JOINT];
  const double* JD = db + ib[TS_I_DOFF_JOINT];
  // nearest moving ancestor-or-self and the constant transform from its frame to each joint frame
  std::vector<int> mov(nj), midx(nj, -1);
  std::vector<Xf> erel(nj), ea(nj);
  int nmj = 0;
  for (int j = 0; j < nj; ++j) {
    const int jt = J[j * TS_JI_STRIDE], par = J[j * TS_JI_STRIDE + 1];
    if (par >= j || par < -1) return "joints are not in parent-first order";
    if ((jt == TS_JT_FREE3D_EULER || jt == TS_JT_SPHERICAL_EULER) && !KT_FREE3D) return "scene exceeds the compiled capacity (free3d-euler / spherical-euler joints)";
    if ((jt == TS_JT_FREE3D_EXP || jt == TS_JT_SPHERICAL_EXP) && !KT_EXP3D) return "scene exceeds the compiled capacity (free3d-exp / spherical-exp joints)";
    if (jt == TS_JT_FREE2D && !KT_FREE3D) return "scene exceeds the compiled capacity (free2d joints)";
    if (jt < TS_JT_FIXED || jt > TS_JT_FREE2D) return "unknown joint type";
    const Xf e0 = xf_load(JD + j * TS_JD_STRIDE + TS_JD_RPJ, JD + j * TS_JD_STRIDE + TS_JD_PPJ);
    const Xf up = (par < 0) ? e0 : xf_mul(erel[par], e0);
    if (jt == TS_JT_FIXED) {
      mov[j] = (par < 0) ? -1 : mov[par];
      erel[j] = up;
    } else {
      ea[j] = up;
      mov[j] = j;
      erel[j] = xf_identity();
      midx[j] = nmj++;
    }
  }
  if (nmj > KT_MAXJ) return "scene exceeds the compiled capacity (moving joints)";
  auto mv_of = [&](int j) { return (j < 0 || mov[j] < 0) ? -1 : midx[mov[j]]; };

  std::vector<int>& oi = out.ib;
  std::vector<double>& od = out.db;
  oi.assign(KI_HEADER, 0);
  od.assign(KD_HEADER, 0.0);
  out.nj_ref = nj;
  for (int i = 0; i < TS_D_HEADER && i < KD_HEADER; ++i) od[i] = db[i];   // h, gravity, tol, ground: same slots
  oi[KI_NMJ] = nmj; oi[KI_N] = n; oi[KI_NU] = nu; oi[KI_NEE] = nee; oi[KI_NMARK] = nmark; oi[KI_NGROUND] = nground;
  oi[KI_NGP] = ngp; oi[KI_NACT] = nact; oi[KI_NSENS] = nsens; oi[KI_MAX_ITER] = ib[TS_I_MAX_ITER];
  oi[KI_MAX_LS] = ib[TS_I_MAX_LS]; oi[KI_NBODY] = nj; oi[KI_NPOINTS] = npoints;
  oi[KI_INTEGRATOR] = integrator;

  // ---- composite spatial inertia per moving joint: sum over the attached bodies of X^T diag(I_i) X,
  //      X = twist transform joint frame -> body frame (DH/Robot.cpp:652-658 gathers diag(I_i) per body)
  std::vector<std::vector<double> > comp(nmj, std::vector<double>(36, 0.0));
  std::vector<int> has_mass(nmj, 0);
  for (int j = 0; j < nj; ++j) {
    const int m = mv_of(j);
    if (m < 0) continue;
    const double* s = JD + j * TS_JD_STRIDE;
    double mass = 0.0;
    for (int i = 0; i < 6; ++i) mass += fabs(s[TS_JD_INERTIA + i]);
    if (!(mass > 0.0)) continue;
    has_mass[m] = 1;
    const Xf e = xf_mul(erel[j], xf_load(s + TS_JD_RJI, s + TS_JD_PJI));   // body frame in the joint frame
    // X = [[R^T, 0], [-R^T [p]x, R^T]]
    double X[6][6] = {{0}};
    const double px[9] = {0, -e.p[2], e.p[1], e.p[2], 0, -e.p[0], -e.p[1], e.p[0], 0};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        X[a][b] = e.R[3 * b + a];
        X[3 + a][3 + b] = e.R[3 * b + a];
        double t = 0.0;
        for (int c = 0; c < 3; ++c) t += e.R[3 * c + a] * px[3 * c + b];
        X[3 + a][b] = -t;
      }
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) {
        double t = 0.0;
        for (int c = 0; c < 6; ++c) t += X[c][a] * s[TS_JD_INERTIA + c] * X[c][b];
        comp[m][6 * a + b] += t;
      }
  }
  // ---- moving joints
  oi[KI_O_JOINT] = (int)oi.size(); oi[KI_D_JOINT] = (int)od.size();
  std::vector<int> anc(nmj, 0);
  for (int j = 0; j < nj; ++j) {
    if (midx[j] < 0) continue;
    const int m = midx[j];
    const int par = J[j * TS_JI_STRIDE + 1];
    const int pm = mv_of(par);
    anc[m] = (1 << m) | (pm >= 0 ? anc[pm] : 0);
    int rec[KJ_ISTRIDE] = {J[j * TS_JI_STRIDE], pm, J[j * TS_JI_STRIDE + 2], J[j * TS_JI_STRIDE + 3], anc[m], has_mass[m], 0, 0};
    oi.insert(oi.end(), rec, rec + KJ_ISTRIDE);
    double d[KJ_DSTRIDE] = {0};
    for (int i = 0; i < 9; ++i) d[KJ_RA + i] = ea[j].R[i];
    for (int i = 0; i < 3; ++i) d[KJ_PA + i] = ea[j].p[i];
    const double* s = JD + j * TS_JD_STRIDE;
    for (int i = 0; i < 3; ++i) { d[KJ_AX0 + i] = s[TS_JD_AX0 + i]; d[KJ_AX1 + i] = s[TS_JD_AX1 + i]; }
    d[KJ_DAMP] = s[TS_JD_DAMP]; d[KJ_LIMLO] = s[TS_JD_LIMLO]; d[KJ_LIMHI] = s[TS_JD_LIMHI]; d[KJ_LIMK] = s[TS_JD_LIMK];
    // 6x6 = [[Ibar, [mc]x], [[mc]x^T, m 1]]
    const std::vector<double>& C = comp[m];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) d[KJ_IBAR + 3 * a + b] = 0.5 * (C[6 * a + b] + C[6 * b + a]);
    d[KJ_MASS] = (C[6 * 3 + 3] + C[6 * 4 + 4] + C[6 * 5 + 5]) / 3.0;
    d[KJ_MC + 0] = 0.5 * (C[6 * 2 + 4] - C[6 * 1 + 5]);   // [mc]x = [[0,-z,y],[z,0,-x],[-y,x,0]] in the upper-right block
    d[KJ_MC + 1] = 0.5 * (C[6 * 0 + 5] - C[6 * 2 + 3]);
    d[KJ_MC + 2] = 0.5 * (C[6 * 1 + 3] - C[6 * 0 + 4]);
    od.insert(od.end(), d, d + KJ_DSTRIDE);
  }
  // ---- bodies (same order and ids as the reference: body j hangs off joint j)
  oi[KI_O_BODY] = (int)oi.size(); oi[KI_D_BODY] = (int)od.size();
  for (int j = 0; j < nj; ++j) {
    const double* s = JD + j * TS_JD_STRIDE;
    const Xf eji = xf_load(s + TS_JD_RJI, s + TS_JD_PJI);
    const Xf emi = xf_mul(erel[j], eji);
    const int m = mv_of(j);
    double mass = 0.0;
    for (int i = 0; i < 6; ++i) mass += fabs(s[TS_JD_INERTIA + i]);
    int rec[KB_ISTRIDE] = {m, J[j * TS_JI_STRIDE + 4], (m >= 0 && mass > 0.0) ? 1 : 0, 0};
    oi.insert(oi.end(), rec, rec + KB_ISTRIDE);
    double d[KB_DSTRIDE] = {0};
    for (int i = 0; i < 9; ++i) d[KB_RMI + i] = emi.R[i];
    for (int i = 0; i < 3; ++i) d[KB_PMI + i] = emi.p[i];
    for (int i = 0; i < 6; ++i) d[KB_INERTIA + i] = s[TS_JD_INERTIA + i];
    if (J[j * TS_JI_STRIDE + 4] == TS_SH_CAPSULE) {        // blob: (radius, length, -)
      d[KB_HALF] = s[TS_JD_HALF];
      d[KB_HALF + 1] = s[TS_JD_HALF + 1] / 2.;
      d[KB_RBOUND] = d[KB_HALF] + d[KB_HALF + 1];
    } else if (J[j * TS_JI_STRIDE + 4] == TS_SH_CYLINDER) {      // blob: (radius, length, -)
      d[KB_HALF] = s[TS_JD_HALF];
      d[KB_HALF + 1] = s[TS_JD_HALF + 1] / 2.;
      d[KB_RBOUND] = sqrt(d[KB_HALF] * d[KB_HALF] + d[KB_HALF + 1] * d[KB_HALF + 1]);
    } else if (J[j * TS_JI_STRIDE + 4] == TS_SH_SPHERE) {  // blob: (radius, -, -)
      d[KB_HALF] = s[TS_JD_HALF];
      d[KB_RBOUND] = s[TS_JD_HALF];
    } else {
      for (int i = 0; i < 3; ++i) d[KB_HALF + i] = s[TS_JD_HALF + i];
      d[KB_RBOUND] = norm3(s + TS_JD_HALF);
    }
    od.insert(od.end(), d, d + KB_DSTRIDE);
  }
  const double* P = db + ib[TS_I_DOFF_POINTS];
  const double* MK = db + ib[TS_I_DOFF_MARKERS];
  const double* MAX = ib[TS_I_DOFF_MARKER_AXES] > 0 ? db + ib[TS_I_DOFF_MARKER_AXES] : (const double*)0;
  int cmw = 0;               // words of the contact bitmask output, force by force (ground first)
  // ---- ground contacts
  oi[KI_O_GROUND] = (int)oi.size(); oi[KI_D_GROUND] = (int)od.size();
  for (int g = 0; g < nground; ++g) {
    const int* r = ib + ib[TS_I_OFF_GROUND] + g * TS_GI_STRIDE;
    const double* c = db + ib[TS_I_DOFF_GROUND] + g * TS_CD_STRIDE;
    // a sphere touches the ground at ONE state-dependent point (DH/CollisionDetection/CollisionDetection.cpp:17-25)
    const bool sph = J[r[0] * TS_JI_STRIDE + 4] == TS_SH_SPHERE;
    if (sph && !KT_SPHERE) return "scene exceeds the compiled capacity (sphere primitives)";
    int rec[KG_ISTRIDE] = {r[0], r[1], sph ? -1 : r[2], cmw};
    cmw += sph ? 1 : (r[2] + 31) / 32;
    oi.insert(oi.end(), rec, rec + KG_ISTRIDE);
    od.insert(od.end(), c, c + KG_DSTRIDE);
  }
  // ---- general-primitive contacts
  std::vector<int> gp_rec_pos;
  oi[KI_O_GP] = (int)oi.size(); oi[KI_D_GP] = (int)od.size();
  for (int f = 0; f < ngp; ++f) {
    const int* r = ib + ib[TS_I_OFF_GP] + f * TS_PI_STRIDE;
    const double* c = db + ib[TS_I_DOFF_GP] + f * TS_CD_STRIDE;
    const int shape2 = J[r[1] * TS_JI_STRIDE + 4];
    if (shape2 != TS_SH_CUBOID && shape2 != TS_SH_CYLINDER && shape2 != TS_SH_SPHERE && shape2 != TS_SH_CAPSULE)
      return "general-primitive contact: only cuboid, cylinder, sphere and capsule primitives are supported";
    if (shape2 == TS_SH_CAPSULE && !KT_SPHERE) return "scene exceeds the compiled capacity (capsule primitives)";
    if (shape2 == TS_SH_CYLINDER && !KT_CYLINDER) return "scene exceeds the compiled capacity (cylinder primitives)";
    if (shape2 == TS_SH_SPHERE && !KT_SPHERE) return "scene exceeds the compiled capacity (sphere primitives)";
    if (r[3] > 32 * KT_MAXPW) return "scene exceeds the compiled capacity (sampled points per general body)";
    int rec[KP_ISTRIDE] = {r[0], r[1], r[2], r[3], cmw, shape2, 0, 0};
    cmw += (r[3] + 31) / 32;
    gp_rec_pos.push_back((int)oi.size());
    oi.insert(oi.end(), rec, rec + KP_ISTRIDE);
    double d[KP_DSTRIDE] = {c[0], c[1], c[2], c[3], 0};
    for (int i = 0; i < 3; ++i) { d[KP_BBOX + i] = 1e300; d[KP_BBOX + 3 + i] = -1e300; }
    for (int k = 0; k < r[3]; ++k) {
      d[4] = fmax(d[4], norm3(P + 3 * (r[2] + k)));
      for (int i = 0; i < 3; ++i) {
        d[KP_BBOX + i] = fmin(d[KP_BBOX + i], P[3 * (r[2] + k) + i]);
        d[KP_BBOX + 3 + i] = fmax(d[KP_BBOX + 3 + i], P[3 * (r[2] + k) + i]);
      }
    }
    od.insert(od.end(), d, d + KP_DSTRIDE);
  }
  // ---- actuators
  oi[KI_O_ACT] = (int)oi.size(); oi[KI_D_ACT] = (int)od.size();
  for (int a = 0; a < nact; ++a) {
    const int* r = ib + ib[TS_I_OFF_ACT] + a * TS_AI_STRIDE;
    const double* c = db + ib[TS_I_DOFF_ACT] + a * TS_AD_STRIDE;
    if (midx[r[0]] < 0) return "actuator on a fixed joint";
    if (r[1] == TS_ACT_POS && !KT_POS_MOTOR) return "scene exceeds the compiled capacity (position-controlled motors)";
    int rec[KA_ISTRIDE] = {midx[r[0]], r[1], r[2], r[3]};
    oi.insert(oi.end(), rec, rec + KA_ISTRIDE);
    od.insert(od.end(), c, c + KA_DSTRIDE);
  }
  // ---- end effectors
  oi[KI_O_EE] = (int)oi.size(); oi[KI_D_EE] = (int)od.size();
  for (int e = 0; e < nee; ++e) {
    const int j = ib[ib[TS_I_OFF_EE] + e * TS_EI_STRIDE];
    const double* pos = db + ib[TS_I_DOFF_EE] + e * TS_ED_STRIDE;
    const Xf& x = erel[j];
    int rec[KE_ISTRIDE] = {mv_of(j), 0};
    oi.insert(oi.end(), rec, rec + KE_ISTRIDE);
    double d[KE_DSTRIDE] = {0};
    for (int i = 0; i < 3; ++i) d[i] = x.R[3 * i] * pos[0] + x.R[3 * i + 1] * pos[1] + x.R[3 * i + 2] * pos[2] + x.p[i];
    od.insert(od.end(), d, d + KE_DSTRIDE);
  }
  // ---- sensors
  oi[KI_O_SENSOR] = (int)oi.size(); oi[KI_D_SENSOR] = (int)od.size();
  for (int s = 0; s < nsens; ++s) {
    const int* r = ib + ib[TS_I_OFF_SENSOR] + s * si_stride;
    const double* c = db + ib[TS_I_DOFF_SENSOR] + s * TS_SD_STRIDE;
    if (r[3] > KT_MAXCAND) return "scene exceeds the compiled capacity (tactile candidate bodies)";
    for (int k = 0; k < r[3]; ++k) {
      const int sh = J[r[4 + k] * TS_JI_STRIDE + 4];
      if (sh != TS_SH_CUBOID && sh != TS_SH_CYLINDER && sh != TS_SH_SPHERE && sh != TS_SH_CAPSULE) return "tactile candidates must be cuboids, cylinders, spheres or capsules";
      if (sh == TS_SH_CAPSULE && !KT_SPHERE) return "scene exceeds the compiled capacity (capsule primitives)";
      if (sh == TS_SH_CYLINDER && !KT_CYLINDER) return "scene exceeds the compiled capacity (cylinder primitives)";
      if (sh == TS_SH_SPHERE && !KT_SPHERE) return "scene exceeds the compiled capacity (sphere primitives)";
    }
    int rec[KS_ISTRIDE] = {r[0], r[1], r[2], r[3]};
    for (int k = 0; k < r[3]; ++k) rec[4 + k] = r[4 + k];
    oi.insert(oi.end(), rec, rec + KS_ISTRIDE);
    double d[KS_DSTRIDE] = {0};
    for (int i = 0; i < 13; ++i) d[i] = c[i];
    for (int i = 0; i < 3; ++i) { d[KS_BBOX + i] = 1e300; d[KS_BBOX + 3 + i] = -1e300; }
    for (int k = 0; k < r[2]; ++k) {
      d[KS_RMARK] = fmax(d[KS_RMARK], norm3(MK + 3 * (r[1] + k)));
      for (int i = 0; i < 3; ++i) {
        d[KS_BBOX + i] = fmin(d[KS_BBOX + i], MK[3 * (r[1] + k) + i]);
        d[KS_BBOX + 3 + i] = fmax(d[KS_BBOX + 3 + i], MK[3 * (r[1] + k) + i]);
      }
    }
    od.insert(od.end(), d, d + KS_DSTRIDE);
  }
  oi[KI_D_POINTS] = (int)od.size();
  od.insert(od.end(), P, P + 3 * npoints);
  // per-word boxes of the general-primitive point sets
  for (int f = 0; f < ngp; ++f) {
    const int* r = ib + ib[TS_I_OFF_GP] + f * TS_PI_STRIDE;
    oi[gp_rec_pos[f] + KP_WBOX] = (int)od.size();
    for (int w0 = 0; w0 < r[3]; w0 += 32) {
      double bx[6] = {1e300, 1e300, 1e300, -1e300, -1e300, -1e300};
      for (int k = w0; k < r[3] && k < w0 + 32; ++k)
        for (int i = 0; i < 3; ++i) {
          bx[i] = fmin(bx[i], P[3 * (r[2] + k) + i]);
          bx[3 + i] = fmax(bx[3 + i], P[3 * (r[2] + k) + i]);
        }
      od.insert(od.end(), bx, bx + 6);
    }
  }
  oi[KI_CMW] = cmw;
  // markers last (the kernels stage everything before them in shared memory): position + per-marker axes
  oi[KI_D_MARKERS] = (int)od.size();
  for (int s = 0; s < nsens; ++s) {
    const int* r = ib + ib[TS_I_OFF_SENSOR] + s * si_stride;
    const double* c = db + ib[TS_I_DOFF_SENSOR] + s * TS_SD_STRIDE;
    for (int k = r[1]; k < r[1] + r[2]; ++k) {
      od.insert(od.end(), MK + 3 * k, MK + 3 * k + 3);
      if (MAX) od.insert(od.end(), MAX + 9 * k, MAX + 9 * k + 9);
      else od.insert(od.end(), c + 4, c + 13);
    }
  }
  (void)nd;
  return "";
}
