// TEST-ONLY host harness.  Compiles the very same per-lane math the sm_100a kernels inline
// (tactilesimulation_b200/csrc/sim_core.cuh) with g++ and runs it with a one-lane tile, so the
// kernel arithmetic can be checked against the oracle on a machine without a GPU.
// It is NOT a product path: tactilesimulation_b200/ never loads this library, and the product
// fails loudly when the CUDA extension is missing.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "../../tactilesimulation_b200/csrc/scene_lower.h"
#include "../../tactilesimulation_b200/csrc/sim_core.cuh"

extern "C" {

// words of the contact bitmask output per env-step (tsim_scene_sizes[TSIM_CMASK_WORDS]); < 0: scene rejected
int emu_cmask_words(const int32_t* ibuf, const double* dbuf) {
  KernelTables kt;
  if (!lower_scene(ibuf, 1 << 30, dbuf, 1 << 30, kt).empty()) return -1;
  return kt.ib[KI_CMW];
}

int emu_forward(const int32_t* ibuf, const double* dbuf, int32_t B, int32_t T, double* q, double* qd,
                const double* u, int64_t u_stride, double* q_traj, double* qd_traj, double* var_out,
                const int32_t* var_row, double* tac_out, const int32_t* tac_row, double* tape, int32_t* status,
                uint32_t* cmask, int32_t* marker_body, double* q_prev, double* qd_prev, int32_t steps_done) {
  KernelTables kt;
  if (!lower_scene(ibuf, 1 << 30, dbuf, 1 << 30, kt).empty()) return 1;
  SceneView S;
  scene_view_init(S, kt.ib.data(), kt.db.data());
  FwdArgs a;
  a.B = B; a.T = T; a.q = q; a.qd = qd; a.u = u; a.u_stride = u_stride; a.q_traj = q_traj; a.qd_traj = qd_traj;
  a.var_out = var_out; a.var_row = var_row; a.tac_out = tac_out; a.tac_row = tac_row; a.tape = tape;
  a.status = status; a.cmask = cmask; a.marker_body = marker_body; a.ls_batch = 0; a.max_newton = 0;
  a.q_prev = q_prev; a.qd_prev = qd_prev; a.steps_done = steps_done;
  // as tsim_forward: calls of 4 steps or more read the tactile field out in a pass of its own over the trajectory
  a.defer_tac = (tac_out && q_traj && qd_traj && T >= 4) ? 1 : 0;
  a.work_counter = 0; a.tac_prezeroed = 0;
  // ... and write the G0 / G1 / gain blocks of the tape in a pass of their own (BDF1 scenes)
  std::vector<double> qs(q, q + (size_t)B * S.n), qds(qd, qd + (size_t)B * S.n);
  a.defer_g0 = (tape && q_traj && qd_traj && T >= 4) ? 1 : 0;
  a.q_start = qs.data(); a.qd_start = qds.data(); a.tape_order = 0;
  std::vector<Work<Dual> > wb(1);
  HostTile tl;
  for (int env = 0; env < B; ++env) env_forward(tl, S, a, env, wb[0]);
  if (a.defer_g0)
    for (long long item = 0; item < (long long)T * B; ++item) env_tape(tl, S, a, item, wb[0]);
  if (a.defer_tac)
    for (long long item = 0; item < (long long)T * B; ++item) env_tactile(tl, S, a, item, wb[0]);
  return 0;
}

int emu_backward(const int32_t* ibuf, const double* dbuf, int32_t B, int32_t T, const double* q_traj,
                 const double* qd_traj, const double* u, int64_t u_stride, const double* tape, const double* df_dq,
                 const int32_t* dq_row, const double* df_dvar, const int32_t* dvar_row, const double* df_dtac,
                 const int32_t* dtac_row, double* carry, double* df_du, double* df_dq0, double* df_dqdot0) {
  KernelTables kt;
  if (!lower_scene(ibuf, 1 << 30, dbuf, 1 << 30, kt).empty()) return 1;
  SceneView S;
  scene_view_init(S, kt.ib.data(), kt.db.data());
  BwdArgs a;
  a.B = B; a.T = T; a.q_traj = q_traj; a.qd_traj = qd_traj; a.u = u; a.u_stride = u_stride; a.tape = tape;
  a.df_dq = df_dq; a.dq_row = dq_row; a.df_dvar = df_dvar; a.dvar_row = dvar_row; a.df_dtac = df_dtac;
  a.dtac_row = dtac_row; a.carry = carry; a.df_du = df_du; a.df_dq0 = df_dq0; a.df_dqdot0 = df_dqdot0;
  // the readout pull-backs in a pass of their own, as tsim_backward does (vjp_kernel), then the sweep
  std::vector<double> vy((size_t)T * B * S.n), vc((size_t)T * B * S.n);
  a.vjp_y = vy.data(); a.vjp_c = vc.data(); a.work_counter = 0; a.vjp_list = 0;
  {
    // two phases as vjp_kernel: env-steps whose pads can be reached are deferred to a second sweep
    std::vector<Work<Dual> > wv(1);
    HostTile tv;
    std::vector<long long> deferred;
    for (long long item = 0; item < (long long)T * B; ++item)
      if (env_vjp(tv, S, a, item, wv[0], true)) deferred.push_back(item);
    for (size_t i = 0; i < deferred.size(); ++i) env_vjp(tv, S, a, deferred[i], wv[0], false);
  }
  std::vector<Work<Dual> > wb(1);
  HostTile tl;
  for (int env = 0; env < B; ++env) env_backward(tl, S, a, env, wb[0]);
  return 0;
}

int emu_readout(const int32_t* ibuf, const double* dbuf, int32_t B, const double* q, const double* qd, double* var_out,
                double* tac_out, int32_t* marker_body, uint32_t* cmask) {
  KernelTables kt;
  if (!lower_scene(ibuf, 1 << 30, dbuf, 1 << 30, kt).empty()) return 1;
  SceneView S;
  scene_view_init(S, kt.ib.data(), kt.db.data());
  std::vector<Work<double> > wb(1);
  wb[0].beta = 0.0;
  HostTile tl;
  for (int env = 0; env < B; ++env)
    env_readout(tl, S, q + (long long)env * S.n, qd + (long long)env * S.n,
                var_out ? var_out + (long long)env * 3 * S.nee : 0, tac_out ? tac_out + (long long)env * 3 * S.nmark : 0,
                marker_body ? marker_body + (long long)env * S.nmark : 0, cmask ? cmask + (long long)env * S.cmw : 0, wb[0]);
  return 0;
}
}

extern "C" const char* emu_lower_error(const int32_t* ibuf, const double* dbuf) {
  static std::string err;
  KernelTables kt;
  err = lower_scene(ibuf, 1 << 30, dbuf, 1 << 30, kt);
  return err.c_str();
}
