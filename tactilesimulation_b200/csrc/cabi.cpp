// Public C ABI (include/tactilesim_b200.h) of libtactilesim_b200.so.
//
// The kernels are compiled three times from csrc/kernels.cu with different compile-time capacities
// (kernel_layout.h): variant 8 (<= 8 reduced dofs, 8 lanes per environment: TactilePush), variant 16
// (<= 16 reduced dofs, 16 lanes per environment: DClaw, TactileInsertion, StableGrasp) and variant 17 (variant 16
// plus sphere primitives, free3d-exp joints, BDF2 / SDIRK2 integration: RollingBall).  Each variant
// exports the whole ABI under suffixed names (tsim_*_v8 / _v16 / _v17); this file owns the public names, picks
// the smallest variant a scene fits at tsim_scene_create time and forwards every other call to it.
#include <stdint.h>
#include <stdlib.h>
#include <string>
#include "../../include/tactilesim_b200.h"

#define TS_DECLARE_VARIANT(V)                                                                                         \
  extern "C" {                                                                                                        \
  const char* tsim_last_error_v##V(void);                                                                             \
  int tsim_scene_create_v##V(const int32_t*, int64_t, const double*, int64_t, int, void**);                           \
  void tsim_scene_destroy_v##V(void*);                                                                                \
  int tsim_scene_sizes_v##V(const void*, int32_t*);                                                                   \
  int tsim_scene_set_lanes_v##V(void*, int);                                                                          \
  int tsim_scene_set_option_v##V(void*, int, int);                                                                    \
  int tsim_scene_set_env_scenes_v##V(void*, int32_t, const int32_t*, int64_t, const double*, int64_t);                \
  int tsim_forward_multistep_v##V(const void*, int32_t, int32_t, double*, double*, double*, double*, int32_t,        \
                                  const double*, int64_t, double*, double*, double*, const int32_t*, double*,        \
                                  const int32_t*, double*, int32_t*, uint32_t*, int32_t*, void*);                     \
  int tsim_scene_kernel_times_v##V(const void*, double*);                                                             \
  int tsim_variant_lu_solve_v##V(int, int, const double*, const double*, double*);                                    \
  int tsim_readout_v##V(const void*, int32_t, const double*, const double*, double*, double*, int32_t*, uint32_t*,    \
                        void*);                                                                                       \
  int tsim_backward_v##V(const void*, int32_t, int32_t, const double*, const double*, const double*, int64_t,        \
                         const double*, const double*, const int32_t*, const double*, const int32_t*, const double*,  \
                         const int32_t*, double*, double*, double*, double*, void*);                                  \
  }
TS_DECLARE_VARIANT(8)
TS_DECLARE_VARIANT(16)
TS_DECLARE_VARIANT(17)

struct tsim_scene {
  int variant;
  void* inner;
};

static thread_local std::string g_err;
static thread_local int g_err_variant = 0;       // 0: g_err holds the message, else the variant that failed last

static int own_fail(const char* m) { g_err = m; g_err_variant = 0; return 1; }
static int fwd_rc(int rc, int variant) { if (rc) g_err_variant = variant; return rc; }

#define DISPATCH(s, call8, call16, call17) \
  ((s)->variant == 8 ? fwd_rc(call8, 8) : ((s)->variant == 16 ? fwd_rc(call16, 16) : fwd_rc(call17, 17)))

extern "C" {

const char* tsim_last_error(void) {
  if (g_err_variant == 8) return tsim_last_error_v8();
  if (g_err_variant == 16) return tsim_last_error_v16();
  if (g_err_variant == 17) return tsim_last_error_v17();
  return g_err.c_str();
}

int tsim_scene_create(const int32_t* ibuf, int64_t n_int, const double* dbuf, int64_t n_dbl, int device, tsim_scene** out) {
  if (!out) return own_fail("tsim_scene_create: null argument");
  void* inner = 0;
  int variant = 8;
  // TSIM_B200_VARIANT=16: development/test knob that runs a small scene on the 16-dof kernels
  const char* force = getenv("TSIM_B200_VARIANT");
  int rc = 1;
  const bool skip8 = force && atoi(force) >= 16, skip16 = force && atoi(force) == 17;
  if (!skip8) rc = tsim_scene_create_v8(ibuf, n_int, dbuf, n_dbl, device, &inner);
  if (skip8 || (rc != 0 && std::string(tsim_last_error_v8()).find("compiled capacit") != std::string::npos)) {
    variant = 16;
    if (!skip16) rc = tsim_scene_create_v16(ibuf, n_int, dbuf, n_dbl, device, &inner);
    if (skip16 || (rc != 0 && std::string(tsim_last_error_v16()).find("compiled capacit") != std::string::npos)) {
      variant = 17;
      rc = tsim_scene_create_v17(ibuf, n_int, dbuf, n_dbl, device, &inner);
    }
  }
  if (rc) return fwd_rc(rc, variant);
  tsim_scene* s = new tsim_scene();
  s->variant = variant;
  s->inner = inner;
  *out = s;
  return 0;
}

void tsim_scene_destroy(tsim_scene* s) {
  if (!s) return;
  if (s->variant == 8) tsim_scene_destroy_v8(s->inner);
  else if (s->variant == 16) tsim_scene_destroy_v16(s->inner);
  else tsim_scene_destroy_v17(s->inner);
  delete s;
}

int tsim_scene_sizes(const tsim_scene* s, int32_t* out) {
  if (!s) return own_fail("tsim_scene_sizes: null scene");
  return DISPATCH(s, tsim_scene_sizes_v8(s->inner, out), tsim_scene_sizes_v16(s->inner, out), tsim_scene_sizes_v17(s->inner, out));
}

int tsim_scene_set_lanes(tsim_scene* s, int lanes) {
  if (!s) return own_fail("tsim_scene_set_lanes: null scene");
  return DISPATCH(s, tsim_scene_set_lanes_v8(s->inner, lanes), tsim_scene_set_lanes_v16(s->inner, lanes),
                  tsim_scene_set_lanes_v17(s->inner, lanes));
}

int tsim_scene_set_option(tsim_scene* s, int key, int value) {
  if (!s) return own_fail("tsim_scene_set_option: null scene");
  return DISPATCH(s, tsim_scene_set_option_v8(s->inner, key, value), tsim_scene_set_option_v16(s->inner, key, value),
                  tsim_scene_set_option_v17(s->inner, key, value));
}

int tsim_scene_set_env_scenes(tsim_scene* s, int32_t B, const int32_t* ibufs, int64_t n_int, const double* dbufs, int64_t n_dbl) {
  if (!s) return own_fail("tsim_scene_set_env_scenes: null scene");
  return DISPATCH(s, tsim_scene_set_env_scenes_v8(s->inner, B, ibufs, n_int, dbufs, n_dbl),
                  tsim_scene_set_env_scenes_v16(s->inner, B, ibufs, n_int, dbufs, n_dbl),
                  tsim_scene_set_env_scenes_v17(s->inner, B, ibufs, n_int, dbufs, n_dbl));
}

int tsim_forward_multistep(const tsim_scene* s, int32_t B, int32_t T, double* q, double* qd, double* q_prev, double* qd_prev,
                           int32_t steps_done, const double* u, int64_t u_step_stride, double* q_traj, double* qd_traj,
                           double* var_out, const int32_t* var_row, double* tac_out, const int32_t* tac_row, double* tape,
                           int32_t* status, uint32_t* contact_masks, int32_t* marker_body, void* stream) {
  if (!s) return own_fail("tsim_forward: null scene");
#define TS_FWD_ARGS s->inner, B, T, q, qd, q_prev, qd_prev, steps_done, u, u_step_stride, q_traj, qd_traj, var_out, var_row, \
                    tac_out, tac_row, tape, status, contact_masks, marker_body, stream
  return DISPATCH(s, tsim_forward_multistep_v8(TS_FWD_ARGS), tsim_forward_multistep_v16(TS_FWD_ARGS),
                  tsim_forward_multistep_v17(TS_FWD_ARGS));
#undef TS_FWD_ARGS
}

int tsim_forward(const tsim_scene* s, int32_t B, int32_t T, double* q, double* qd, const double* u, int64_t u_step_stride,
                 double* q_traj, double* qd_traj, double* var_out, const int32_t* var_row, double* tac_out,
                 const int32_t* tac_row, double* tape, int32_t* status, uint32_t* contact_masks, int32_t* marker_body,
                 void* stream) {
  return tsim_forward_multistep(s, B, T, q, qd, 0, 0, 0, u, u_step_stride, q_traj, qd_traj, var_out, var_row, tac_out, tac_row,
                                tape, status, contact_masks, marker_body, stream);
}

int tsim_scene_kernel_times(const tsim_scene* s, double* ms) {
  if (!s) return own_fail("tsim_scene_kernel_times: null scene");
  return DISPATCH(s, tsim_scene_kernel_times_v8(s->inner, ms), tsim_scene_kernel_times_v16(s->inner, ms),
                  tsim_scene_kernel_times_v17(s->inner, ms));
}

int tsim_debug_lu_solve(int n, int device, int nsys, const double* A, const double* b, double* x) {
  if (n == 8) return fwd_rc(tsim_variant_lu_solve_v8(device, nsys, A, b, x), 8);
  if (n == 16) return fwd_rc(tsim_variant_lu_solve_v16(device, nsys, A, b, x), 16);
  return own_fail("tsim_debug_lu_solve: n must be 8 or 16 (the dof capacities of the kernel variants)");
}

int tsim_readout(const tsim_scene* s, int32_t B, const double* q, const double* qd, double* var_out, double* tac_out,
                 int32_t* marker_body, uint32_t* contact_masks, void* stream) {
  if (!s) return own_fail("tsim_readout: null scene");
  return DISPATCH(s, tsim_readout_v8(s->inner, B, q, qd, var_out, tac_out, marker_body, contact_masks, stream),
                  tsim_readout_v16(s->inner, B, q, qd, var_out, tac_out, marker_body, contact_masks, stream),
                  tsim_readout_v17(s->inner, B, q, qd, var_out, tac_out, marker_body, contact_masks, stream));
}

int tsim_backward(const tsim_scene* s, int32_t B, int32_t T, const double* q_traj, const double* qd_traj, const double* u,
                  int64_t u_step_stride, const double* tape, const double* df_dq, const int32_t* dq_row,
                  const double* df_dvar, const int32_t* dvar_row, const double* df_dtac, const int32_t* dtac_row,
                  double* carry, double* df_du, double* df_dq0, double* df_dqdot0, void* stream) {
  if (!s) return own_fail("tsim_backward: null scene");
  return DISPATCH(s,
                  tsim_backward_v8(s->inner, B, T, q_traj, qd_traj, u, u_step_stride, tape, df_dq, dq_row, df_dvar, dvar_row,
                                   df_dtac, dtac_row, carry, df_du, df_dq0, df_dqdot0, stream),
                  tsim_backward_v16(s->inner, B, T, q_traj, qd_traj, u, u_step_stride, tape, df_dq, dq_row, df_dvar, dvar_row,
                                    df_dtac, dtac_row, carry, df_du, df_dq0, df_dqdot0, stream),
                  tsim_backward_v17(s->inner, B, T, q_traj, qd_traj, u, u_step_stride, tape, df_dq, dq_row, df_dvar, dvar_row,
                                    df_dtac, dtac_row, carry, df_du, df_dq0, df_dqdot0, stream));
}

}  // extern "C"
