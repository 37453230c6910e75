/* C ABI of the B200 tactile simulator (libtactilesim_b200.so).
 *
 * Drop-in boundary for ONE path of eanswer/TactileSimulation: the per-timestep differentiable
 * simulation step + dense tactile force field + reverse-time adjoint that the reference reaches
 * through pybind11 (R = /root/reference, DH = R/externals/DiffHand/core/projects/redmax):
 *
 *   tsim_scene_create    <- Simulation(xml_file_path)            DH/python_interface.cpp:86-88
 *                           (the XML is compiled on the host by tactilesimulation_b200.scene;
 *                            the blob replaces the pointer graph of DH/Simulation_Constructor.cpp)
 *   tsim_scene_sizes     <- ndof_r/ndof_m/ndof_u/ndof_var/ndof_tactile  DH/python_interface.cpp:93-98
 *   tsim_forward         <- set_u + forward(num_steps, ..., save_last_frame_var_only) + get_q /
 *                           get_qdot / get_variables / get_tactile_force_vector
 *                           DH/python_interface.cpp:118-135,159,229-231 ; DH/Simulation.cpp:1057-1148
 *   tsim_readout         <- get_variables / get_tactile_force_vector at the current state
 *                           DH/Robot.cpp:359-370
 *   tsim_backward        <- backward() and backward_steps(n) with backward_info.df_dq/df_dvar/
 *                           df_dtactile in, backward_results.df_du/df_dq0/df_dqdot0 out
 *                           DH/python_interface.cpp:64-83,232-236 ; DH/Simulation.cpp:1569-1713,1876-1971
 *
 * A scene handle is not thread-safe (like a Simulation object of the reference): issue the calls on one handle from one
 * thread, on one stream at a time.  tsim_forward / tsim_backward take their scratch (work counters, per-env-step
 * vectors of the trajectory-only passes) from the device's stream-ordered pool (cudaMallocAsync) and return it before
 * they return; nothing is retained between calls except the pool's cache.
 *
 * Conventions: all data pointers are DEVICE pointers (fp64, C-contiguous, env-major inside a
 * step: [step][env][component]); `stream` is a cudaStream_t passed as void*; every call is
 * asynchronous on that stream.  Functions return 0 on success, non-zero on error with a message
 * available from tsim_last_error().  Newton non-convergence is not an error (as in the reference,
 * DH/Simulation.cpp:1218-1222); it is reported per env-step in `status`.
 */
#ifndef TACTILESIM_B200_H
#define TACTILESIM_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct tsim_scene tsim_scene;

/* status word of one env-step: bits 0-7 Newton iterations, bits 8-15 line-search evaluations,
 * bit 16 not converged, bit 17 NaN state */
#define TSIM_STAT_NOT_CONVERGED (1 << 16)
#define TSIM_STAT_NAN (1 << 17)

/* indices into the array filled by tsim_scene_sizes */
enum { TSIM_NJ = 0, TSIM_NDOF_R, TSIM_NDOF_M, TSIM_NDOF_U, TSIM_NDOF_VAR, TSIM_NDOF_TACTILE, TSIM_N_MARKERS,
       TSIM_TAPE_DOUBLES /* per env-step: 3*ndof_r^2 + ndof_u */,
       TSIM_CMASK_WORDS /* 32-bit words of the active contact-point bitmask per env-step */,
       TSIM_INTEGRATOR /* TSIM_INT_*: options.integrator of the scene, DH/python_interface.cpp:36-44 */, TSIM_N_SIZES };
enum { TSIM_INT_BDF1 = 0, TSIM_INT_BDF2 = 1 /* first step SDIRK2, DH/Simulation.cpp:1079-1086 */, TSIM_INT_SDIRK2 = 2 };

const char* tsim_last_error(void);

/* Uploads a packed scene (host pointers; layout: tactilesimulation_b200/csrc/scene_layout.h) to
 * `device` and returns a handle. */
int tsim_scene_create(const int32_t* ibuf, int64_t n_int, const double* dbuf, int64_t n_dbl, int device,
                      tsim_scene** out);
void tsim_scene_destroy(tsim_scene* scene);
int tsim_scene_sizes(const tsim_scene* scene, int32_t* out /* [TSIM_N_SIZES] */);
/* lanes cooperating on one environment: 8, 16 or 32 (default 8) */
int tsim_scene_set_lanes(tsim_scene* scene, int lanes_per_env);
/* Solver options of a scene handle (no reference counterpart: the reference searches sequentially).
 *   TSIM_OPT_LS_BATCH  1 (default): after two rejected line-search trials the lanes of the tile evaluate the
 *                      following step lengths at once; 0: strictly sequential search.  Same accepted step.
 *   TSIM_OPT_MAX_NEWTON 0 (default): the reference's cap on Newton iterations per step, max(20 ndof_r, max_iter)
 *                      (DH/Simulation.cpp:1155); > 0: a lower cap.  A step that hits the cap is reported as not
 *                      converged in `status`, like in the reference; in a lock-step batch one such step can cost as
 *                      much as a whole trajectory, so throughput-bound users may trade exactness on those steps.
 *   TSIM_OPT_VJP_PASS  1 (default): tsim_backward pulls the readout cotangents (df_dvar, df_dtactile) back in a pass of
 *                      its own over all T x B env-steps, spread over the whole GPU, before the reverse sweep (which is
 *                      sequential per environment); 0: inside the sweep.  Bit-identical results.
 *   TSIM_OPT_TAC_PASS  1 (default): tsim_forward calls of 4 steps or more read the tactile field out in a pass of
 *                      their own over the recorded trajectory, spread over the whole GPU, after the step loop;
 *                      0: at the end of every step inside the loop.  Same values.
 *   TSIM_OPT_TAPE_PASS 1 (default): likewise for the blocks G0, G1 and d f_r/d u of the adjoint tape (one residual
 *                      evaluation per env-step at the converged state, which depends on the trajectory only); needs
 *                      q_traj and qd_traj.  0: evaluated inside the step loop.  Same values. */
enum { TSIM_OPT_LS_BATCH = 0, TSIM_OPT_MAX_NEWTON = 1, TSIM_OPT_VJP_PASS = 2, TSIM_OPT_TAC_PASS = 3, TSIM_OPT_TAPE_PASS = 4,
       TSIM_N_OPTS };
int tsim_scene_set_option(tsim_scene* scene, int key, int value);

/* Per-environment parameters (domain randomisation: what the reference does by calling update_joint_damping /
 * update_body_size / update_endeffector_position / update_joint_location / update_body_density / update_contact_parameters /
 * update_tactile_parameters on each environment's own Simulation before reset(): DH/python_interface.cpp:181-211,
 * DH/Robot.cpp:571-650; R/envs/dclaw_rotate_env.py:173-178, stable_grasp_env.py:122-128, tactile_insertion_env.py:254-275).
 *   ibufs [B][n_int], dbufs [B][n_dbl]  (HOST) B packed scenes of the handle's topology, one per environment of the batch
 * The scenes are lowered one by one; counts and connectivity must equal the handle's (values may differ).  Afterwards
 * tsim_forward / tsim_readout / tsim_backward calls with this B read every environment's own parameters (joint damping,
 * contact and tactile coefficients, body sizes / inertias, joint locations, end-effector positions ...); other batch
 * sizes are refused.  B = 0 returns to one parameter set for all environments. */
int tsim_scene_set_env_scenes(tsim_scene* scene, int32_t B, const int32_t* ibufs, int64_t n_int, const double* dbufs,
                              int64_t n_dbl);

/* Advances B environments by T implicit (BDF1/Newton) steps.
 *   q, qd        [B,n]        state, in/out
 *   u            u[t*u_step_stride + env*nu + i]; u_step_stride = 0 holds one action for all T steps
 *   q_traj,qd_traj [T,B,n]    states after each step, or NULL (required later by tsim_backward)
 *   var_out      [rows,B,nvar]   end-effector variables of step t at row var_row[t] (NULL map: row t;
 *                                row < 0: skipped), or NULL
 *   tac_out      [rows,B,ntac]   tactile field (marker-major: shear.axis0, shear.axis1, normal)
 *   tape         [T,B,W]      adjoint tape, W = sizes[TSIM_TAPE_DOUBLES] = 3 n^2 + nu per env-step: H = dg/dq1,
 *                             G0 = dg/dq0, G1 = dg/dqdot0 (n x n, row-major) and d f_r/d u per control; NULL = no-grad mode
 *   status       [T,B] or NULL
 *   contact_masks [T,B,W] or NULL: active contact-point bitmasks, W = sizes[TSIM_CMASK_WORDS] words per
 *                                env-step, force by force in scene order: ground contacts first, then the
 *                                general-primitive contacts, ceil(points/32) words each (TactilePush: W = 4,
 *                                word 0 ground-box, words 1-3 pad-box); bit k = sampled point k is active
 *   marker_body  [rows,B,M] or NULL: contacted body id per marker (-1 none), rows as tac_out */
int tsim_forward(const tsim_scene* scene, int32_t B, int32_t T, double* q, double* qd, const double* u,
                 int64_t u_step_stride, double* q_traj, double* qd_traj, double* var_out, const int32_t* var_row,
                 double* tac_out, const int32_t* tac_row, double* tape, int32_t* status, uint32_t* contact_masks,
                 int32_t* marker_body, void* stream);

/* tsim_forward for scenes whose integrator keeps more than one state (options.integrator = "BDF2", the default of
 * the reference and the integrator of examples/RollingBallExp; DH/Simulation.cpp:1076-1092, 1353-1564): the
 * reference keeps _q_his inside the Simulation, here the caller owns it.
 *   q_prev, qd_prev [B,n]  state one step before (q, qd), in/out: read when steps_done > 0, always written
 *                          (the state before the last step of this call); may be NULL for a whole trajectory
 *                          simulated by ONE call from reset (steps_done = 0)
 *   steps_done             steps already taken since reset(): 0 makes the first step of this call the SDIRK2
 *                          start-up step of BDF2
 * BDF1 scenes ignore the three arguments; tsim_forward(...) is tsim_forward_multistep(..., NULL, NULL, 0, ...).
 * The adjoint tape exists for BDF1 only (tape must be NULL otherwise), like Simulation::backward. */
int tsim_forward_multistep(const tsim_scene* scene, int32_t B, int32_t T, double* q, double* qd, double* q_prev,
                           double* qd_prev, int32_t steps_done, const double* u, int64_t u_step_stride,
                           double* q_traj, double* qd_traj, double* var_out, const int32_t* var_row, double* tac_out,
                           const int32_t* tac_row, double* tape, int32_t* status, uint32_t* contact_masks,
                           int32_t* marker_body, void* stream);

/* Device time (ms, CUDA events on the call's stream) of the kernels of the LAST tsim_forward / tsim_backward call on this
 * handle: ms[TSIM_K_*], -1 for a kernel that did not run.  The caller synchronises the stream first.  (Measurement aid:
 * bench.py's roofline figures; no reference counterpart -- the reference's print_time_report is host timing.) */
enum { TSIM_K_FWD = 0 /* step loop */, TSIM_K_TAPE /* G0 / G1 pass */, TSIM_K_TAC /* tactile readout pass */,
       TSIM_K_VJP /* readout pull-back pass */, TSIM_K_BWD /* reverse sweep */, TSIM_N_KERNELS };
int tsim_scene_kernel_times(const tsim_scene* scene, double* ms /* [TSIM_N_KERNELS] */);

/* Measurement aid: fp64 FMA peak of `device`, measured by a kernel of independent DFMA chains at full occupancy
 * (csrc/microbench.cu).  out[0] = GFLOP/s (FMA = 2 flops), out[1] = kernel ms, out[2] = SM count, out[3] = max SM MHz.
 * The path is fp64 (DH/Common.h:24) and latency / issue bound, not HBM bound: this is the
 * denominator of bench.py's `roofline.fp64`.  Error text: tsim_debug_last_error(). */
int tsim_debug_fp64_peak(int device, double* out /* [4] */);
const char* tsim_debug_last_error(void);

/* Test aid: the n x n solves of the Newton iteration and of the adjoint sweep (row-owner elimination with partial
 * pivoting, csrc/sim_core.cuh lu_rows_solve_pivot; the reference: Eigen partialPivLu, DH/Simulation.cpp:1178, :1640) on
 * `nsys` given systems.  n = 8 or 16 (the dof capacities of the kernel variants; smaller systems are padded with the
 * identity by the caller, as the kernels do).  HOST buffers: A [nsys][n][n] row-major, b [nsys][n], x [nsys][n] out. */
int tsim_debug_lu_solve(int n, int device, int nsys, const double* A, const double* b, double* x);

/* Readouts at a given state (no stepping). Any output may be NULL. */
int tsim_readout(const tsim_scene* scene, int32_t B, const double* q, const double* qd, double* var_out,
                 double* tac_out, int32_t* marker_body, uint32_t* contact_masks, void* stream);

/* Reverse sweep over the T steps recorded by tsim_forward (same B, T, u, q_traj, qd_traj, tape).
 *   df_dq/df_dvar/df_dtac  cotangents [rows,B,*] with optional row maps [T] as above (NULL = no cotangent)
 *   carry        [B,2,n] in/out: pending adjoint contributions of later steps.  Zero it before the
 *                reverse sweep of the LAST chunk of a trajectory and pass it unchanged to the
 *                sweeps of earlier chunks (this is what backward_steps() keeps in BackwardInfo::_z).
 *   df_du        [T,B,nu] out, or NULL
 *   df_dq0, df_dqdot0 [B,n] out, or NULL: gradient w.r.t. the state at the start of the chunk
 *                (meaningful for the first chunk: Simulation::backward's df_dq0 / df_dqdot0) */
int tsim_backward(const tsim_scene* scene, int32_t B, int32_t T, const double* q_traj, const double* qd_traj,
                  const double* u, int64_t u_step_stride, const double* tape, const double* df_dq,
                  const int32_t* dq_row, const double* df_dvar, const int32_t* dvar_row, const double* df_dtac,
                  const int32_t* dtac_row, double* carry, double* df_du, double* df_dq0, double* df_dqdot0,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif
