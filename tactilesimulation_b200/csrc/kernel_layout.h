// Kernel tables: what the sm_100a kernels actually index.  Produced on the host by
// lower_scene() (scene_lower.h) from the portable scene blob (scene_layout.h) at
// tsim_scene_create time; never leaves the library.
//
// Lowering performed (all batch-invariant, done once per scene handle):
//   * fixed joints are folded away: every body / end-effector hangs off its nearest MOVING
//     ancestor joint through a precomposed constant transform, and every moving joint carries the
//     precomposed constant offset from its nearest moving ancestor (the reference walks these
//     identity-motion joints every evaluation, DH/Joint/Joint.cpp:119-165);
//   * the bodies rigidly attached to one moving joint are merged into ONE composite rigid body
//     (6x6 spatial inertia summed in the joint frame), so the dynamics sweep touches each moving
//     joint once instead of each reference body once;
//   * bounding radii of contact point sets / marker grids / boxes for exact-safe culling;
//   * ancestor bitmasks per moving joint (mass-matrix columns).
#pragma once

// Compile-time capacities.  The library holds the kernels twice (build.py): variant 8 (defaults below,
// 8 lanes per environment: TactilePush) and variant 16 (-DTS_VARIANT=16: up to 16 reduced dofs, 16 lanes per
// environment: DClaw, TactileInsertion); csrc/cabi.cpp picks the variant per scene.
#ifndef TS_VARIANT
#define TS_VARIANT 8
#endif
#if TS_VARIANT == 17
// variant 17 = the capacities of variant 16 plus the features of the rolling-ball scene (BASELINE configs[0]):
// sphere primitives, free3d-exp joints, dense point sets, BDF2 / SDIRK2 time integration (the adjoint is BDF1's)
#define KT_MAXJ 12
#define KT_MAXN 16
#define KT_MAXB 24
#define KT_MAXU 16
#define KT_MAXPW 72      // <= 2304 sampled points per general body (20x20x20 cuboid surface grid: 2168)
#define KT_CYLINDER 1
#define KT_MAXCAND 16
#define KT_FREE3D 1
#define KT_POS_MOTOR 1
#define KT_SPHERE 1      // sphere SDF primitives (contact force, tactile candidates, ground contact of a sphere)
#define KT_EXP3D 1       // free3d-exp joints
#define KT_MULTISTEP 1   // BDF2 / SDIRK2 integrators (forward only: the tactile adjoint of the reference is backward_BDF1)
#elif TS_VARIANT == 16
#define KT_MAXJ 12       // moving joints
#define KT_MAXN 16       // reduced dofs
#define KT_MAXB 24       // bodies
#define KT_MAXU 16       // controls
#define KT_MAXPW 5       // 32-bit words of an active-point bitmask: <= 160 sampled points per general body
#define KT_CYLINDER 1    // cylinder SDF primitives (contact force, tactile candidates)
#define KT_MAXCAND 16    // tactile candidate bodies per sensor (StableGrasp: 15)
#define KT_FREE3D 1      // free3d-euler joints
#define KT_POS_MOTOR 1   // position-controlled motors
#else
#define KT_MAXJ 8        // moving joints
#define KT_MAXN 8        // reduced dofs
#define KT_MAXB 16       // bodies
#define KT_MAXU 8        // controls
#define KT_MAXPW 3       // <= 96 sampled points per general body
#define KT_CYLINDER 0    // cuboid primitives only: the TactilePush hot path carries no cylinder code
#define KT_MAXCAND 4     // tactile candidate bodies per sensor
#define KT_FREE3D 0      // revolute / prismatic / planar / translational joints only
#define KT_POS_MOTOR 0   // force-controlled motors only
#endif

#ifndef KT_SPHERE
#define KT_SPHERE 0
#define KT_EXP3D 0
#define KT_MULTISTEP 0
#endif

enum {
  KI_NMJ = 0, KI_N, KI_NU, KI_NEE, KI_NMARK, KI_NGROUND, KI_NGP, KI_NACT, KI_NSENS, KI_MAX_ITER, KI_MAX_LS,
  KI_NBODY, KI_NPOINTS, KI_CMW /* words of the per-env-step contact bitmask output */, KI_INTEGRATOR /* TS_INT_* */,
  KI_O_JOINT = 16, KI_O_BODY, KI_O_GROUND, KI_O_GP, KI_O_ACT, KI_O_EE, KI_O_SENSOR,
  KI_D_JOINT = 24, KI_D_BODY, KI_D_GROUND, KI_D_GP, KI_D_ACT, KI_D_EE, KI_D_SENSOR, KI_D_POINTS, KI_D_MARKERS,
  KI_HEADER = 40
};
enum { KD_H = 0, KD_GRAV = 1, KD_TOL = 4, KD_GN = 5, KD_GX = 8, KD_HEADER = 16 };

// moving joint: int {type, parent (moving index, -1 = world), qoff, ndof, ancestor-or-self bitmask, has_mass}
#define KJ_ISTRIDE 8
// dbl {Ra(9) pa(3): constant offset from the parent moving frame | axis0(3) axis1(3) | damping lo hi limk |
//      composite spatial inertia of every body rigidly attached to this joint, in the joint frame:
//      Ibar(9) rotational inertia about the joint origin, mc(3) = sum m_i c_i, m = sum m_i}
#define KJ_DSTRIDE 40
#define KJ_RA 0
#define KJ_PA 9
#define KJ_AX0 12
#define KJ_AX1 15
#define KJ_DAMP 18
#define KJ_LIMLO 19
#define KJ_LIMHI 20
#define KJ_LIMK 21
#define KJ_IBAR 22
#define KJ_MC 31
#define KJ_MASS 34
// body: int {moving joint (-1 = static), shape, dynamic (has mass and a moving joint), unused}
#define KB_ISTRIDE 4
// dbl {Rmi(9) pmi(3): body frame in its moving joint frame | inertia(6) | cuboid: half-size(3), cylinder: radius,
//      half-length, - | bounding radius}
#define KB_DSTRIDE 24
#define KB_RMI 0
#define KB_PMI 9
#define KB_INERTIA 12
#define KB_HALF 18
#define KB_RBOUND 21
// ground contact: int {body, point_off, point_cnt (-1: the body is a sphere, one state-dependent contact point),
// first word in the contact bitmask output}; dbl {kn kt mu damping}
#define KG_ISTRIDE 4
#define KG_DSTRIDE 4
// general-primitive contact: int {body1, body2, point_off, point_cnt, first word in the contact bitmask output,
// shape of body2, offset (doubles) of the per-word boxes, -}; dbl {kn kt mu damping r_points | bounding box of the
// points in the body-1 frame: lo(3) hi(3)}.  Per-word boxes: lo(3) hi(3) of every 32 consecutive points in the body-1
// frame (dense point sets are walked word by word: a word whose box the primitive cannot reach is skipped).
#define KP_ISTRIDE 8
#define KP_WBOX 6
#define KP_DSTRIDE 12
#define KP_BBOX 5
// actuator: int {moving joint, mode, uoff, ndof}; dbl {cmin[3] cmax[3] P[3] D[3]}
#define KA_ISTRIDE 4
#define KA_DSTRIDE 12
// end effector: int {moving joint (-1 = world), -}; dbl {pos(3) in that frame, -}
#define KE_ISTRIDE 2
#define KE_DSTRIDE 4
// markers: KM_STRIDE doubles each {position(3) axis0(3) axis1(3) normal(3)} in the pad body frame; they stay in
// global memory (read once per marker and readout), everything before them is staged in shared memory
#define KM_STRIDE 12
// sensor: int {body, marker_off, marker_cnt, ncand, cand[KT_MAXCAND]};
// dbl {kn kt mu damping axis0 axis1 normal r_markers | bounding box of the markers in the pad frame: lo(3) hi(3)}
#define KS_ISTRIDE (4 + KT_MAXCAND)
#define KS_DSTRIDE 24
#define KS_RMARK 13
#define KS_BBOX 14
