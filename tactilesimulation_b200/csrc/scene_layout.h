// Flat scene ("topology blob") layout shared by the host packer
// (tactilesimulation_b200/layout.py) and the CUDA kernels.  One scene = one int32 buffer
// and one float64 buffer; both are batch-invariant and live in HBM once per handle.
// What each record restates: see tactilesimulation_b200/scene.py (reference citations there).
#pragma once

#define TS_MAGIC 0x54533230  // "TS20"
#define TS_VERSION 5         // sensor records carry up to 4 (version 3), 8 (4) or 16 (5) candidate bodies; all are accepted
#define TS_VERSION_MIN 3

// ---- joint types / shapes / actuator modes (scene.py uses the same values)
#define TS_JT_FIXED 0
#define TS_JT_REVOLUTE 1
#define TS_JT_PRISMATIC 2
#define TS_JT_PLANAR 3
#define TS_JT_TRANSLATIONAL 4
#define TS_JT_FREE3D_EULER 5    // q = (p, r): translation + XYZ Euler angles (DH/Joint/JointFree3DEuler.cpp)
#define TS_JT_FREE3D_EXP 6      // q = (p, r): translation + exponential coordinates (DH/Joint/JointFree3DExp.cpp)
#define TS_JT_SPHERICAL_EULER 7 // q = r: XYZ Euler angles (DH/Joint/JointSphericalEuler.cpp)
#define TS_JT_SPHERICAL_EXP 8   // q = r: exponential coordinates (DH/Joint/JointSphericalExp.cpp)
#define TS_JT_FREE2D 9          // q = (x, y, theta): planar motion in the joint's x-y plane (DH/Joint/JointFree2D.cpp)
#define TS_SH_NONE 0
#define TS_SH_CUBOID 1
#define TS_SH_CYLINDER 2
#define TS_SH_SPHERE 3
#define TS_SH_CAPSULE 4         // axis z; blob: (radius, length, -)   (DH/Body/BodyCapsule.cpp)
// time integrators (DH/Simulation.cpp:1076-1092); header slot TS_I_INTEGRATOR, 0 in blobs written before it existed
#define TS_INT_BDF1 0
#define TS_INT_BDF2 1           // first step SDIRK2, then BDF2
#define TS_INT_SDIRK2 2
#define TS_ACT_FORCE 0
#define TS_ACT_POS 1

// (compile-time capacities of the kernels live in kernel_layout.h)

// ---- int header
enum {
  TS_I_MAGIC = 0, TS_I_VERSION, TS_I_NJ, TS_I_NDOF_R, TS_I_NDOF_U, TS_I_NEE, TS_I_NMARKERS,
  TS_I_NGROUND, TS_I_NGP, TS_I_NACT, TS_I_NSENSORS, TS_I_MAX_ITER, TS_I_MAX_LS, TS_I_NPOINTS,
  TS_I_INTEGRATOR, TS_I_RES1,
  // offsets (in elements) of the int sections
  TS_I_OFF_JOINT = 16, TS_I_OFF_GROUND, TS_I_OFF_GP, TS_I_OFF_ACT, TS_I_OFF_EE, TS_I_OFF_SENSOR,
  // offset (doubles) of the optional per-marker (axis0, axis1, normal) section, 9 doubles per marker; 0 = the
  // per-sensor axes of the sensor record apply to all its markers (blobs written before abstract sensors)
  TS_I_DOFF_MARKER_AXES = 22,
  // int offset of the per-marker image positions (row, col); host-side only (flow images), 0 = none
  TS_I_OFF_MARKER_IMAGE = 23,
  // offsets (in elements) of the double sections
  TS_I_DOFF_JOINT = 24, TS_I_DOFF_GROUND, TS_I_DOFF_GP, TS_I_DOFF_ACT, TS_I_DOFF_EE,
  TS_I_DOFF_SENSOR, TS_I_DOFF_POINTS, TS_I_DOFF_MARKERS,
  TS_I_HEADER = 32
};

// int records
#define TS_JI_STRIDE 8      // jtype, parent, qoff, ndof, shape, unused x3
#define TS_GI_STRIDE 4      // body, point_off, point_cnt, unused
#define TS_PI_STRIDE 4      // body1, body2, point_off, point_cnt
#define TS_AI_STRIDE 4      // joint, mode, uoff, ndof
#define TS_EI_STRIDE 2      // joint, unused
#define TS_SI_STRIDE_V3 8   // body, marker_off, marker_cnt, ncand, cand[4]
#define TS_SI_STRIDE_V4 12  // body, marker_off, marker_cnt, ncand, cand[8]
#define TS_SI_STRIDE 20     // body, marker_off, marker_cnt, ncand, cand[16]

// double header: h, g(3), tol, ground normal(3), ground origin(3)
enum { TS_D_H = 0, TS_D_GRAV = 1, TS_D_TOL = 4, TS_D_GN = 5, TS_D_GX = 8, TS_D_HEADER = 16 };

// double records
// joint: E_pj0 R(9 row-major) p(3) | axis0(3) axis1(3) | damping lim_lo lim_hi lim_k |
//        E_ji R(9) p(3) | inertia(6) | half-size(3) (cuboid) / r,l (cylinder) | pad
#define TS_JD_STRIDE 48
#define TS_JD_RPJ 0
#define TS_JD_PPJ 9
#define TS_JD_AX0 12
#define TS_JD_AX1 15
#define TS_JD_DAMP 18
#define TS_JD_LIMLO 19
#define TS_JD_LIMHI 20
#define TS_JD_LIMK 21
#define TS_JD_RJI 22
#define TS_JD_PJI 31
#define TS_JD_INERTIA 34
#define TS_JD_HALF 40
#define TS_CD_STRIDE 4      // kn kt mu damping            (ground, gp)
#define TS_AD_STRIDE 12     // cmin[3] cmax[3] P[3] D[3]
#define TS_ED_STRIDE 4      // pos(3) pad
#define TS_SD_STRIDE 16     // kn kt mu damping | axis0(3) axis1(3) normal(3) | pad
