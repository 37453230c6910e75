// sm_100a kernels + C ABI (include/tactilesim_b200.h) of the B200 tactile simulator.
//
// Execution model: one environment per TILE of LPE lanes (8/16/32) of a warp, one persistent
// tile per environment for all T steps of a call (the tile-uniform state lives in shared memory, the
// scene tables are staged once per CTA in shared memory).  Lane k of a tile owns reduced
// coordinate k: it carries the Dual tangent along q_k through the matrix-free residual, holds
// column k / row k of the Newton matrix for the row-owner LU with partial pivoting, and strides over tactile
// markers / contact points in the readout and adjoint passes.  See sim_core.cuh.
//
// A forward call is three kernels and a backward call three launches: only what is sequential per environment (the
// Newton step loop, the reverse sweep) runs in the persistent one-tile-per-environment kernels fwd_kernel / bwd_kernel;
// what depends on the recorded trajectory only -- tactile read-out (tac_kernel), the G0 / G1 blocks of the tape
// (tape_kernel), the pull-back of the readout cotangents (vjp_kernel, two phases) -- runs over all T x B env-steps in
// persistent kernels that draw env-steps from a work counter.
//
// fwd_kernel: the warps of a block advance evaluation round by evaluation round together (one block-wide vote per
// round: the residual code is ~110 KB, an SM cannot fetch it for seven warps at seven places), the tiles advance
// through their time steps independently, and (variant 8) the active contact points of the whole block are evaluated
// by all its tiles together (gp_points_coop: DevTile<LPE, COOP = true>, CoopArea in shared memory).
// Every kernel exists once more per variant with per-environment parameter tables (template parameter PE,
// tsim_scene_set_env_scenes).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>

#ifndef TS_NO_SYNC_EVALS
#define TS_SYNC_EVALS 1
#endif
#include "kernel_layout.h"
// 224 threads = 7 warps = 28 environments per block: 4096 environments make 147 blocks, one per SM
// (148 SMs), and the block is the lockstep domain of TS_SYNC_EVALS (sim_core.cuh, step_forward).
#ifndef TS_BLOCK
#define TS_BLOCK 224
#endif
// Block-cooperative evaluation of the active contact points in the step loop (sim_core.cuh, gp_points_coop): variant 8
// (cuboid contacts only, one lane per reduced coordinate).  -DTS_NO_COOP_GP builds the A/B library without it.
#if TS_VARIANT == 8 && !defined(TS_NO_COOP_GP)
#define TS_COOP_GP 1
#endif
// this translation unit is one VARIANT of the library (kernel_layout.h): every ABI name gets the variant's
// suffix; csrc/cabi.cpp owns the public names and dispatches per scene
#define TS_CAT_(a, b) a##b
#define TS_CAT(a, b) TS_CAT_(a, b)
#define TSV(x) TS_CAT(x, TS_CAT(_v, TS_VARIANT))
#define tsim_scene TSV(tsim_scene)
#define tsim_last_error TSV(tsim_last_error)
#define tsim_scene_create TSV(tsim_scene_create)
#define tsim_scene_destroy TSV(tsim_scene_destroy)
#define tsim_scene_sizes TSV(tsim_scene_sizes)
#define tsim_scene_set_lanes TSV(tsim_scene_set_lanes)
#define tsim_scene_set_option TSV(tsim_scene_set_option)
#define tsim_scene_set_env_scenes TSV(tsim_scene_set_env_scenes)
#define tsim_forward TSV(tsim_forward)
#define tsim_forward_multistep TSV(tsim_forward_multistep)
#define tsim_readout TSV(tsim_readout)
#define tsim_backward TSV(tsim_backward)
#define tsim_debug_set_prof TSV(tsim_debug_set_prof)
#define tsim_variant_lu_solve TSV(tsim_variant_lu_solve)
#define tsim_scene_kernel_times TSV(tsim_scene_kernel_times)
// Everything below lives in a per-variant namespace: the two variants instantiate templates and kernels
// with identical signatures but different capacities, and must not share symbols.
namespace TSV(tsimns) {
#include "../../include/tactilesim_b200.h"
#include "scene_lower.h"
#include "sim_core.cuh"

#ifdef TS_PROFILE
__device__ long long* g_prof = 0;      // [threads][16] cycle accumulators (development builds only)
#endif

template <int SUBL> struct SubTileOf;       // (defined after DevTile)
template <int LPE_, bool COOP_ = false>
struct DevTile {
  static const int LPE = LPE_;
  static const bool COOP = COOP_;      // the tiles of the block evaluate the active contact points together (fwd_kernel)
  int lane;
  unsigned mask;
  void* coop;                          // CoopArea<LPE> of the block (shared memory), COOP only
  int tile_id;                         // tile index inside the block
  // block-wide line-search service (sim_core.cuh, ls_service_round): tiles with one lane per row, outside the
  // cooperative step loop; the tile states of the block's tiles (shared memory) are ts_stride doubles apart from ts0
  static const int NTILES = TS_BLOCK / LPE_;
  static const bool LS_SERVICE = TS_LS_SERVICE && LPE_ >= TS_MAXN && !COOP_;
  double* ts0;
  int ts_stride;
  __device__ __forceinline__ TileState* peer_state(int r) const { return (TileState*)(ts0 + (size_t)r * ts_stride); }
  // block-wide barriers that the lanes of a warp may reach diverged (tiles run independent control flow)
  __device__ __forceinline__ void cta_sync_unaligned() const { asm volatile("barrier.sync 0;" ::: "memory"); }
  __device__ __forceinline__ bool cta_or_unaligned(bool p) const {
    int r;
    asm volatile("{\n\t.reg .pred pin, pout;\n\tsetp.ne.s32 pin, %1, 0;\n\tbarrier.red.or.pred pout, 0, pin;\n\tselp.s32 %0, 1, 0, pout;\n\t}"
                 : "=r"(r) : "r"((int)p) : "memory");
    return r != 0;
  }
#ifdef TS_PROFILE
  mutable long long acc[16];
#endif
  HD double bcast(double v, int src) const {
#ifdef __CUDA_ARCH__
    return __shfl_sync(mask, v, src, LPE_);
#else
    return v;
#endif
  }
  HD int bcasti(int v, int src) const {
#ifdef __CUDA_ARCH__
    return __shfl_sync(mask, v, src, LPE_);
#else
    return v;
#endif
  }
  HD double sum(double v) const {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = LPE_ / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, LPE_);
#endif
    return v;
  }
  HD void cta_sync() const {
#ifdef __CUDA_ARCH__
    __syncthreads();
#endif
  }
  // the lanes of a tile share the value half of their work space (WorkSplit): converge + order memory
  HD void tile_sync() const {
#ifdef __CUDA_ARCH__
    __syncwarp(mask);
#endif
  }
  // block-wide vote; with TS_NO_SYNC_EVALS (A/B builds) every warp runs free
  HD bool cta_any(bool p) const {
#if defined(__CUDA_ARCH__) && defined(TS_SYNC_EVALS)
    return __syncthreads_or(p) != 0;
#elif defined(__CUDA_ARCH__)
    return __any_sync(0xffffffffu, p) != 0;
#else
    return p;
#endif
  }
  HD bool warp_all(bool p) const {
#ifdef __CUDA_ARCH__
    return __all_sync(0xffffffffu, p) != 0;
#else
    return p;
#endif
  }
  HD bool warp_any(bool p) const {
#ifdef __CUDA_ARCH__
    return __any_sync(0xffffffffu, p) != 0;
#else
    return p;
#endif
  }
  // sub-tile of SUBL consecutive lanes of this tile (SUBL = 1: one lane alone = the host policy): the value-only trial
  // evaluations of the batched line search (sim_core.cuh, trial_norm)
  template <int SUBL> __device__ __forceinline__ typename SubTileOf<SUBL>::type sub() const { return SubTileOf<SUBL>::make(); }
  // ---- whole-warp helpers: every lane of the warp must call them together
  static const int TPW = 32 / LPE_;
  HD int tile_in_warp() const { return (threadIdx.x & 31) / LPE_; }
  // bit t = predicate of tile t (the predicate is uniform inside a tile)
  HD unsigned tiles_ballot(bool p) const {
#ifdef __CUDA_ARCH__
    const unsigned b = __ballot_sync(0xffffffffu, p);
    unsigned r = 0;
#pragma unroll
    for (int t = 0; t < 32 / LPE_; ++t) r |= ((b >> (t * LPE_)) & 1u) << t;
    return r;
#else
    return p ? 1u : 0u;
#endif
  }
  HD double warp_shfl(double v, int src) const {
#ifdef __CUDA_ARCH__
    return __shfl_sync(0xffffffffu, v, src);
#else
    return v;
#endif
  }
  HD unsigned warp_shfl_u(unsigned v, int src) const {
#ifdef __CUDA_ARCH__
    return __shfl_sync(0xffffffffu, v, src);
#else
    return v;
#endif
  }
  // sum over the tiles of the warp of the value held by the same tile-lane
  HD double sum_tiles(double v) const {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = LPE_; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
  }
  // predicate of every lane of the tile, bit i = lane i
  HD unsigned ballot(bool p) const {
#ifdef __CUDA_ARCH__
    const unsigned b = __ballot_sync(mask, p);
    return (LPE_ == 32) ? b : ((b >> ((threadIdx.x & 31) - lane)) & ((1u << LPE_) - 1u));
#else
    return p ? 1u : 0u;
#endif
  }
};

// resident blocks per SM the launch bounds and the shared-memory carve-out are sized for
#ifndef TS_BPS
#define TS_BPS 1
#endif
// ... of the read-out / pull-back passes (tac_kernel, vjp_kernel): value-heavy marker loops that gain from occupancy
// (variant 8: two blocks of 128-register threads per SM, tac_kernel 5.2 -> 4.0 ms, vjp_kernel 6.0 -> 5.6 ms; the tile
// regions of the larger variants do not fit twice in shared memory)
#ifndef TS_PASS_BPS
#if TS_VARIANT == 8
#define TS_PASS_BPS 2
#else
#define TS_PASS_BPS 1
#endif
#endif

// stage the scene blob in shared memory (doubles first, then ints, 8-byte aligned)
// nd = doubles BEFORE the marker table; the markers (the bulk of a scene with dense sensors) stay in global memory
__device__ __forceinline__ void stage_scene(SceneView& S, const int* ib, int ni, const double* db, int nd,
                                            unsigned char* smem) {
  double* sd = (double*)smem;
  int* si = (int*)(smem + (size_t)nd * sizeof(double));
  for (int i = threadIdx.x; i < nd; i += blockDim.x) sd[i] = db[i];
  for (int i = threadIdx.x; i < ni; i += blockDim.x) si[i] = ib[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    scene_view_init(S, si, sd);   // S is the block's shared view: registers stay free for the math
    S.mk = db + si[KI_D_MARKERS];
  }
  __syncthreads();
}

// Shared memory after the scene: one region per tile holding the VALUE half of the Dual work space
// (identical in every lane of the tile) and the sensor/candidate frames.  The stride is odd (in
// doubles) so that the tiles of a warp, which read the same offset of their own region, hit
// different banks.
__host__ __device__ inline int tile_region_doubles(int nmj) {
  return (nmj * WK_REC + (int)((sizeof(Frames) + 7) / 8) + (int)(sizeof(TileState) / 8)) | 1;
}
__host__ __device__ inline size_t scene_bytes(int ni, int nd) {
  return (((size_t)nd * sizeof(double) + (size_t)ni * sizeof(int)) + 15) & ~(size_t)15;
}
template <int LPE>
__device__ __forceinline__ void bind_work(WorkSplit& W, const SceneView& S, int ni, int nd, unsigned char* smem) {
  double* base = (double*)(smem + scene_bytes(ni, nd)) + (size_t)(threadIdx.x / LPE) * tile_region_doubles(S.nj);
  W.sv = base;
  W.ts = (TileState*)(base + S.nj * WK_REC);
  W.fr = (Frames*)(base + S.nj * WK_REC + sizeof(TileState) / 8);
  W.beta = 0.0;
}

template <int SUBL> struct SubTileOf {
  typedef DevTile<SUBL, false> type;
  static __device__ __forceinline__ type make() {
    type t;
    const int wl = threadIdx.x & 31;
    t.lane = wl % SUBL;
    t.mask = (SUBL == 32) ? 0xffffffffu : (((1u << SUBL) - 1u) << (wl - t.lane));
    t.coop = 0;
    t.tile_id = 0;
    t.ts0 = 0;
    t.ts_stride = 0;
    return t;
  }
};
template <> struct SubTileOf<1> {            // one lane alone: the host policy (no cross-lane operations at all)
  typedef HostTile type;
  static __device__ __forceinline__ type make() { return HostTile(); }
};

template <int LPE, bool COOP = false>
__device__ __forceinline__ DevTile<LPE, COOP> make_tile() {
  DevTile<LPE, COOP> tl;
  tl.lane = threadIdx.x % LPE;
  const int wl = threadIdx.x & 31;
  tl.mask = (LPE == 32) ? 0xffffffffu : (((1u << LPE) - 1u) << (wl - tl.lane));
  tl.coop = 0;
  tl.tile_id = threadIdx.x / LPE;
  tl.ts0 = 0;
  tl.ts_stride = 0;
  return tl;
}
// cooperative step loop: one lane per reduced coordinate (one evaluation per round), 8-lane tiles
#ifdef TS_COOP_GP
#define TS_COOP_FOR(LPE) ((LPE) == 8 && TS_MAXN <= 8)
template <int LPE> __host__ __device__ inline size_t coop_bytes() { return TS_COOP_FOR(LPE) ? sizeof(CoopArea<LPE>) : 0; }
#else
#define TS_COOP_FOR(LPE) false
template <int LPE> __host__ __device__ inline size_t coop_bytes() { return 0; }
#endif

// PE (per-environment parameters, tsim_scene_set_env_scenes): every environment has its own lowered double table in
// global memory (read through L1); the tile works on a private copy of the scene view whose table pointer is its
// environment's.  The integer tables (topology) and the marker table are the batch's.  Separate instantiations: the
// broadcast kernels keep the table in shared memory and pay nothing.
#define TS_PE_VIEW(ENV)                                                                   \
  SceneView Sl;                                                                          \
  const SceneView* Sp = &S;                                                              \
  if (PE) { Sl = S; Sl.db = a.env_db + (long long)(ENV) * a.env_stride; Sp = &Sl; }

template <int LPE, bool PE>
__global__ void __launch_bounds__(TS_BLOCK, TS_BPS) fwd_kernel(const int* ib, int ni, const double* db, int nd, FwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ SceneView S;
  stage_scene(S, ib, ni, db, nd, smem);
  const int env = (blockIdx.x * blockDim.x + threadIdx.x) / LPE;
  TS_PE_VIEW(env < a.B ? env : a.B - 1)
  DevTile<LPE, TS_COOP_FOR(LPE) && !PE> tl = make_tile<LPE, TS_COOP_FOR(LPE) && !PE>();
  WorkSplit WD;
  bind_work<LPE>(WD, S, ni, nd, smem);
  // the cooperative area follows the tile regions
  tl.coop = smem + ((scene_bytes(ni, nd) + (size_t)(TS_BLOCK / LPE) * tile_region_doubles(S.nj) * sizeof(double) + 15) & ~(size_t)15);
  tl.ts0 = (double*)(smem + scene_bytes(ni, nd)) + S.nj * WK_REC;       // tile states of the block (bind_work)
  tl.ts_stride = tile_region_doubles(S.nj);
#ifdef TS_PROFILE
  for (int i = 0; i < 16; ++i) tl.acc[i] = 0;
  const long long t_begin = clock64();
#endif
  env_forward(tl, *Sp, a, env, WD);   // tiles past the batch stay in the block-wide votes
#ifdef TS_PROFILE
  tl.acc[7] = clock64() - t_begin;
  if (g_prof) for (int i = 0; i < 16; ++i) g_prof[(long long)(blockIdx.x * blockDim.x + threadIdx.x) * 16 + i] = tl.acc[i];
#endif
}

// Tactile fields of all T x B env-steps of a forward call (env_tactile), one tile per env-step, env-steps drawn from a
// counter like in vjp_kernel.
template <int LPE, bool PE>
__global__ void __launch_bounds__(TS_BLOCK, TS_PASS_BPS) tac_kernel(const int* ib, int ni, const double* db, int nd, FwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ SceneView S;
  stage_scene(S, ib, ni, db, nd, smem);
  DevTile<LPE> tl = make_tile<LPE>();
  WorkSplit WD;
  bind_work<LPE>(WD, S, ni, nd, smem);
  const long long items = (long long)a.T * a.B;
  const int tpw = 32 / LPE;
  for (;;) {
    unsigned base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(a.work_counter, (unsigned)tpw);
    base = __shfl_sync(0xffffffffu, base, 0);
    if ((long long)base >= items) break;
    const long long item = (long long)base + (threadIdx.x & 31) / LPE;
    if (item < items) {
      TS_PE_VIEW(item % a.B)
      env_tactile(tl, *Sp, a, item, WD);
    }
    __syncwarp();
  }
}

// (the adjoint exists for BDF1 scenes; tsim_forward / tsim_backward refuse a tape for the other integrators)
// G0 / G1 / gain blocks of the tape for all T x B env-steps of a forward call (env_tape), env-steps drawn from a counter.
template <int LPE, bool PE>
__global__ void __launch_bounds__(TS_BLOCK, TS_BPS) tape_kernel(const int* ib, int ni, const double* db, int nd, FwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ SceneView S;
  stage_scene(S, ib, ni, db, nd, smem);
  DevTile<LPE> tl = make_tile<LPE>();
  WorkSplit WD;
  bind_work<LPE>(WD, S, ni, nd, smem);
  const long long items = (long long)a.T * a.B;
  unsigned* counter = a.work_counter + 1;
  // the residual code is large: the warps of a block take their env-steps together (one block-wide fetch per round)
  // so that the block streams through it at the same time, as in fwd_kernel
  __shared__ unsigned sbase;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) sbase = atomicAdd(counter, (unsigned)(TS_BLOCK / LPE));
    __syncthreads();
    const unsigned base = sbase;
    if ((long long)base >= items) break;
    const long long slot = (long long)base + threadIdx.x / LPE;
    if (slot < items) {
      const long long item = a.tape_order ? (long long)a.tape_order[slot] : slot;
      TS_PE_VIEW(item % a.B)
      env_tape(tl, *Sp, a, item, WD);
    }
  }
}

template <int LPE, bool PE>
__global__ void __launch_bounds__(TS_BLOCK, TS_BPS) bwd_kernel(const int* ib, int ni, const double* db, int nd, BwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ SceneView S;
  stage_scene(S, ib, ni, db, nd, smem);
  const int env = (blockIdx.x * blockDim.x + threadIdx.x) / LPE;
  if (env >= a.B) return;
  TS_PE_VIEW(env)
  DevTile<LPE> tl = make_tile<LPE>();
  WorkSplit WD;
  bind_work<LPE>(WD, S, ni, nd, smem);
  env_backward(tl, *Sp, a, env, WD);
}

// Readout pull-backs of all T x B env-steps (vjp_terms), one tile per env-step.  The work per env-step depends on
// whether markers are in contact, so the warps draw their env-steps from a counter (TPW consecutive ones at a time:
// consecutive environments of one step, whose loads coalesce) instead of owning a fixed share.
// Two launches: phase 0 takes every env-step, finishes those whose pads nothing can reach (kinematics + variables
// only) and lists the others; phase 1 takes the listed ones -- so the tiles of a warp run work of the same kind.
template <int LPE, bool PE>
__global__ void __launch_bounds__(TS_BLOCK, TS_PASS_BPS) vjp_kernel(const int* ib, int ni, const double* db, int nd, BwdArgs a,
                                                               int phase) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ SceneView S;
  stage_scene(S, ib, ni, db, nd, smem);
  DevTile<LPE> tl = make_tile<LPE>();
  WorkSplit WD;
  bind_work<LPE>(WD, S, ni, nd, smem);
  const long long items = phase ? (long long)a.work_counter[2] : (long long)a.T * a.B;
  const int tpw = 32 / LPE;
  for (;;) {
    unsigned base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(a.work_counter + phase, (unsigned)tpw);
    base = __shfl_sync(0xffffffffu, base, 0);
    if ((long long)base >= items) break;
    const long long slot = (long long)base + (threadIdx.x & 31) / LPE;
    if (slot < items) {
      const long long item = phase ? (long long)a.vjp_list[slot] : slot;
      TS_PE_VIEW(item % a.B)
      const bool deferred = env_vjp(tl, *Sp, a, item, WD, phase == 0);
      if (deferred && tl.lane == 0) a.vjp_list[atomicAdd(a.work_counter + 2, 1u)] = (int)item;
    }
    __syncwarp();
  }
}

struct EnvTables { const double* env_db; long long env_stride; };
template <int LPE, bool PE>
__global__ void __launch_bounds__(TS_BLOCK) readout_kernel(const int* ib, int ni, const double* db, int nd, int B,
                                                          const double* q, const double* qd, double* var_out,
                                                          double* tac_out, int* marker_body, unsigned* cmask, EnvTables a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ SceneView S;
  stage_scene(S, ib, ni, db, nd, smem);
  const int env = (blockIdx.x * blockDim.x + threadIdx.x) / LPE;
  if (env >= B) return;
  TS_PE_VIEW(env)
  DevTile<LPE> tl = make_tile<LPE>();
  Work<double> wb;
  wb.beta = 0.0;
  double ql[TS_MAXN], qdl[TS_MAXN];
  for (int i = 0; i < TS_MAXN; ++i) {
    ql[i] = (i < S.n) ? q[(long long)env * S.n + i] : 0.0;
    qdl[i] = (i < S.n) ? qd[(long long)env * S.n + i] : 0.0;
  }
  env_readout(tl, *Sp, ql, qdl, var_out ? var_out + (long long)env * 3 * S.nee : (double*)0,
              tac_out ? tac_out + (long long)env * 3 * S.nmark : (double*)0,
              marker_body ? marker_body + (long long)env * S.nmark : (int*)0,
              cmask ? cmask + (long long)env * S.cmw : (unsigned*)0, wb);
}

// ------------------------------------------------------------------ host side of the C ABI
struct tsim_scene {
  int device;
  int* d_ib;
  double* d_db;
  int ni, nd, nd_all;
  int lanes;
  int nmj;                 // moving joints of the lowered scene
  int opts[TSIM_N_OPTS];
  int sizes[TSIM_N_SIZES];
  // per-environment parameters (tsim_scene_set_env_scenes): lowered double tables [env_B][nd] on the device, or null
  double* d_env_db;
  int env_B;
  std::vector<int> base_ib;          // lowered integer tables of the handle: every environment must lower to the same
  // events around the kernels of the last tsim_forward / tsim_backward call (tsim_scene_kernel_times)
  cudaEvent_t ev[7];
  mutable int ran[TSIM_N_KERNELS];
};

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }

// Stream-ordered scratch of a call, returned to the pool on EVERY exit path of the call.
struct Scratch {
  void* p;
  cudaStream_t st;
  Scratch(cudaStream_t s) : p(0), st(s) {}
  ~Scratch() { if (p) cudaFreeAsync(p, st); }
};
// The scratch comes from a memory pool of the library's own (one per device, kept between calls: release threshold
// = max), not from the device's default pool, whose settings belong to the application (torch's allocator does not see
// memory cached there).
static cudaError_t scratch_alloc(void** p, size_t bytes, int device, cudaStream_t st) {
  static cudaMemPool_t pools[64] = {0};
  if (device < 0 || device >= 64) return cudaMallocAsync(p, bytes, st);
  if (!pools[device]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    cudaMemPool_t pool;
    cudaError_t e = cudaMemPoolCreate(&pool, &props);
    if (e != cudaSuccess) return e;
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    pools[device] = pool;
  }
  return cudaMallocFromPoolAsync(p, bytes, pools[device], st);
}
#define CK(x)                                                                                          \
  do {                                                                                                 \
    cudaError_t e_ = (x);                                                                              \
    if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_));               \
  } while (0)

// scene tables + one work-space region per tile of the block
static size_t scene_smem(const tsim_scene* s) {
  return scene_bytes(s->ni, s->nd) + (size_t)(TS_BLOCK / s->lanes) * tile_region_doubles(s->nmj) * sizeof(double);
}

template <class K>
static int prep(K kern, size_t smem, int bps = TS_BPS) {
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // one block per SM: ask for no more shared memory than the block uses, the rest of the 256 KB is L1
  // for the per-lane tangents (local memory)
#ifndef TS_NO_CARVEOUT
  int pct = (int)((smem + 1024) * bps * 100 / (228 * 1024)) + 1;
  if (pct > 100) pct = 100;
#ifdef TS_FORCE_CARVEOUT
  pct = TS_FORCE_CARVEOUT;               // A/B builds: how much the L1 left beside the shared memory matters
#endif
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
#endif
  return 0;
}

// Launch KERN<lanes, PE> with the handle's lanes per environment; the per-environment-parameter instantiation exists for
// the variant's own lane count only (TS_MAXN lanes: one per reduced coordinate).
#if TS_MAXN <= 8
#define TS_LAUNCH_LANES8(KERN, GRID, SMEM8, BPS, ...)                                                        \
  if (s->lanes == 8) { if (prep(KERN<8, false>, SMEM8, BPS)) return 1; KERN<8, false><<<GRID, TS_BLOCK, SMEM8, st>>>(__VA_ARGS__); } else
#else
#define TS_LAUNCH_LANES8(KERN, GRID, SMEM8, BPS, ...)
#endif
#define TS_LAUNCH(KERN, GRID, SMEM, SMEM8, BPS, ...)                                                         \
  do {                                                                                                       \
    if (pe) { if (prep(KERN<TS_MAXN, true>, SMEM, BPS)) return 1; KERN<TS_MAXN, true><<<GRID, TS_BLOCK, SMEM, st>>>(__VA_ARGS__); } \
    else TS_LAUNCH_LANES8(KERN, GRID, SMEM8, BPS, __VA_ARGS__)                                               \
    if (s->lanes == 16) { if (prep(KERN<16, false>, SMEM, BPS)) return 1; KERN<16, false><<<GRID, TS_BLOCK, SMEM, st>>>(__VA_ARGS__); } \
    else { if (prep(KERN<32, false>, SMEM, BPS)) return 1; KERN<32, false><<<GRID, TS_BLOCK, SMEM, st>>>(__VA_ARGS__); } \
    CK(cudaGetLastError());                                                                                  \
  } while (0)
// per-environment tables in use for a call of batch B?  (they must have been set for exactly this batch)
#define TS_PE_CHECK(NAME)                                                                                    \
  const bool pe = s->d_env_db != 0;                                                                          \
  if (pe && B != s->env_B) return fail(NAME ": the handle holds per-environment parameters for another batch size"); \
  if (pe && s->lanes != TS_MAXN) return fail(NAME ": per-environment parameters need the default lanes per environment");

// Test aid (tsim_debug_lu_solve): the row-owner elimination with partial pivoting of the Newton / adjoint solves on
// given systems of the variant's capacity, one tile per system.
template <int LPE>
__global__ void lu_debug_kernel(const double* A, const double* b, double* x, int nsys) {
  __shared__ double scr[32 / LPE][TS_MAXN * TS_MAXN];
  DevTile<LPE> tl = make_tile<LPE>();
  const int sys = blockIdx.x * (32 / LPE) + threadIdx.x / LPE;
  const int s = sys < nsys ? sys : nsys - 1;
  double a[TS_MAXN], xs[TS_MAXN];
  for (int c = 0; c < TS_MAXN; ++c) a[c] = tl.lane < TS_MAXN ? A[((long long)s * TS_MAXN + tl.lane) * TS_MAXN + c] : 0.0;
  const double bb = tl.lane < TS_MAXN ? b[(long long)s * TS_MAXN + tl.lane] : 0.0;
  lu_rows_solve_pivot(tl, a, bb, xs, scr[threadIdx.x / LPE]);
  if (sys < nsys && tl.lane == 0) for (int i = 0; i < TS_MAXN; ++i) x[(long long)s * TS_MAXN + i] = xs[i];
}

extern "C" {

int tsim_variant_lu_solve(int device, int nsys, const double* A, const double* b, double* x) {
  if (!A || !b || !x || nsys < 1) return fail("tsim_debug_lu_solve: bad argument");
  const size_t na = (size_t)nsys * TS_MAXN * TS_MAXN * sizeof(double), nb = (size_t)nsys * TS_MAXN * sizeof(double);
  double *dA = 0, *db = 0, *dx = 0;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&dA, na);
  if (e == cudaSuccess) e = cudaMalloc(&db, nb);
  if (e == cudaSuccess) e = cudaMalloc(&dx, nb);
  if (e == cudaSuccess) e = cudaMemcpy(dA, A, na, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(db, b, nb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const int per = 32 / TS_MAXN;
    lu_debug_kernel<TS_MAXN><<<(nsys + per - 1) / per, 32>>>(dA, db, dx, nsys);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(x, dx, nb, cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(db); cudaFree(dx);
  if (e != cudaSuccess) return fail(std::string("tsim_debug_lu_solve: ") + cudaGetErrorString(e));
  return 0;
}

#ifdef TS_PROFILE
int tsim_debug_set_prof(void* p) { return cudaMemcpyToSymbol(g_prof, &p, sizeof(p)) != cudaSuccess; }
#endif

const char* tsim_last_error(void) { return g_err.c_str(); }

int tsim_scene_create(const int32_t* ibuf, int64_t n_int, const double* dbuf, int64_t n_dbl, int device,
                      tsim_scene** out) {
  if (!ibuf || !dbuf || !out) return fail("tsim_scene_create: null argument");
  if (n_int < TS_I_HEADER || ibuf[TS_I_MAGIC] != TS_MAGIC || ibuf[TS_I_VERSION] < TS_VERSION_MIN || ibuf[TS_I_VERSION] > TS_VERSION)
    return fail("tsim_scene_create: not a scene blob of this version");
  // lower the portable scene description to the kernel tables (fixed joints folded, culling radii, ...)
  KernelTables kt;
  const std::string err = lower_scene(ibuf, n_int, dbuf, n_dbl, kt);
  if (!err.empty()) return fail("tsim_scene_create: " + err);
  CK(cudaSetDevice(device));
  tsim_scene* s = new tsim_scene();
  struct Guard {                       // a failure below destroys what was built so far
    tsim_scene* s;
    ~Guard() { if (s) tsim_scene_destroy(s); }
  } guard = {s};
  s->device = device;
  s->d_ib = 0; s->d_db = 0;
  for (int i = 0; i < 7; ++i) s->ev[i] = 0;
  s->ni = (int)kt.ib.size();
  s->nd_all = (int)kt.db.size();
  s->nd = kt.ib[KI_D_MARKERS];       // doubles staged in shared memory: everything before the marker table
  s->lanes = TS_MAXN;        // one lane per reduced coordinate
  s->nmj = kt.ib[KI_NMJ];
  s->d_env_db = 0;
  s->env_B = 0;
  s->base_ib = kt.ib;
  for (int i = 0; i < 7; ++i) CK(cudaEventCreate(&s->ev[i]));
  for (int i = 0; i < TSIM_N_KERNELS; ++i) s->ran[i] = 0;
  s->opts[TSIM_OPT_LS_BATCH] = 1;
  s->opts[TSIM_OPT_MAX_NEWTON] = 0;
  s->opts[TSIM_OPT_VJP_PASS] = 1;
  s->opts[TSIM_OPT_TAC_PASS] = 1;
  s->opts[TSIM_OPT_TAPE_PASS] = 1;
  CK(cudaMalloc(&s->d_ib, sizeof(int) * s->ni));
  CK(cudaMalloc(&s->d_db, sizeof(double) * s->nd_all));
  CK(cudaMemcpy(s->d_ib, kt.ib.data(), sizeof(int) * s->ni, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s->d_db, kt.db.data(), sizeof(double) * s->nd_all, cudaMemcpyHostToDevice));
  const int n = ibuf[TS_I_NDOF_R];
  s->sizes[TSIM_NJ] = ibuf[TS_I_NJ];
  s->sizes[TSIM_NDOF_R] = n;
  s->sizes[TSIM_NDOF_M] = 6 * ibuf[TS_I_NJ];
  s->sizes[TSIM_NDOF_U] = ibuf[TS_I_NDOF_U];
  s->sizes[TSIM_NDOF_VAR] = 3 * ibuf[TS_I_NEE];
  s->sizes[TSIM_NDOF_TACTILE] = 3 * ibuf[TS_I_NMARKERS];
  s->sizes[TSIM_N_MARKERS] = ibuf[TS_I_NMARKERS];
  s->sizes[TSIM_TAPE_DOUBLES] = 3 * n * n + ibuf[TS_I_NDOF_U];
  s->sizes[TSIM_CMASK_WORDS] = kt.ib[KI_CMW];
  s->sizes[TSIM_INTEGRATOR] = kt.ib[KI_INTEGRATOR];
  guard.s = 0;
  *out = s;
  return 0;
}

void tsim_scene_destroy(tsim_scene* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaFree(s->d_ib);
  cudaFree(s->d_db);
  if (s->d_env_db) cudaFree(s->d_env_db);
  for (int i = 0; i < 7; ++i) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
  delete s;
}

int tsim_scene_sizes(const tsim_scene* s, int32_t* out) {
  if (!s || !out) return fail("tsim_scene_sizes: null argument");
  for (int i = 0; i < TSIM_N_SIZES; ++i) out[i] = s->sizes[i];
  return 0;
}

int tsim_scene_set_lanes(tsim_scene* s, int lanes) {
  if (!s) return fail("tsim_scene_set_lanes: null scene");
  if ((lanes != 8 && lanes != 16 && lanes != 32) || lanes < TS_MAXN)
    return fail("tsim_scene_set_lanes: lanes must be 8, 16 or 32 and at least the dof capacity of the scene's kernel variant");
  s->lanes = lanes;
  return 0;
}


int tsim_scene_set_option(tsim_scene* s, int key, int value) {
  if (!s) return fail("tsim_scene_set_option: null scene");
  if (key < 0 || key >= TSIM_N_OPTS) return fail("tsim_scene_set_option: unknown option");
  s->opts[key] = value;
  return 0;
}

// Per-environment parameters: B packed scenes of the handle's topology (the host edits a copy of the scene per
// environment: tactilesimulation_b200.scene.update_*), lowered here one by one; the integer tables must come out
// identical to the handle's, the double tables before the marker table are uploaded as [B][nd].  B = 0 drops them.
int tsim_scene_set_env_scenes(tsim_scene* s, int32_t B, const int32_t* ibufs, int64_t n_int, const double* dbufs, int64_t n_dbl) {
  if (!s) return fail("tsim_scene_set_env_scenes: null scene");
  CK(cudaSetDevice(s->device));
  if (s->d_env_db) { CK(cudaDeviceSynchronize()); CK(cudaFree(s->d_env_db)); s->d_env_db = 0; s->env_B = 0; }
  if (B <= 0) return 0;
  if (!ibufs || !dbufs) return fail("tsim_scene_set_env_scenes: null argument");
  std::vector<double> tables((size_t)B * s->nd);
  for (int e = 0; e < B; ++e) {
    KernelTables kt;
    const std::string err = lower_scene(ibufs + (size_t)e * n_int, n_int, dbufs + (size_t)e * n_dbl, n_dbl, kt);
    if (!err.empty()) return fail("tsim_scene_set_env_scenes: environment " + std::to_string(e) + ": " + err);
    if (kt.ib != s->base_ib || (int)kt.db.size() != s->nd_all)
      return fail("tsim_scene_set_env_scenes: environment " + std::to_string(e) +
                  " does not have the topology of the handle (per-environment parameters may change values, not counts)");
    memcpy(tables.data() + (size_t)e * s->nd, kt.db.data(), sizeof(double) * s->nd);
  }
  CK(cudaMalloc(&s->d_env_db, sizeof(double) * tables.size()));
  CK(cudaMemcpy(s->d_env_db, tables.data(), sizeof(double) * tables.size(), cudaMemcpyHostToDevice));
  s->env_B = B;
  return 0;
}

// (the public tsim_forward is tsim_forward_multistep(..., NULL, NULL, 0, ...): csrc/cabi.cpp)
int tsim_forward_multistep(const tsim_scene* s, int32_t B, int32_t T, double* q, double* qd, double* q_prev,
                           double* qd_prev, int32_t steps_done, const double* u, int64_t u_step_stride,
                           double* q_traj, double* qd_traj, double* var_out, const int32_t* var_row, double* tac_out,
                           const int32_t* tac_row, double* tape, int32_t* status, uint32_t* contact_masks,
                           int32_t* marker_body, void* stream) {
  if (!s) return fail("tsim_forward: null scene");
  if (B <= 0 || T < 0) return fail("tsim_forward: bad batch or step count");
  if (!q || !qd || !u) return fail("tsim_forward: q, qd and u are required");
  if (s->sizes[TSIM_INTEGRATOR] != TSIM_INT_BDF1 && tape)
    return fail("tsim_forward: the adjoint tape exists for BDF1 scenes only (as Simulation::backward of the reference)");
  if (steps_done < 0 || (steps_done > 0 && s->sizes[TSIM_INTEGRATOR] == TSIM_INT_BDF2 && (!q_prev || !qd_prev)))
    return fail("tsim_forward_multistep: continuing a BDF2 trajectory needs q_prev and qd_prev");
  if ((q_prev == 0) != (qd_prev == 0)) return fail("tsim_forward_multistep: q_prev and qd_prev go together");
  if (T == 0) return 0;
  CK(cudaSetDevice(s->device));
  FwdArgs a;
  a.B = B; a.T = T; a.q = q; a.qd = qd; a.u = u; a.u_stride = u_step_stride; a.q_traj = q_traj; a.qd_traj = qd_traj;
  a.var_out = var_out; a.var_row = var_row; a.tac_out = tac_out; a.tac_row = tac_row; a.tape = tape;
  a.status = status; a.cmask = contact_masks; a.marker_body = marker_body;
  a.ls_batch = s->opts[TSIM_OPT_LS_BATCH];
  a.max_newton = s->opts[TSIM_OPT_MAX_NEWTON];
  a.q_prev = q_prev; a.qd_prev = qd_prev; a.steps_done = steps_done;
  a.defer_tac = 0; a.tac_prezeroed = 0; a.work_counter = 0;
  TS_PE_CHECK("tsim_forward")
  a.env_db = s->d_env_db; a.env_stride = s->nd;
  const size_t smem = scene_smem(s);
  // the step loop adds the cooperative contact-point area of its block (8-lane tiles of variant 8)
  const size_t smem_fwd = s->lanes == 8 && coop_bytes<8>() ? ((smem + 15) & ~(size_t)15) + coop_bytes<8>() : smem;
  const long long threads = (long long)B * s->lanes;
  const int grid = (int)((threads + TS_BLOCK - 1) / TS_BLOCK);
  cudaStream_t st = (cudaStream_t)stream;
  // The tactile field is read out by a pass of its own over the recorded trajectory (tac_kernel) when the call covers
  // more than a few steps; scratch (trajectory if the caller keeps none, work counter) is stream-ordered.
  Scratch scratch(st);
  const bool big = T >= 4 && (long long)T * B < (1ll << 31) - 64;
  const bool tac_pass = tac_out && s->opts[TSIM_OPT_TAC_PASS] != 0 && big;
  // ... and so are the G0 / G1 / gain blocks of the tape (tape_kernel), which need the state at the start of the call
  const bool tape_pass = tape && q_traj && qd_traj && s->opts[TSIM_OPT_TAPE_PASS] != 0 && big;
  a.defer_g0 = 0; a.q_start = 0; a.qd_start = 0; a.tape_order = 0;
  if (tac_pass || tape_pass) {
    const size_t nvec = (size_t)T * B * s->sizes[TSIM_NDOF_R], nst = (size_t)B * s->sizes[TSIM_NDOF_R];
    const int own_traj = (q_traj ? 0 : 1) + (qd_traj ? 0 : 1);    // trajectories the caller does not keep: scratch
    const size_t norder = (((size_t)T * B * sizeof(int)) + 15) & ~(size_t)15;
    const size_t need = 16 + own_traj * nvec * sizeof(double) + (tape_pass ? 2 * nst * sizeof(double) + norder : 0);
    CK(scratch_alloc(&scratch.p, need, s->device, st));
    unsigned char* p = (unsigned char*)scratch.p;
    a.work_counter = (unsigned*)p;
    p += 16;
    if (!q_traj) { a.q_traj = (double*)p; p += nvec * sizeof(double); }
    if (!qd_traj) { a.qd_traj = (double*)p; p += nvec * sizeof(double); }
    CK(cudaMemsetAsync(a.work_counter, 0, 16, st));
    a.defer_tac = tac_pass ? 1 : 0;
    // identity row map: the whole [T,B,3M] field is this call's; a memset (issued with the pass, inside its timer)
    // runs at the HBM rate and the pass then only writes where a body can reach a pad (the sign of a zero is not
    // part of the contract)
    if (tac_pass && !tac_row) a.tac_prezeroed = 1;
    if (tape_pass) {
      double* qs = (double*)p;
      CK(cudaMemcpyAsync(qs, q, nst * sizeof(double), cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(qs + nst, qd, nst * sizeof(double), cudaMemcpyDeviceToDevice, st));
      a.q_start = qs; a.qd_start = qs + nst;
      a.tape_order = (int*)(qs + 2 * nst);
      a.defer_g0 = 1;
    }
  }
  s->ran[TSIM_K_FWD] = 1; s->ran[TSIM_K_TAPE] = tape_pass ? 1 : 0; s->ran[TSIM_K_TAC] = tac_pass ? 1 : 0;
  CK(cudaEventRecord(s->ev[0], st));
  TS_LAUNCH(fwd_kernel, grid, smem, smem_fwd, TS_BPS, s->d_ib, s->ni, s->d_db, s->nd, a);
  CK(cudaEventRecord(s->ev[1], st));
  int tgrid = 1;
  if (tac_pass || tape_pass) {
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, s->device));
    const long long want = ((long long)T * B * s->lanes + TS_BLOCK - 1) / TS_BLOCK;
    tgrid = (int)(want < (long long)nsm * TS_BPS ? want : (long long)nsm * TS_BPS);
  }
  if (tape_pass) {
    TS_LAUNCH(tape_kernel, tgrid, smem, smem, TS_BPS, s->d_ib, s->ni, s->d_db, s->nd, a);
  }
  CK(cudaEventRecord(s->ev[2], st));
  if (tac_pass) {
    if (a.tac_prezeroed) CK(cudaMemsetAsync(tac_out, 0, (size_t)T * B * s->sizes[TSIM_NDOF_TACTILE] * sizeof(double), st));
    TS_LAUNCH(tac_kernel, tgrid * TS_PASS_BPS / TS_BPS, smem, smem, TS_PASS_BPS, s->d_ib, s->ni, s->d_db, s->nd, a);
  }
  CK(cudaEventRecord(s->ev[3], st));
  return 0;                                  // (~Scratch returns the scratch to the pool, stream-ordered)
}

int tsim_readout(const tsim_scene* s, int32_t B, const double* q, const double* qd, double* var_out, double* tac_out,
                 int32_t* marker_body, uint32_t* contact_masks, void* stream) {
  if (!s) return fail("tsim_readout: null scene");
  if (B <= 0 || !q || !qd) return fail("tsim_readout: bad arguments");
  CK(cudaSetDevice(s->device));
  TS_PE_CHECK("tsim_readout")
  EnvTables et;
  et.env_db = s->d_env_db; et.env_stride = s->nd;
  const size_t smem = scene_smem(s);
  const long long threads = (long long)B * s->lanes;
  const int grid = (int)((threads + TS_BLOCK - 1) / TS_BLOCK);
  cudaStream_t st = (cudaStream_t)stream;
  TS_LAUNCH(readout_kernel, grid, smem, smem, TS_BPS, s->d_ib, s->ni, s->d_db, s->nd, B, q, qd, var_out, tac_out, marker_body, contact_masks, et);
  return 0;
}

int tsim_backward(const tsim_scene* s, int32_t B, int32_t T, const double* q_traj, const double* qd_traj,
                  const double* u, int64_t u_step_stride, const double* tape, const double* df_dq,
                  const int32_t* dq_row, const double* df_dvar, const int32_t* dvar_row, const double* df_dtac,
                  const int32_t* dtac_row, double* carry, double* df_du, double* df_dq0, double* df_dqdot0,
                  void* stream) {
  if (!s) return fail("tsim_backward: null scene");
  if (s->sizes[TSIM_INTEGRATOR] != TSIM_INT_BDF1)
    return fail("tsim_backward: the adjoint exists for BDF1 scenes only (as Simulation::backward of the reference)");
  if (B <= 0 || T <= 0) return fail("tsim_backward: bad batch or step count");
  if (!q_traj || !qd_traj || !u || !tape || !carry)
    return fail("tsim_backward: q_traj, qd_traj, u, tape and carry are required (run tsim_forward with a tape first)");
  CK(cudaSetDevice(s->device));
  BwdArgs a;
  a.B = B; a.T = T; a.q_traj = q_traj; a.qd_traj = qd_traj; a.u = u; a.u_stride = u_step_stride; a.tape = tape;
  a.df_dq = df_dq; a.dq_row = dq_row; a.df_dvar = df_dvar; a.dvar_row = dvar_row; a.df_dtac = df_dtac;
  a.dtac_row = dtac_row; a.carry = carry; a.df_du = df_du; a.df_dq0 = df_dq0; a.df_dqdot0 = df_dqdot0;
  a.vjp_y = 0; a.vjp_c = 0; a.work_counter = 0; a.vjp_list = 0;
  TS_PE_CHECK("tsim_backward")
  a.env_db = s->d_env_db; a.env_stride = s->nd;
  const size_t smem = scene_smem(s);
  const long long threads = (long long)B * s->lanes;
  const int grid = (int)((threads + TS_BLOCK - 1) / TS_BLOCK);
  cudaStream_t st = (cudaStream_t)stream;
  // Pass 1: readout pull-backs of all env-steps (balanced over the whole GPU); pass 2: the reverse sweep reads them.
  // Scratch: 2 x [T,B,n] doubles + the work counter, stream-ordered allocation (retained by the device's pool).
  CK(cudaEventRecord(s->ev[4], st));
  Scratch scratch(st);
  const bool split = (df_dvar || df_dtac) && s->opts[TSIM_OPT_VJP_PASS] != 0 && (long long)T * B < (1ll << 31) - 64;
  if (split) {
    const size_t nvec = (size_t)T * B * s->sizes[TSIM_NDOF_R];
    CK(scratch_alloc(&scratch.p, 2 * nvec * sizeof(double) + 16 + (size_t)T * B * sizeof(int), s->device, st));
    a.vjp_y = (double*)scratch.p;
    a.vjp_c = a.vjp_y + nvec;
    a.work_counter = (unsigned*)(a.vjp_c + nvec);
    a.vjp_list = (int*)(a.work_counter + 4);
    CK(cudaMemsetAsync(a.work_counter, 0, 16, st));
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, s->device));
    const long long want = ((long long)T * B * s->lanes + TS_BLOCK - 1) / TS_BLOCK;
    const int vgrid = (int)(want < (long long)nsm * TS_PASS_BPS ? want : (long long)nsm * TS_PASS_BPS);
    for (int phase = 0; phase < 2; ++phase) {
      TS_LAUNCH(vjp_kernel, vgrid, smem, smem, TS_PASS_BPS, s->d_ib, s->ni, s->d_db, s->nd, a, phase);
    }
  }
  s->ran[TSIM_K_VJP] = split ? 1 : 0; s->ran[TSIM_K_BWD] = 1;
  CK(cudaEventRecord(s->ev[5], st));
  TS_LAUNCH(bwd_kernel, grid, smem, smem, TS_BPS, s->d_ib, s->ni, s->d_db, s->nd, a);
  CK(cudaEventRecord(s->ev[6], st));
  return 0;
}

int tsim_scene_kernel_times(const tsim_scene* s, double* ms) {
  if (!s || !ms) return fail("tsim_scene_kernel_times: null argument");
  CK(cudaSetDevice(s->device));
  const int a0[TSIM_N_KERNELS] = {0, 1, 2, 4, 5}, a1[TSIM_N_KERNELS] = {1, 2, 3, 5, 6};
  for (int k = 0; k < TSIM_N_KERNELS; ++k) {
    ms[k] = -1.0;
    if (!s->ran[k]) continue;
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, s->ev[a0[k]], s->ev[a1[k]]));
    ms[k] = (double)t;
  }
  return 0;
}

}  // extern "C"
}  // namespace
