// Instantiations of ds_step_kernel for ONE (integrator, mode, rotor-count) triple: compile with
//   -DDS_INST_INTEG=0|1 (QUAT | RPY)  -DDS_INST_MODE=0|1|2 (fused physics-then-control | physics only | fused control-then-physics)
//   -DDS_INST_NU6=0|1 (all types have <= 4 rotors | some type has 6).
#include "ds_kernels.cuh"
#include "ds_step_inst.cuh"

#ifndef DS_INST_INTEG
#error "compile ds_step_inst.cu with -DDS_INST_INTEG=0|1 -DDS_INST_MODE=0|1|2 -DDS_INST_NU6=0|1"
#endif

// two tile stages of dynamic shared memory (> 48 KB: opt in once per instantiation and device).  `grid` arrives as the
// DS_MIN_CTAS-per-SM cap; variants whose registers and shared memory let more CTAs share an SM (the single-vehicle-env
// variants: ~80 registers, no downwash snapshot, no type table in shared memory -> 5 CTAs) get a grid to match, so that a
// fifth tile per SM is in flight on the HBM-bound K = 2 workloads.
template <class ARGS, void (*KERNEL)(const ARGS)>
static void launch5(const ARGS& args, int n_tiles, int grid, cudaStream_t st) {
  constexpr int kSmem = 2 * ds_stage_bytes<DS_INST_MODE>();
  static int occ[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!occ[dev]) {
    cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(KERNEL, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, KERNEL, DS_TILE, kSmem) != cudaSuccess || n < DS_MIN_CTAS) n = DS_MIN_CTAS;
    occ[dev] = n;
  }
  if (occ[dev] > DS_MIN_CTAS && grid < n_tiles) {
    grid = grid / DS_MIN_CTAS * occ[dev];
    if (grid > n_tiles) grid = n_tiles;
  }
  KERNEL<<<grid, DS_TILE, kSmem, st>>>(args);
}

// homo: the swarm's single airframe type (its table rides in the kernel parameters), nullptr for mixed swarms.  The
// homogeneous variants exist for the quaternion integrator without extensions (the throughput configurations).
template <int DW, bool NU6, bool WS, int FX, bool EXT, bool RC>
static void launch4b(const DsArgs& a, const DsTypeDev* homo, int grid, cudaStream_t st) {
#if DS_INST_INTEG == 0
  if (!EXT && homo) {
    DsArgsH ah;
    ah.a = a;
    ah.tp = *homo;
    launch5<DsArgsH, ds_step_kernel<DS_INST_INTEG, DW, NU6, WS, DS_INST_MODE, FX, false, true, RC>>(ah, a.n_tiles, grid, st);
    return;
  }
#endif
  launch5<DsArgs, ds_step_kernel<DS_INST_INTEG, DW, NU6, WS, DS_INST_MODE, FX, EXT, false, RC>>(a, a.n_tiles, grid, st);
}

// RC: some type has a centre-of-mass offset (quaternion integrator only; the extension variants always carry it)
template <int DW, bool NU6, bool WS, int FX, bool EXT>
static void launch4(const DsArgs& a, const DsTypeDev* homo, int grid, cudaStream_t st) {
#if DS_INST_INTEG == 0
  if (EXT || a.rc_kind != 0) launch4b<DW, NU6, WS, FX, EXT, true>(a, homo, grid, st);
  else launch4b<DW, NU6, WS, FX, false, false>(a, homo, grid, st);
#else
  launch4b<DW, NU6, WS, FX, EXT, false>(a, homo, grid, st);
#endif
}

template <int DW, bool NU6, int FX, bool EXT>
static void launch3(bool warpsync, const DsArgs& a, const DsTypeDev* homo, int grid, cudaStream_t st) {
  // warp-level sync of the downwash snapshot needs every env inside one warp: D | 32 (always true of the D = 16 variant)
  if (DW == 2 || warpsync) launch4<DW, NU6, true, FX, EXT>(a, homo, grid, st);
  else launch4<(DW == 2 ? 1 : DW), NU6, false, FX, EXT>(a, homo, grid, st);
}

template <int DW, bool NU6>
static void launch2(bool warpsync, const DsArgs& a, const DsTypeDev* homo, int grid, cudaStream_t st) {
  // FX: ground effect + drag resolved at compile time for the all-add-ons configuration (straight-line substep body),
  // run-time flags otherwise
#if DS_INST_INTEG == 0
  // EXT: motor model / angular-acceleration filter (extensions beyond the reference; quaternion integrator only)
  if (a.ext) { launch3<DW, NU6, -1, true>(warpsync, a, homo, grid, st); return; }
#endif
  // (the ground plane is a run-time flag as well: it rides in the FX = -1 variants only)
  if (DW != 0 && (a.flags & 3u) == 3u && !(a.flags & 64u)) { launch3<DW, NU6, 3, false>(warpsync, a, homo, grid, st); return; }
#if DS_INST_INTEG == 0
  // single-vehicle envs (BASELINE configs[1] / [2]): no add-ons, or ground effect + drag, at compile time too
  if (DW == 0 && !(a.flags & 64u)) {
    if ((a.flags & 3u) == 0u) { launch3<0, NU6, 0, false>(warpsync, a, homo, grid, st); return; }
    if ((a.flags & 3u) == 3u) { launch3<0, NU6, 3, false>(warpsync, a, homo, grid, st); return; }
  }
#endif
  launch3<DW, NU6, -1, false>(warpsync, a, homo, grid, st);
}

#define DS_CONCAT3_(a, b, c) a##b##_##c
#define DS_CONCAT3(a, b, c) DS_CONCAT3_(a, b, c)
#if DS_INST_INTEG == 0
#define DS_INST_NAME DS_CONCAT3(ds_launch_step_q, DS_INST_MODE, DS_INST_NU6)
#else
#define DS_INST_NAME DS_CONCAT3(ds_launch_step_r, DS_INST_MODE, DS_INST_NU6)
#endif

void DS_INST_NAME(int dw, bool warpsync, const DsArgs& a, const DsTypeDev* homo, int grid, cudaStream_t st) {
  constexpr bool NU6 = DS_INST_NU6 != 0;
  if (dw == 2) launch2<2, NU6>(warpsync, a, homo, grid, st);
  else if (dw == 1) launch2<1, NU6>(warpsync, a, homo, grid, st);
  else launch2<0, NU6>(warpsync, a, homo, grid, st);
}
