// Instantiations of ds_step_kernel for ONE (integrator, mode, rotor-count) triple: compile with
//   -DDS_INST_INTEG=0|1 (QUAT | RPY)  -DDS_INST_MODE=0|1|2 (fused physics-then-control | physics only | fused control-then-physics)
//   -DDS_INST_NU6=0|1 (all types have <= 4 rotors | some type has 6).
#include "ds_kernels.cuh"
#include "ds_step_inst.cuh"

#ifndef DS_INST_INTEG
#error "compile ds_step_inst.cu with -DDS_INST_INTEG=0|1 -DDS_INST_MODE=0|1|2 -DDS_INST_NU6=0|1"
#endif

// two tile stages of dynamic shared memory (> 48 KB: opt in once per instantiation and device)
template <void (*KERNEL)(const DsArgs)>
static void launch4(const DsArgs& a, int grid, cudaStream_t st) {
  constexpr int kSmem = 2 * ds_stage_bytes<DS_INST_MODE>();
  static bool opted[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !opted[dev]) {
    cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    opted[dev] = true;
  }
  KERNEL<<<grid, DS_TILE, kSmem, st>>>(a);
}

template <int DW, bool NU6, int FX, bool EXT>
static void launch3(bool warpsync, const DsArgs& a, int grid, cudaStream_t st) {
  // warp-level sync of the downwash snapshot needs every env inside one warp: D | 32 (always true of the D = 16 variant)
  if (DW == 2 || warpsync) launch4<ds_step_kernel<DS_INST_INTEG, DW, NU6, true, DS_INST_MODE, FX, EXT>>(a, grid, st);
  else launch4<ds_step_kernel<DS_INST_INTEG, (DW == 2 ? 1 : DW), NU6, false, DS_INST_MODE, FX, EXT>>(a, grid, st);
}

template <int DW, bool NU6>
static void launch2(bool warpsync, const DsArgs& a, int grid, cudaStream_t st) {
  // FX: ground effect + drag resolved at compile time for the all-add-ons configuration (straight-line substep body),
  // run-time flags otherwise
#if DS_INST_INTEG == 0
  // EXT: motor model / angular-acceleration filter (extensions beyond the reference; quaternion integrator only)
  if (a.ext) { launch3<DW, NU6, -1, true>(warpsync, a, grid, st); return; }
#endif
  if (DW != 0 && (a.flags & 3u) == 3u) launch3<DW, NU6, 3, false>(warpsync, a, grid, st);
  else launch3<DW, NU6, -1, false>(warpsync, a, grid, st);
}

#define DS_CONCAT3_(a, b, c) a##b##_##c
#define DS_CONCAT3(a, b, c) DS_CONCAT3_(a, b, c)
#if DS_INST_INTEG == 0
#define DS_INST_NAME DS_CONCAT3(ds_launch_step_q, DS_INST_MODE, DS_INST_NU6)
#else
#define DS_INST_NAME DS_CONCAT3(ds_launch_step_r, DS_INST_MODE, DS_INST_NU6)
#endif

void DS_INST_NAME(int dw, bool warpsync, const DsArgs& a, int grid, cudaStream_t st) {
  constexpr bool NU6 = DS_INST_NU6 != 0;
  if (dw == 2) launch2<2, NU6>(warpsync, a, grid, st);
  else if (dw == 1) launch2<1, NU6>(warpsync, a, grid, st);
  else launch2<0, NU6>(warpsync, a, grid, st);
}
