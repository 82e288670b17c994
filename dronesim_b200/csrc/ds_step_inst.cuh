// Host-side dispatcher of the ds_step_kernel instantiations.
//
// ds_step_kernel<INTEG, DW, NU6, WARPSYNC, MODE, FX, EXT> has ~150 instantiations of a 2,600-instruction kernel;
// they are split over twelve translation units (ds_step_inst.cu compiled with -DDS_INST_INTEG=0|1 -DDS_INST_MODE=0|1|2
// -DDS_INST_NU6=0|1)
// that nvcc builds in parallel.  ds_api.cu only sees the declarations below.
#pragma once
#include <cuda_runtime.h>

#include "ds_device.cuh"

// one translation unit each: ds_launch_step_<q|r><mode>_<nu6>
#define DS_DECL(n) void n(int dw, bool warpsync, const DsArgs& a, const DsTypeDev* homo, int grid, cudaStream_t st)
DS_DECL(ds_launch_step_q0_0); DS_DECL(ds_launch_step_q0_1); DS_DECL(ds_launch_step_q1_0); DS_DECL(ds_launch_step_q1_1);
DS_DECL(ds_launch_step_q2_0); DS_DECL(ds_launch_step_q2_1);
DS_DECL(ds_launch_step_r0_0); DS_DECL(ds_launch_step_r0_1); DS_DECL(ds_launch_step_r1_0); DS_DECL(ds_launch_step_r1_1);
DS_DECL(ds_launch_step_r2_0); DS_DECL(ds_launch_step_r2_1);
#undef DS_DECL

// mode: 0 fused physics-then-control, 1 physics only, 2 fused control-then-physics
// homo: host copy of the single type's table when every slot flies the same type (nullptr otherwise)
static inline void ds_launch_step(int integ, int mode, int dw, bool nu6, bool warpsync, const DsArgs& a, const DsTypeDev* homo,
                                  int grid, cudaStream_t st) {
  typedef void (*fn_t)(int, bool, const DsArgs&, const DsTypeDev*, int, cudaStream_t);
  static const fn_t table[2][3][2] = {
      {{ds_launch_step_q0_0, ds_launch_step_q0_1}, {ds_launch_step_q1_0, ds_launch_step_q1_1}, {ds_launch_step_q2_0, ds_launch_step_q2_1}},
      {{ds_launch_step_r0_0, ds_launch_step_r0_1}, {ds_launch_step_r1_0, ds_launch_step_r1_1}, {ds_launch_step_r2_0, ds_launch_step_r2_1}}};
  table[integ ? 1 : 0][mode][nu6 ? 1 : 0](dw, warpsync, a, homo, grid, st);
}
