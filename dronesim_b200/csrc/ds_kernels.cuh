// Kernels of the dronesim_b200 core.
//
//  ds_step_kernel<INTEG, DW, NU6, WARPSYNC, MODE>
//      MODE 0 (fused): K physics substeps + one INDI evaluation per vehicle  (examples/fly_INDI.py:217-245)
//      MODE 1 (physics only): BaseAviary.step with an external action        (BaseAviary.py:428-555)
//  ds_control_kernel<NU6>      INDIControl.computeControl on resident or external state
//  ds_obs_kernel               CtrlAviary._computeObs: state vector + adjacency bitmask + done / reward
//  ds_reset_kernel             BaseAviary._housekeeping + INDIControl.reset
//
// Grid: persistent CTAs looping over tiles of tile_v vehicles (whole environments per tile, so the
// downwash neighbour exchange never leaves the CTA); grid size = min(tiles, SMs x resident CTAs).
//
// Thread -> vehicle map: vehicles are env-major in memory (v = env * D + slot) and thread t of a tile
// handles local vehicle t, so a warp reads 512 contiguous bytes of every state array and (for D | 32)
// holds whole envs: the downwash snapshot is exchanged with __syncwarp only.  Sorting the threads of a
// tile by airframe class (warp-uniform control law / rotor count) was measured and rejected: it removed
// 2.6 % of the issued instructions (divergence only idles 4.5 % of the lanes in the 8 quad + 8 hexa env)
// but needed a CTA-wide barrier per substep and ran 7 % slower (profiles/r01_notes.md).
#pragma once
#include "ds_control.cuh"
#include "ds_device.cuh"
#include "ds_physics.cuh"

__device__ __forceinline__ void ds_load_types(const DsArgs& a, DsTypeDev* sh_types) {
  const int words = a.n_types * (int)(sizeof(DsTypeDev) / 16);
  const float4* src = reinterpret_cast<const float4*>(a.types);
  float4* dst = reinterpret_cast<float4*>(sh_types);
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

__device__ __forceinline__ float ds_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ds_warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Rollout statistics: per-thread accumulators live in SHARED memory (one column per thread) so they cost no
// registers across the persistent tile loop; one warp-shuffle reduction + atomics per CTA at the end.
enum { ST_NCTRL = 0, ST_ERR2, ST_SAT, ST_WLS_SLOW, ST_WLS_FAIL, ST_NONFINITE, ST_MINZ, ST_DONE, ST_COUNT };

__device__ __forceinline__ void ds_flush_stats(const float* sh_stat, double* stats) {
  const int t = threadIdx.x;
  float v[ST_COUNT];
#pragma unroll
  for (int i = 0; i < ST_COUNT; ++i) v[i] = sh_stat[i * DS_TILE + t];
#pragma unroll
  for (int i = 0; i < ST_COUNT; ++i) v[i] = (i == ST_MINZ) ? ds_warp_min(v[i]) : ds_warp_sum(v[i]);
  if ((t & 31) == 0) {
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i)
      if (i != ST_MINZ && v[i] != 0.f) atomicAdd(stats + i, (double)v[i]);
    // min altitude: doubles order like their bit patterns for non-negative values only, so use a CAS loop
    unsigned long long* p = reinterpret_cast<unsigned long long*>(stats + ST_MINZ);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) > (double)v[ST_MINZ]) {
      unsigned long long assumed = old;
      old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong((double)v[ST_MINZ]));
      if (old == assumed) break;
    }
  }
}

// Bulk L2 prefetch (TMA unit, no shared-memory staging): one instruction pulls `bytes` (multiple of 16) of a
// contiguous slab into L2.  One thread per CTA issues these for the CTA's NEXT tile at the start of the
// current one, so the DRAM reads of tile i+1 overlap the K substeps of tile i and the tile-start / control-
// phase loads become L2 hits instead of exposed DRAM latency.
__device__ __forceinline__ void ds_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <bool NU6, int MODE>
__device__ __forceinline__ void ds_prefetch_tile(const DsArgs& a, int tile) {
  const long long v0 = (long long)tile * a.tile_v;
  const int cnt = (int)min((long long)DS_TILE, (long long)a.n - v0);  // vehicles of the tile that exist
  if (cnt <= 0) return;
  const uint32_t b16 = (uint32_t)cnt * 16u;
  ds_prefetch_l2(a.s_pos + v0, b16); ds_prefetch_l2(a.s_quat + v0, b16);
  ds_prefetch_l2(a.s_vel + v0, b16); ds_prefetch_l2(a.s_om + v0, b16);
  if (MODE == 0) {
    ds_prefetch_l2(a.s_lv + v0, b16); ds_prefetch_l2(a.s_lr + v0, b16); ds_prefetch_l2(a.s_c0 + v0, b16);
    if (NU6) {  // float2 array: keep the slab 16-byte aligned and inside the allocation
      const long long v0e = v0 & ~1LL;
      const uint32_t b8 = (uint32_t)(((cnt + (int)(v0 - v0e)) * 8) & ~15);
      if (b8) ds_prefetch_l2(a.s_c1 + v0e, b8);
    }
    if (a.tmode == 0) {
      ds_prefetch_l2(a.t_pos + v0, b16);
      if (a.t_vel) ds_prefetch_l2(a.t_vel + v0, b16);
      if (a.t_acc) ds_prefetch_l2(a.t_acc + v0, b16);
    } else if (a.tmode == 2) {
      ds_prefetch_l2(a.t_vel + v0, b16);
    } else if (a.t_off) {
      ds_prefetch_l2(a.t_off + v0, b16);
    }
  }
}

__device__ __forceinline__ CtrlTarget ds_fetch_target(const DsArgs& a, const DsTypeDev& tp, const CtrlState& cs, int v,
                                                      int& wp) {
  CtrlTarget t;
  if (a.tmode == 2) {  // VelocityAviary._preprocessAction (VelocityAviary.py:236-257)
    float4 q = __ldg(a.t_vel + v);
    float n2 = q.x * q.x + q.y * q.y + q.z * q.z;
    float k = (n2 != 0.f) ? tp.speed_limit * fabsf(q.w) / sqrtf(n2) : 0.f;
    float roll, pitch, yaw;
    ds_euler(cs.qx, cs.qy, cs.qz, cs.qw, roll, pitch, yaw);
    t.x = cs.px; t.y = cs.py; t.z = cs.pz; t.yaw = yaw;  // hold position and yaw (state[0:3], state[9])
    t.vx = k * q.x; t.vy = k * q.y; t.vz = k * q.z;
    t.ax = t.ay = t.az = 0.f;
  } else if (a.tmode == 0) {
    float4 p = __ldg(a.t_pos + v);
    t.x = p.x; t.y = p.y; t.z = p.z; t.yaw = p.w;
    t.vx = t.vy = t.vz = t.ax = t.ay = t.az = 0.f;
    if (a.t_vel) { float4 q = __ldg(a.t_vel + v); t.vx = q.x; t.vy = q.y; t.vz = q.z; }
    if (a.t_acc) { float4 q = __ldg(a.t_acc + v); t.ax = q.x; t.ay = q.y; t.az = q.z; }
  } else {
    const float4* row = a.t_table + 3 * wp;
    float4 p = __ldg(row), q = __ldg(row + 1), r = __ldg(row + 2);
    t.x = p.x; t.y = p.y; t.z = p.z; t.yaw = p.w;
    t.vx = q.x; t.vy = q.y; t.vz = q.z; t.ax = r.x; t.ay = r.y; t.az = r.z;
    if (a.t_off) { float4 o = __ldg(a.t_off + v); t.x += o.x; t.y += o.y; t.z += o.z; }
    if (a.advance_wp) wp = (wp < a.num_wp - 1) ? wp + 1 : 0;  // fly_INDI.py:242-245
  }
  return t;
}

template <int INTEG, int DW, bool NU6, bool WARPSYNC, int MODE>
__global__ void __launch_bounds__(DS_TILE, DS_MIN_CTAS) ds_step_kernel(const DsArgs a) {
  __shared__ __align__(16) DsTypeDev sh_types[DS_MAX_TYPES_DEV];
  __shared__ uint8_t sh_slot_type[32];
  __shared__ __align__(16) float4 sh_pos[DW ? 2 * DS_DW_BUF : 1];
  __shared__ float sh_stat[ST_COUNT * DS_TILE];
  ds_load_types(a, sh_types);
  if (threadIdx.x < 32) sh_slot_type[threadIdx.x] = (threadIdx.x < a.D) ? a.slot_type[threadIdx.x] : 0;
  const bool stats_on = (a.flags & 8u) != 0;
  if (stats_on) {
#pragma unroll
    for (int i = 0; i < ST_COUNT; ++i) sh_stat[i * DS_TILE + threadIdx.x] = (i == ST_MINZ) ? 3.0e38f : 0.f;
  }
  __syncthreads();

  constexpr int NU = NU6 ? 6 : 4;
  const int tid = threadIdx.x;
  const bool lane_ok = tid < a.tile_v;
  const int lv = lane_ok ? tid : 0;  // idle lanes shadow local vehicle 0 (no stores) so barriers stay uniform
  const int slot = lv % a.D;
  // row of the env's slot 0 in the downwash snapshot: D + 1 padded rows per env, or a 32-row block (symmetric variant)
  const int env_row0 = (DW == 2) ? (lv / 16) * DS_DW_SYM_ROWS : (lv / a.D) * (a.D + DS_DW_PAD);
  const int my_row = env_row0 + slot;
  const int type_id = sh_slot_type[slot];
  const DsTypeDev& tp = sh_types[type_id];

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int v = tile * a.tile_v + lv;
    const bool valid = lane_ok && v < a.n;
    const int vv = valid ? v : 0;
    if (tid == 0 && tile + (int)gridDim.x < a.n_tiles) ds_prefetch_tile<NU6, MODE>(a, tile + gridDim.x);

    // ---- physics inputs.  The controller memory (last_vel, last_rates, cmd) is loaded only when the control
    // law runs, so that it does not occupy registers during the K substeps.
    const float4 P = a.s_pos[vv], Q = a.s_quat[vv], V = a.s_vel[vv], W = a.s_om[vv];
    PhysState s = {P.x, P.y, P.z, Q.x, Q.y, Q.z, Q.w, V.x, V.y, V.z, W.x, W.y, W.z};
    float prev_rpm_sum = V.w;
    float lthrust = P.w;
    int wp = __float_as_int(W.w);

    CtrlMem m;
    CtrlOut o = {0.f, 0.f, 0.f, 0.f, 0, 0};
    float perr = 0.f;
    uint32_t done_bits = 0;
    auto control = [&]() {  // INDIControl.computeControl on the resident state
      const float4 LV = a.s_lv[vv], LR = a.s_lr[vv], C0 = a.s_c0[vv];
      m.lvx = LV.x; m.lvy = LV.y; m.lvz = LV.z; m.lrx = LR.x; m.lry = LR.y; m.lrz = LR.z; m.lthrust = lthrust;
      m.cmd[0] = C0.x; m.cmd[1] = C0.y; m.cmd[2] = C0.z; m.cmd[3] = C0.w;
      if (NU6) { const float2 C1 = a.s_c1[vv]; m.cmd[4] = C1.x; m.cmd[5] = C1.y; } else { m.cmd[4] = m.cmd[5] = 0.f; }
      done_bits = __float_as_uint(LV.w);
      CtrlState cs = {s.px, s.py, s.pz, s.qx, s.qy, s.qz, s.qw, s.vx, s.vy, s.vz, s.wx, s.wy, s.wz};
      CtrlTarget t = ds_fetch_target(a, tp, cs, vv, wp);
      ds_indi_control<NU6>(tp, a.wls, type_id, cs, t, a.inv_ctrl_dt, m, o, false);
      perr = sqrtf(o.pex * o.pex + o.pey * o.pey + o.pez * o.pez);
      lthrust = m.lthrust;
    };

    // ---- the action the physics applies
    float act[6];
    if (MODE == 0 && a.order == 1) {  // VelocityAviary order: control, then physics with the new command
      control();
#pragma unroll
      for (int i = 0; i < NU; ++i) act[i] = m.cmd[i];
    } else if (MODE == 1) {  // external action, clipped (CtrlAviary.py:258-263)
      const float* ea = a.ext_action + (size_t)vv * 6;
#pragma unroll
      for (int i = 0; i < NU; ++i) act[i] = ds_clampf(ea[i], tp.rotor[i].pmin, tp.rotor[i].pmax);
    } else if (a.use_act) {  // first step after reset: the caller's initial action (fly_INDI.py:214)
      const float4 A0 = a.s_a0[vv];
      act[0] = A0.x; act[1] = A0.y; act[2] = A0.z; act[3] = A0.w;
      if (NU6) { const float2 A1 = a.s_a1[vv]; act[4] = A1.x; act[5] = A1.y; }
#pragma unroll
      for (int i = 0; i < NU; ++i) act[i] = ds_clampf(act[i], tp.rotor[i].pmin, tp.rotor[i].pmax);
    } else {  // the resident controller command, already clipped by the controller (INDIControl.py:487)
      const float4 C0 = a.s_c0[vv];
      act[0] = C0.x; act[1] = C0.y; act[2] = C0.z; act[3] = C0.w;
      if (NU6) { const float2 C1 = a.s_c1[vv]; act[4] = C1.x; act[5] = C1.y; }
    }

    ds_physics<INTEG, DW, NU6, WARPSYNC>(a, tp, env_row0, my_row, sh_pos, act, s, prev_rpm_sum);
    if (MODE == 0 && a.order == 0) control();
    if (MODE == 1) done_bits = __float_as_uint(a.s_lv[vv].w);

    // ---- done predicate on the fresh state (fly_INDI_TrajectoryTrack.py:249-250)
    if (a.goal_en) {
      float dx = s.px - a.goal_x, dy = s.py - a.goal_y, dz = s.pz - a.goal_z;
      if (sqrtf(dx * dx + dy * dy + dz * dz) < a.goal_r) done_bits |= 1u;
    }
    if (a.floor_en && s.pz < a.z_min) done_bits |= 2u;
    if (a.time_hit) done_bits |= 4u;

    if (valid) {
      a.s_pos[v] = make_float4(s.px, s.py, s.pz, lthrust);
      a.s_quat[v] = make_float4(s.qx, s.qy, s.qz, s.qw);
      a.s_vel[v] = make_float4(s.vx, s.vy, s.vz, prev_rpm_sum);
      a.s_om[v] = make_float4(s.wx, s.wy, s.wz, __int_as_float(wp));
      if (MODE == 0) {
        a.s_lv[v] = make_float4(m.lvx, m.lvy, m.lvz, __uint_as_float(done_bits));
        a.s_lr[v] = make_float4(m.lrx, m.lry, m.lrz, perr);
        a.s_c0[v] = make_float4(m.cmd[0], m.cmd[1], m.cmd[2], m.cmd[3]);
        if (NU6) a.s_c1[v] = make_float2(m.cmd[4], m.cmd[5]);
      } else {
        reinterpret_cast<float*>(a.s_lv + v)[3] = __uint_as_float(done_bits);
      }
      if (a.store_act) {
        a.s_a0[v] = make_float4(act[0], act[1], act[2], act[3]);
        if (NU6) a.s_a1[v] = make_float2(act[4], act[5]);
      }
      if (stats_on) {
        float* st = sh_stat + tid;
        if (MODE == 0) {
          st[ST_NCTRL * DS_TILE] += 1.f;
          st[ST_ERR2 * DS_TILE] += perr * perr;
          if (o.sat) st[ST_SAT * DS_TILE] += (float)o.sat;
          if (o.wls_iter != 1 && o.wls_iter != 0) st[ST_WLS_SLOW * DS_TILE] += 1.f;
          if (o.wls_iter < 0) st[ST_WLS_FAIL * DS_TILE] += 1.f;
        }
        const bool fin = isfinite(s.px) && isfinite(s.py) && isfinite(s.pz) && isfinite(s.qw) && isfinite(s.vx) && isfinite(s.wx);
        if (!fin) st[ST_NONFINITE * DS_TILE] += 1.f;
        st[ST_MINZ * DS_TILE] = fminf(st[ST_MINZ * DS_TILE], s.pz);
        if (done_bits) st[ST_DONE * DS_TILE] += 1.f;
      }
    }
  }
  if (stats_on) ds_flush_stats(sh_stat, a.stats);
}

// ---------------------------------------------------------------------------------------------
// control only (resident or external state; MODE 1 = rate/thrust entry of RPYTAviary)
// ---------------------------------------------------------------------------------------------
template <bool NU6, int MODE>
__global__ void __launch_bounds__(DS_TILE) ds_control_kernel(const DsArgs a) {
  __shared__ __align__(16) DsTypeDev sh_types[DS_MAX_TYPES_DEV];
  __shared__ uint8_t sh_slot_type[32];
  ds_load_types(a, sh_types);
  if (threadIdx.x < 32) sh_slot_type[threadIdx.x] = (threadIdx.x < a.D) ? a.slot_type[threadIdx.x] : 0;
  __syncthreads();
  constexpr int NU = NU6 ? 6 : 4;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < a.n; v += gridDim.x * blockDim.x) {
    const int type_id = sh_slot_type[v % a.D];
    const DsTypeDev& tp = sh_types[type_id];
    CtrlState cs;
    float4 P = a.s_pos[v], W = a.s_om[v];
    if (a.ext_state) {  // BaseControl.computeControlFromState slicing (BaseControl.py:92-103)
      const float* e = a.ext_state + (size_t)v * 22;
      cs.px = e[0]; cs.py = e[1]; cs.pz = e[2];
      cs.qx = e[3]; cs.qy = e[4]; cs.qz = e[5]; cs.qw = e[6];
      cs.vx = e[10]; cs.vy = e[11]; cs.vz = e[12];
      float d = cs.qx * cs.qx + cs.qy * cs.qy + cs.qz * cs.qz + cs.qw * cs.qw;
      Mat3 R = ds_rot(cs.qx, cs.qy, cs.qz, cs.qw, 2.0f / d);  // world -> body rates (INDIControl.py:428-430)
      cs.wx = R.m00 * e[13] + R.m10 * e[14] + R.m20 * e[15];
      cs.wy = R.m01 * e[13] + R.m11 * e[14] + R.m21 * e[15];
      cs.wz = R.m02 * e[13] + R.m12 * e[14] + R.m22 * e[15];
    } else {
      float4 Q = a.s_quat[v], V = a.s_vel[v];
      cs.px = P.x; cs.py = P.y; cs.pz = P.z; cs.qx = Q.x; cs.qy = Q.y; cs.qz = Q.z; cs.qw = Q.w;
      cs.vx = V.x; cs.vy = V.y; cs.vz = V.z; cs.wx = W.x; cs.wy = W.y; cs.wz = W.z;
    }
    float4 LV = a.s_lv[v], LR = a.s_lr[v], C0 = a.s_c0[v];
    CtrlMem m;
    m.lvx = LV.x; m.lvy = LV.y; m.lvz = LV.z; m.lrx = LR.x; m.lry = LR.y; m.lrz = LR.z; m.lthrust = P.w;
    m.cmd[0] = C0.x; m.cmd[1] = C0.y; m.cmd[2] = C0.z; m.cmd[3] = C0.w;
    if (NU6) { float2 C1 = a.s_c1[v]; m.cmd[4] = C1.x; m.cmd[5] = C1.y; } else { m.cmd[4] = m.cmd[5] = 0.f; }
    CtrlOut o = {0.f, 0.f, 0.f, 0.f, 0, 0};
    int wp = __float_as_int(W.w);
    if (MODE == 0) {
      CtrlTarget t = ds_fetch_target(a, tp, cs, v, wp);
      ds_indi_control<NU6>(tp, a.wls, type_id, cs, t, a.inv_ctrl_dt, m, o, true);
    } else {  // INDIControl._INDIRateControl (INDIControl.py:413-490)
      float4 rt = a.rate_thrust[v];
      float nu[4];
      ds_rate_loop(tp, cs, a.inv_ctrl_dt, rt.x, rt.y, rt.z, m, nu);
      nu[3] = rt.w - m.lthrust;
      m.lthrust = rt.w;
      ds_allocate_quad<NU6>(tp, nu, m, o);
    }
    a.s_pos[v] = make_float4(P.x, P.y, P.z, m.lthrust);
    a.s_om[v] = make_float4(W.x, W.y, W.z, __int_as_float(wp));
    a.s_lv[v] = make_float4(m.lvx, m.lvy, m.lvz, LV.w);
    a.s_lr[v] = make_float4(m.lrx, m.lry, m.lrz, sqrtf(o.pex * o.pex + o.pey * o.pey + o.pez * o.pez));
    a.s_c0[v] = make_float4(m.cmd[0], m.cmd[1], m.cmd[2], m.cmd[3]);
    if (NU6) a.s_c1[v] = make_float2(m.cmd[4], m.cmd[5]);
    if (a.cmd_out) {
      float* c = a.cmd_out + (size_t)v * 6;
#pragma unroll
      for (int i = 0; i < 6; ++i) c[i] = (i < NU) ? m.cmd[i] : 0.f;
    }
    if (a.pos_e_out) { float* e = a.pos_e_out + (size_t)v * 3; e[0] = o.pex; e[1] = o.pey; e[2] = o.pez; }
    if (a.yaw_err_out) a.yaw_err_out[v] = o.yaw_err;
  }
}

// ---------------------------------------------------------------------------------------------
// observation (CtrlAviary._computeObs, CtrlAviary.py:212-232)
// ---------------------------------------------------------------------------------------------
struct DsObsArgs {
  const float4 *s_pos, *s_quat, *s_vel, *s_om, *s_lv, *s_c0;
  const float2* s_c1;
  const uint8_t* slot_type;
  const DsTypeDev* types;
  float* obs;
  uint32_t* neighbors;
  uint8_t* done_env;
  float* reward_env;
  int n, D, nu6;
  float radius;
};

__global__ void __launch_bounds__(256) ds_obs_kernel(const DsObsArgs a) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < a.n; v += gridDim.x * blockDim.x) {
    const int slot = v % a.D, env0 = v - slot;
    float4 P = a.s_pos[v];
    if (a.obs) {
      float4 Q = a.s_quat[v], V = a.s_vel[v], W = a.s_om[v], C0 = a.s_c0[v];
      float2 C1 = a.nu6 ? a.s_c1[v] : make_float2(0.f, 0.f);
      float roll, pitch, yaw;
      ds_euler(Q.x, Q.y, Q.z, Q.w, roll, pitch, yaw);  // BaseAviary.py:729
      float d = Q.x * Q.x + Q.y * Q.y + Q.z * Q.z + Q.w * Q.w;
      Mat3 R = ds_rot(Q.x, Q.y, Q.z, Q.w, 2.0f / d);
      float* o = a.obs + (size_t)v * 22;  // BaseAviary.py:780-790
      o[0] = P.x; o[1] = P.y; o[2] = P.z;
      o[3] = Q.x; o[4] = Q.y; o[5] = Q.z; o[6] = Q.w;
      o[7] = roll; o[8] = pitch; o[9] = yaw;
      o[10] = V.x; o[11] = V.y; o[12] = V.z;
      o[13] = R.m00 * W.x + R.m01 * W.y + R.m02 * W.z;  // world angular velocity
      o[14] = R.m10 * W.x + R.m11 * W.y + R.m12 * W.z;
      o[15] = R.m20 * W.x + R.m21 * W.y + R.m22 * W.z;
      o[16] = C0.x; o[17] = C0.y; o[18] = C0.z; o[19] = C0.w; o[20] = C1.x; o[21] = C1.y;
    }
    if (a.neighbors) {  // BaseAviary._getAdjacencyMatrix (BaseAviary.py:901-921), strict <
      uint32_t bits = 1u << slot;
      for (int j = 0; j < a.D; ++j) {
        if (j == slot) continue;
        float4 O = a.s_pos[env0 + j];
        float dx = P.x - O.x, dy = P.y - O.y, dz = P.z - O.z;
        if (sqrtf(dx * dx + dy * dy + dz * dz) < a.radius) bits |= 1u << j;
      }
      a.neighbors[v] = bits;
    }
    if (slot == 0 && (a.done_env || a.reward_env)) {
      // env done: slot 0 reached the goal (the example tests drone "0"), any slot under the floor / out of time
      uint32_t any = 0;
      for (int j = 0; j < a.D; ++j) {
        uint32_t b = __float_as_uint(a.s_lv[env0 + j].w);
        any |= (j == 0) ? b : (b & 6u);
      }
      if (a.done_env) a.done_env[v / a.D] = any ? 1 : 0;
      if (a.reward_env) a.reward_env[v / a.D] = -1.0f;  // CtrlAviary.py:267-278
    }
  }
}

// ---------------------------------------------------------------------------------------------
// reset (BaseAviary._housekeeping BaseAviary.py:640-714, INDIControl.reset INDIControl.py:109-146)
// ---------------------------------------------------------------------------------------------
struct DsResetArgs {
  float4 *s_pos, *s_quat, *s_vel, *s_om, *s_lv, *s_lr, *s_c0, *s_a0;
  float2 *s_c1, *s_a1;
  const float *pos0, *rpy0, *vel0, *action0;
  const int32_t* wp0;
  const uint8_t* slot_type;
  const DsTypeDev* types;
  const float* init_cmd;     // [n_types]
  const float* init_thrust;  // [n_types]
  int n, n_pad, D;
};

__global__ void __launch_bounds__(256) ds_reset_kernel(const DsResetArgs a) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < a.n_pad; v += gridDim.x * blockDim.x) {
    const bool real = v < a.n;
    const int type_id = real ? a.slot_type[v % a.D] : 0;
    float px = 0.f, py = 0.f, pz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
    double r = 0.0, p = 0.0, y = 0.0;
    if (real) {
      px = a.pos0[3 * v]; py = a.pos0[3 * v + 1]; pz = a.pos0[3 * v + 2];
      if (a.rpy0) { r = a.rpy0[3 * v]; p = a.rpy0[3 * v + 1]; y = a.rpy0[3 * v + 2]; }
      if (a.vel0) { vx = a.vel0[3 * v]; vy = a.vel0[3 * v + 1]; vz = a.vel0[3 * v + 2]; }
    }
    // p.getQuaternionFromEuler(INIT_RPYS) (BaseAviary.py:688) in FP64, rounded once
    double sph = sin(0.5 * r), cph = cos(0.5 * r), sth = sin(0.5 * p), cth = cos(0.5 * p);
    double sps = sin(0.5 * y), cps = cos(0.5 * y);
    double qx = sph * cth * cps - cph * sth * sps, qy = cph * sth * cps + sph * cth * sps;
    double qz = cph * cth * sps - sph * sth * cps, qw = cph * cth * cps + sph * sth * sps;
    double n = 1.0 / sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
    const float ic = real ? a.init_cmd[type_id] : 0.f;
    const int nu = real ? a.types[type_id].n_u : 0;
    a.s_pos[v] = make_float4(px, py, pz, real ? a.init_thrust[type_id] : 0.f);
    a.s_quat[v] = make_float4((float)(qx * n), (float)(qy * n), (float)(qz * n), (float)(qw * n));
    a.s_vel[v] = make_float4(vx, vy, vz, real ? a.types[type_id].rpm0_sum : 0.f);
    a.s_om[v] = make_float4(0.f, 0.f, 0.f, __int_as_float((real && a.wp0) ? a.wp0[v] : 0));
    a.s_lv[v] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
    a.s_lr[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    a.s_c0[v] = make_float4(nu > 0 ? ic : 0.f, nu > 1 ? ic : 0.f, nu > 2 ? ic : 0.f, nu > 3 ? ic : 0.f);
    a.s_c1[v] = make_float2(nu > 4 ? ic : 0.f, nu > 5 ? ic : 0.f);
    float ac[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (real && a.action0)
      for (int i = 0; i < 6; ++i) ac[i] = (i < nu) ? a.action0[6 * v + i] : 0.f;
    a.s_a0[v] = make_float4(ac[0], ac[1], ac[2], ac[3]);
    a.s_a1[v] = make_float2(ac[4], ac[5]);
  }
}

// ---------------------------------------------------------------------------------------------
// diagnostic: the WLS allocator alone (fast path + FP64 active-set slow path), one problem per thread
// ---------------------------------------------------------------------------------------------
__global__ void ds_wls_kernel(const DsTypeDev* types, const DsWlsDev* wls, int type_id, const float* v, const float* cmd,
                              float* du_out, int* iter_out, int n, int force_slow) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const DsTypeDev& tp = types[type_id];
  const DsWlsDev* P = wls + type_id;
  float nu[6], du[6];
  bool feasible = true;
  for (int k = 0; k < 6; ++k) nu[k] = v[6 * i + k];
  for (int k = 0; k < 6; ++k) {
    const float* a = tp.alloc + k * 6;
    du[k] = a[0] * nu[0] + a[1] * nu[1] + a[2] * nu[2] + a[3] * nu[3] + a[4] * nu[4] + a[5] * nu[5];
    float umin = tp.rotor[k].pmin - cmd[6 * i + k], umax = tp.rotor[k].pmax - cmd[6 * i + k];
    feasible = feasible && !(du[k] >= umax + 1.0f || du[k] <= umin - 1.0f);
  }
  int it = 1;
  if (!feasible || force_slow) {
    double vv[6], umin[6], umax[6], u[6];
    for (int k = 0; k < 6; ++k) {
      vv[k] = (double)nu[k];
      umin[k] = P->pmin[k] - (double)cmd[6 * i + k];
      umax[k] = P->pmax[k] - (double)cmd[6 * i + k];
      u[k] = 0.0;
    }
    it = ds_wls_alloc(P, vv, umin, umax, u);
    for (int k = 0; k < 6; ++k) du[k] = (it > 0) ? (float)u[k] : 0.f;
  }
  for (int k = 0; k < 6; ++k) du_out[6 * i + k] = du[k];
  iter_out[i] = it;
}
