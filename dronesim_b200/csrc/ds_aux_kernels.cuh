// The small kernels around the step kernel: control only, observation, reset, WLS diagnostic.
// Included by ds_api.cu alone (the step kernel's instantiations are separate translation units).
#pragma once
#include "ds_kernels.cuh"

// ---------------------------------------------------------------------------------------------
// control only (resident or external state; MODE 1 = rate/thrust entry of RPYTAviary)
// ---------------------------------------------------------------------------------------------
template <bool NU6, int MODE, bool EXT>
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
    if (EXT) { const float4 AF = a.s_af[v]; m.afx = AF.x; m.afy = AF.y; m.afz = AF.z; }
    if (MODE == 0) {
      const float4* tg = ds_staged_target(a);
      CtrlTarget t = ds_fetch_target(a, tp, cs, v, wp, tg ? tg + v : nullptr);
      ds_indi_control<NU6, EXT>(tp, a.wls, type_id, cs, t, a.inv_ctrl_dt, a.acc_b, m, o, true);
    } else {  // INDIControl._INDIRateControl (INDIControl.py:413-490)
      float4 rt = a.rate_thrust[v];
      float nu[4];
      ds_rate_loop<EXT>(tp, cs, a.inv_ctrl_dt, a.acc_b, rt.x, rt.y, rt.z, m, nu);
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
    if (EXT) a.s_af[v] = make_float4(m.afx, m.afy, m.afz, 0.f);
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
  const float4 *s_pos, *s_quat, *s_vel, *s_om, *s_lv, *s_lr, *s_c0;
  const float2* s_c1;
  const uint8_t* slot_type;
  const DsTypeDev* types;
  float* obs;
  uint32_t* neighbors;
  uint8_t* done_env;
  float* reward_env;
  int n, D, nu6;
  int reward_mode;  // 0: constant -1 (the reference); 1: -mean |pos_e| of the env
  float radius2;  // fl(NEIGHBOURHOOD_RADIUS^2)
};

// BaseAviary._getDroneStateVector (BaseAviary.py:780-790) of vehicle v, zero padded to 22 floats
__device__ __forceinline__ void ds_state_vector(const DsObsArgs& a, int v, float o[22]) {
  const float4 P = a.s_pos[v], Q = a.s_quat[v], V = a.s_vel[v], W = a.s_om[v], C0 = a.s_c0[v];
  const float2 C1 = a.nu6 ? a.s_c1[v] : make_float2(0.f, 0.f);
  float roll, pitch, yaw;
  ds_euler(Q.x, Q.y, Q.z, Q.w, roll, pitch, yaw);  // BaseAviary.py:729
  float d = Q.x * Q.x + Q.y * Q.y + Q.z * Q.z + Q.w * Q.w;
  Mat3 R = ds_rot(Q.x, Q.y, Q.z, Q.w, 2.0f / d);
  o[0] = P.x; o[1] = P.y; o[2] = P.z;
  o[3] = Q.x; o[4] = Q.y; o[5] = Q.z; o[6] = Q.w;
  o[7] = roll; o[8] = pitch; o[9] = yaw;
  o[10] = V.x; o[11] = V.y; o[12] = V.z;
  o[13] = R.m00 * W.x + R.m01 * W.y + R.m02 * W.z;  // world angular velocity
  o[14] = R.m10 * W.x + R.m11 * W.y + R.m12 * W.z;
  o[15] = R.m20 * W.x + R.m21 * W.y + R.m22 * W.z;
  o[16] = C0.x; o[17] = C0.y; o[18] = C0.z; o[19] = C0.w; o[20] = C1.x; o[21] = C1.y;
}

__global__ void __launch_bounds__(256) ds_obs_kernel(const DsObsArgs a) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < a.n; v += gridDim.x * blockDim.x) {
    const int slot = v % a.D, env0 = v - slot;
    float4 P = a.s_pos[v];
    if (a.obs) {
      float o[22];
      ds_state_vector(a, v, o);
      float* dst = a.obs + (size_t)v * 22;
#pragma unroll
      for (int i = 0; i < 22; ++i) dst[i] = o[i];
    }
    if (a.neighbors) {  // BaseAviary._getAdjacencyMatrix (BaseAviary.py:901-921), strict <
      uint32_t bits = 1u << slot;
      for (int j = 0; j < a.D; ++j) {
        if (j == slot) continue;
        float4 O = a.s_pos[env0 + j];
        // |p_i - p_j|^2 < r^2 with every operation correctly rounded in a fixed order (no FMA contraction, no square
        // root): the bit is a pure function of the FP32 positions, reproduced exactly by the same float32 expression
        const float dx = P.x - O.x, dy = P.y - O.y, dz = P.z - O.z;
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d2 < a.radius2) bits |= 1u << j;
      }
      a.neighbors[v] = bits;
    }
    if (slot == 0 && (a.done_env || a.reward_env)) {
      // env done: slot 0 reached the goal (the example tests drone "0"), any slot under the floor / out of time
      uint32_t any = 0;
      for (int j = 0; j < a.D; ++j) {
        uint32_t b = __float_as_uint(a.s_lv[env0 + j].w) & ~DS_PENDING_ACTION;
        any |= (j == 0) ? b : (b & 6u);
      }
      if (a.done_env) a.done_env[v / a.D] = any ? 1 : 0;
      if (a.reward_env) {
        float rw = -1.0f;  // CtrlAviary.py:267-278
        if (a.reward_mode == 1) {  // extension: minus the env's mean position error of the last control step
          float sum = 0.f;
          for (int j = 0; j < a.D; ++j) sum += a.s_lr[env0 + j].w;
          rw = -sum / (float)a.D;
        }
        a.reward_env[v / a.D] = rw;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// trajectory capture in the reference Logger's layout (dronesim/utils/Logger.py:51-53: states[drone][state][sample]):
// one thread per logged vehicle writes its state vector into column `col` of a [n_log][22][capacity] array
// ---------------------------------------------------------------------------------------------
__global__ void ds_log_kernel(const DsObsArgs a, const int32_t* __restrict__ ids, int n_log, float* __restrict__ states,
                              int capacity, int col) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_log) return;
  float o[22];
  ds_state_vector(a, ids[i], o);
  float* dst = states + (size_t)i * 22 * capacity + col;
#pragma unroll
  for (int k = 0; k < 22; ++k) dst[(size_t)k * capacity] = o[k];
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
  // masked reset (ds_reset_envs): only envs with mask[env] != 0 are touched; their time origin becomes step_now
  const uint8_t* mask;
  int32_t* env_t0;
  int step_now;
};

__global__ void __launch_bounds__(256) ds_reset_kernel(const DsResetArgs a) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < a.n_pad; v += gridDim.x * blockDim.x) {
    const bool real = v < a.n;
    if (a.mask) {
      if (!real || !a.mask[v / a.D]) continue;
      if (v % a.D == 0 && a.env_t0) a.env_t0[v / a.D] = a.step_now;
    }
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
    // a masked reset cannot use the launch-wide "first action pending" flag: the vehicle carries it in bit 31
    a.s_lv[v] = make_float4(0.f, 0.f, 0.f, __uint_as_float((a.mask && a.action0) ? DS_PENDING_ACTION : 0u));
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
// deferred WLS slow path: the problems the fused kernel queued (first iterate outside the +-1.0 slack, wls_alloc.py:264),
// one per thread: FP64 active-set solution, cmd = clip(cmd + du) (INDIControl_6DOF.py:630-631); the fused kernel held
// the command of these vehicles, and nothing reads it before the next step's physics.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ds_wls_fixup_kernel(const DsArgs a) {
  const int n = *a.wls_count;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const int v = a.wls_index[q];
    const int type_id = a.slot_type[v % a.D];
    const DsTypeDev& tp = a.types[type_id];
    const DsWlsDev* P = a.wls + type_id;
    const float4 C0 = a.s_c0[v];
    const float2 C1 = a.s_c1[v];
    const float cmd[6] = {C0.x, C0.y, C0.z, C0.w, C1.x, C1.y};
    double vv[6], umin[6], umax[6], u[6];
    for (int i = 0; i < 6; ++i) {
      vv[i] = (double)a.wls_nu[(size_t)v * 6 + i];
      umin[i] = P->pmin[i] - (double)cmd[i];
      umax[i] = P->pmax[i] - (double)cmd[i];
      u[i] = 0.0;
    }
    const int it = ds_wls_alloc(P, vv, umin, umax, u);
    float out[6];
    int sat = 0;
    for (int i = 0; i < 6; ++i) {
      const float c = cmd[i] + ((it > 0) ? (float)u[i] : 0.f);  // non-convergence: hold the command
      out[i] = ds_clampf(c, tp.rotor[i].pmin, tp.rotor[i].pmax);
      sat += (out[i] != c);
    }
    a.s_c0[v] = make_float4(out[0], out[1], out[2], out[3]);
    a.s_c1[v] = make_float2(out[4], out[5]);
    if (a.flags & 8u) {
      if (sat) atomicAdd(a.stats + ST_SAT, (double)sat);
      if (it < 0) atomicAdd(a.stats + ST_WLS_FAIL, 1.0);
    }
  }
  // the last block out empties the queue for the next step kernel (every block has read the count by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(a.wls_count + 1, 1) == (int)gridDim.x - 1) {
      a.wls_count[0] = 0;
      a.wls_count[1] = 0;
    }
  }
}

// extension state after reset: rotor speeds of the all-zero action, filter state zero
__global__ void __launch_bounds__(256) ds_reset_ext_kernel(const DsResetArgs a, float4* s_r0, float2* s_r1, float4* s_af) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < a.n_pad; v += gridDim.x * blockDim.x) {
    float r[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (a.mask && (v >= a.n || !a.mask[v / a.D])) continue;
    if (v < a.n) {
      const DsTypeDev& tp = a.types[a.slot_type[v % a.D]];
      for (int i = 0; i < tp.n_u; ++i) r[i] = tp.rotor[i].cnst;  // last_clipped_action = 0 after reset (BaseAviary.py:659-662)
    }
    s_r0[v] = make_float4(r[0], r[1], r[2], r[3]);
    s_r1[v] = make_float2(r[4], r[5]);
    s_af[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// diagnostic: the WLS allocator alone (fast path + FP64 active-set slow path), one problem per thread
// ---------------------------------------------------------------------------------------------
__global__ void ds_wls_kernel(const DsTypeDev* types, const DsWlsDev* wls, int type_id, const float* v, const float* cmd,
                              float* du_out, int* iter_out, int* w_out, int n, int force_slow) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const DsTypeDev& tp = types[type_id];
  const DsWlsDev* P = wls + type_id;
  float nu[6], du[6];
  bool feasible = true;
  for (int k = 0; k < 6; ++k) nu[k] = v[6 * i + k];
  for (int k = 0; k < 6; ++k) {
    const float2* a2 = tp.alloc2[k / 2];  // rows in pairs: row k is the .x (even k) / .y (odd k) half
    float a[6];
    for (int j = 0; j < 6; ++j) a[j] = (k & 1) ? a2[j].y : a2[j].x;
    du[k] = a[0] * nu[0] + a[1] * nu[1] + a[2] * nu[2] + a[3] * nu[3] + a[4] * nu[4] + a[5] * nu[5];
    float umin = tp.rotor[k].pmin - cmd[6 * i + k], umax = tp.rotor[k].pmax - cmd[6 * i + k];
    feasible = feasible && (du[k] < umax + (1.0f - DS_WLS_MARGIN)) && (du[k] > umin - (1.0f - DS_WLS_MARGIN));
  }
  int it = 1;
  int W[6] = {0, 0, 0, 0, 0, 0};  // a feasible first iterate leaves the working set empty (wls_alloc.py:171)
  if (!feasible || force_slow) {
    double vv[6], umin[6], umax[6], u[6];
    for (int k = 0; k < 6; ++k) {
      vv[k] = (double)nu[k];
      umin[k] = P->pmin[k] - (double)cmd[6 * i + k];
      umax[k] = P->pmax[k] - (double)cmd[6 * i + k];
      u[k] = 0.0;
    }
    it = ds_wls_alloc(P, vv, umin, umax, u, W);
    for (int k = 0; k < 6; ++k) du[k] = (it > 0) ? (float)u[k] : 0.f;
  }
  for (int k = 0; k < 6; ++k) du_out[6 * i + k] = du[k];
  iter_out[i] = it;
  if (w_out) for (int k = 0; k < 6; ++k) w_out[6 * i + k] = W[k];
}
