// Device-side data layout and math of the dronesim_b200 core (sm_100a).
//
// One thread advances one vehicle.  The resident state is a structure of float4 arrays so that a
// warp moves 512 contiguous bytes per load/store instruction (LDG.128 / STG.128).  Per-type
// airframe constants (rotor geometry, inertia, mixer / allocation matrix, gains) live in shared
// memory, indexed by the lane's type id.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ds_lanes.cuh"

// 128 vehicles per tile x 4 resident CTAs per SM (16 warps, 128 registers per thread): with the dynamic tile tickets the
// smaller tile halves the tail of a launch; measured 1-5 % faster than 256 x 2 on every workload (profiles/r01_notes.md)
#ifndef DS_TILE
#define DS_TILE 128          // threads per CTA = vehicles per tile (when drones_per_env divides it)
#endif
#ifndef DS_MIN_CTAS
#define DS_MIN_CTAS 4        // resident CTAs per SM the step kernel is compiled for (register budget)
#endif
#define DS_MAX_TYPES_DEV 8
#define DS_DW_ROWS (DS_TILE + DS_TILE / 2)       // float4 rows of one downwash position snapshot: (DS_TILE / D) envs x (D + 1) padded rows, D >= 2
#define DS_DW_BUF (2 * DS_TILE)                  // rows reserved per snapshot buffer: the symmetric D = 16 variant stores every row twice (DS_TILE / 16 envs x 32 rows)

#define DS_TYPE_PAD 24
struct __align__(16) DsRotorDev {
  float ax, ay, az, scale;   // thrust axis (body)            | PWM2RPM_SCALE
  float mx, my, mz, cnst;    // torque / unit thrust about CoM = (r - rc) x a + spin (km/kf) t | PWM2RPM_CONST
  float gx, gy, gz, pmin;    // (r - rc) x a  (ground-effect thrust has no reaction torque) | MIN_PWM
  float rx, ry, rz, pmax;    // rotor site relative to the centre of mass, r - rc (ground-effect heights) | MAX_PWM
};

struct __align__(16) DsTypeDev {
  DsRotorDev rotor[6];       // 96 floats: control-step constants (motor map, clip limits, hoisted rotor wrench)
  // ---- substep-loop constants, laid out as the register PAIRS the packed FP32 arithmetic (FFMA2) consumes, 16-byte
  // groups so that one LDS.128 delivers two pairs (ds_physics.cuh)
  float2 gh[3][3];           // rotor pair p = (2p, 2p+1): (rx, rx'), (ry, ry'), (rz, rz') - two rotor heights per FFMA2
  float2 gh_pad;
  float2 gw[6][3];           // rotor i: (ax, ay), (az, gx), (gy, gz) - wrench pairs (Fx,Fy) (Fz,tx) (ty,tz) += g_i * ...
  float2 Jc[3];              // inertia tensor, column pairs (J00,J10) (J01,J11) (J02,J12)
  float Jr[3];               // ... and its third row
  float dtm;                 // TIMESTEP / mass: the velocity update is u += dtm (R F) - TIMESTEP g
  float2 Jdc[3];             // J^-1 * TIMESTEP in the same layout: the rate update is w += Jd (tau - w x J w)
  float Jdr[3];
  float gnd_clip;            // GND_EFF_H_CLIP
  float2 nrc_xy;             // -rc (centre of mass in the base frame): p_base = c - R rc
  float nrc_z;
  float dw_k1n;              // -DW_COEFF_1 * (PROP_RADIUS/4)^2
  float2 ndk_xy;             // -DRAG_COEFF * 2 pi / 60
  float ndk_z;
  float kf;
  float dw_k2, dw_k3;        // DW_COEFF_2,3 divided by sqrt(0.5 log2 e): exp(-0.5 (d/beta)^2) = exp2(-(d/beta')^2)
  float gnd_k;               // GND_EFF_COEFF * (PROP_RADIUS/4)^2
  float pad0_;
  // ---- per control step / set-up
  float rc[3];
  float kp, kd;
  float att[3];
  float rate[3];
  float2 alloc2[3][6];       // allocation matrix, rows in pairs: alloc2[p][j] = (A[2p][j], A[2p+1][j]) - two commands per FFMA2
  float2 plo[3], phi[3];     // (MIN_PWM, MAX_PWM) of rotor pairs; rotors beyond n_u: 0 (their command stays 0)
  int n_u;
  int law;
  float rpm0_sum;            // sum_i PWM2RPM_CONST_i  (rpm of the all-zero action, BaseAviary.py:659-662)
  int has_rc;                // centre of mass vs base-frame origin (QUAT integrator): 0 same point, 1 general offset, 2 offset along body z only
  float speed_limit;         // MAX_SPEED_KMH * 1000 / 3600 (VelocityAviary.py:92-94)
  float lat[3];              // sum_i (r_i - rc): arm of the quad model's lateral noise force (BaseAviary.py:1528-1536)
  int rotor_model;           // 0 quad (_quad_copter_physics), 1 morphing hexa (_morphing_hexa_physics), 2 quad "advanced" (:1493-1512)
  float kf_over_km;          // turns the stored reaction-torque column (m - g) = spin km/kf t into spin t
  float adv[15];             // rotor_model 2: the 14 oblique-flow coefficients (Data_section5_ObliqueFlow) + propeller radius [m]
  float pad_[DS_TYPE_PAD];   // stride = 8 (mod 32) words: four types sit in disjoint shared-memory banks
};
static_assert(sizeof(DsTypeDev) % 16 == 0, "DsTypeDev must be float4-copyable");
static_assert((sizeof(DsTypeDev) / 4) % 32 == 8, "DsTypeDev bank stride");

// FP64 side table for the WLS active-set slow path (rarely touched, stays in global / L2)
struct DsWlsDev {
  double B[36];    // [n_v][n_u] = G1 / 0.05, row stride 6
  double Wv[6];
  double gamma;
  double pmin[6], pmax[6];
  int n_u, n_v;
};

struct DsArgs {
  // resident state (structure of float4 arrays)
  float4* s_pos;   // x y z | last_thrust
  float4* s_quat;  // x y z w
  float4* s_vel;   // x y z | rpm sum of last applied action
  float4* s_om;    // p q r | wp counter bits
  float4* s_lv;    // last_vel | done bits
  float4* s_lr;    // last_rates | |pos_e|
  float4* s_c0;    // controller cmd 0..3
  float2* s_c1;    // controller cmd 4..5
  float4* s_a0;    // last clipped action 0..3 (facade path / first step after reset)
  float2* s_a1;
  // extension state (north_star items beyond the reference; allocated only when enabled)
  float4* s_r0;    // actual rotor speed 0..3 (first-order motor model)
  float2* s_r1;    // actual rotor speed 4..5
  float4* s_af;    // filtered angular-acceleration estimate x y z | -
  // deferred WLS slow path: problems queued by the fused kernel for ds_wls_fixup_kernel (6-DOF types only)
  int* wls_count;  // entries queued by the step kernel; ds_wls_fixup_kernel re-arms it (and wls_count[1], its exit counter)
  int* wls_index;  // [n]
  float* wls_nu;   // [n][6]
  // dynamic tile scheduler: tiles beyond the first two of each CTA are handed out by an atomic counter; the holder of the
  // launch's last ticket re-arms it (no memset between launches, and a captured CUDA graph can be replayed as is)
  int* tile_counter;
  int* reserved_;  // was the exit counter of the end-of-kernel handshake.  Kept: without it every later member moves by 8 B in
                   // the constant bank and the 16-drone kernel comes out 1 % slower (ptxas pairs its parameter loads differently)
  const DsTypeDev* types;
  const DsWlsDev* wls;
  const uint8_t* slot_type;
  double* stats;
  int n;            // vehicles
  int D;            // drones per env
  int tile_v;       // vehicles per tile ( (DS_TILE / D) * D )
  int n_tiles;
  int K;
  int n_types;
  int homo_type;    // the one type id every slot flies (HOMO kernel variants), -1: mixed
  uint32_t flags;
  int order;
  int rc_kind;      // centre-of-mass offsets of the swarm's types: 0 none, 1 some general offset, 2 all along body z
  int use_act;      // physics reads the action from s_a0/s_a1 instead of the controller cmd
  int store_act;    // physics stores the clipped action to s_a0/s_a1
  float dt;         // TIMESTEP
  float gravity;
  float floor_z;    // DS_FLAG_GROUND_PLANE: hard floor for the centre of mass (run-time-flag kernel variants, a.flags bit 6)
  float dtg;        // TIMESTEP * gravity
  float qh;         // 0.25 TIMESTEP^2: (half angle)^2 of a substep = qh |w|^2
  float qk[4];      // TIMESTEP x Taylor coefficients of 0.5 sin(h)/h in h^2 (ds_quat_step)
  float ctrl_dt, inv_ctrl_dt;
  int ext;          // 1: the EXT kernel variant runs (motor model and / or angular-acceleration filter on)
  float motor_a;    // 1 - exp(-dt / tau_motor); >= 1: static map (BaseAviary.py:1487-1490)
  float acc_b;      // 1 - exp(-2 pi f_c ctrl_dt); >= 1: raw finite difference (INDIControl.py:432-439)
  // rotor noise (BaseAviary.py:1429-1432, 1518-1525), EXT variant only; sigma = 0: off
  float noise_f, noise_m;
  uint32_t seed_lo, seed_hi;
  uint32_t veh0;    // global id of this shard's vehicle 0 (env_offset * D): the noise stream does not depend on sharding
  uint32_t step0;   // step_counter at launch: substep index of k = 0
  // targets
  int tmode, num_wp, advance_wp;
  const float4* t_pos;
  const float4* t_vel;
  const float4* t_acc;
  const float4* t_table;
  const float4* t_off;
  const int32_t* t_wp;   // table mode: the caller's waypoint index per vehicle (nullptr: the resident counter)
  // done predicate
  int goal_en, floor_en, time_hit;
  const int32_t* env_t0; // per-env step counter at its last masked reset (nullptr: no masked reset yet, time_hit applies)
  int step_end, max_steps;  // step counter after this launch; per-env time limit: step_end - env_t0[env] >= max_steps
  float goal_x, goal_y, goal_z, goal_r2;   // goal_r2 = fl(r * r): the predicate compares squared distances
  float z_min;
  // per-env outputs of the fused step (optional, DEVICE): reduced with warp shuffles inside the step kernel when every env
  // sits inside one warp (D | 32); the host falls back to the observation kernel otherwise
  uint8_t* env_done;   // [n_envs]
  float* env_reward;   // [n_envs]
  int reward_mode;
  // external I/O of the non-fused entry points
  const float* ext_action;   // [n][6]
  const float* ext_state;    // [n][22]
  const float4* rate_thrust; // [n]
  float* cmd_out;            // [n][6]
  float* pos_e_out;          // [n][3]
  float* yaw_err_out;        // [n]
};

// ---------------------------------------------------------------------------------------------
// small math
// ---------------------------------------------------------------------------------------------
#define DS_PENDING_ACTION 0x80000000u  // done-bits word, bit 31: the action array holds this vehicle's first action (masked reset)
#define DS_PI_F 3.14159265358979323846f
#define DS_GIMBAL 0.99999f

// single-instruction MUFU forms ds_rcp / ds_ex2 / ds_rsqrt: ds_lanes.cuh

struct Mat3 { float m00, m01, m02, m10, m11, m12, m20, m21, m22; };

// btMatrix3x3::setRotation (oracle/pyb_math.py getMatrixFromQuaternion), s = 2/|q|^2
__device__ __forceinline__ Mat3 ds_rot(float x, float y, float z, float w, float s) {
  float xs = x * s, ys = y * s, zs = z * s;
  float wx = w * xs, wy = w * ys, wz = w * zs;
  float xx = x * xs, xy = x * ys, xz = x * zs;
  float yy = y * ys, yz = y * zs, zz = z * zs;
  Mat3 R;
  R.m00 = 1.0f - (yy + zz); R.m01 = xy - wz;          R.m02 = xz + wy;
  R.m10 = xy + wz;          R.m11 = 1.0f - (xx + zz); R.m12 = yz - wx;
  R.m20 = xz - wy;          R.m21 = yz + wx;          R.m22 = 1.0f - (xx + yy);
  return R;
}

// utils/math.py:75-80
__device__ __forceinline__ float ds_norm_ang(float x) {
  const float two_pi = 6.28318530717958647692f;
  if (x > DS_PI_F) x -= two_pi * ceilf((x - DS_PI_F) / two_pi);
  else if (x < -DS_PI_F) x += two_pi * ceilf((-DS_PI_F - x) / two_pi);
  return x;
}

// atan2 in ~20 instructions: min / max ratio through one MUFU reciprocal, degree-8 polynomial in t^2 on [0, 1]
// (least-squares Chebyshev fit of atan(t) / t; 1.1e-7 max absolute error evaluated in FP32), quadrant fix-ups by
// selects.  Total error <= 2e-7 rad, the same order as libm's atan2f (2 ulp); used by the controller only.
__device__ __forceinline__ float ds_atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = mn * ds_rcp(mx);
  const float s = t * t;
  float p = 0.0028340641874819994f;
  p = fmaf(p, s, -0.016005029901862144f); p = fmaf(p, s, 0.042587608098983765f); p = fmaf(p, s, -0.07495445758104324f);
  p = fmaf(p, s, 0.10636754333972931f);   p = fmaf(p, s, -0.14202570915222168f); p = fmaf(p, s, 0.19992484152317047f);
  p = fmaf(p, s, -0.3333306610584259f);   p = fmaf(p, s, 1.0f);
  p *= t;
  p = (ay > ax) ? (0.5f * DS_PI_F - p) : p;
  p = (x < 0.f) ? (DS_PI_F - p) : p;
  p = (mx == 0.f) ? 0.f : p;  // atan2(0, 0)
  return copysignf(p, y);
}

// pybullet getEulerFromQuaternion (oracle/pyb_math.py); returns gimbal flag.  FAST: the controller's variant
// (ds_atan2_fast; asin(s) = atan2(s, sqrt((1 - s)(1 + s))), |s| < 0.99999 here); the state vector / integrator use libm.
template <bool FAST = false>
__device__ __forceinline__ bool ds_euler(float x, float y, float z, float w, float& roll, float& pitch, float& yaw) {
  float sarg = -2.0f * (x * z - w * y);
  if (sarg <= -DS_GIMBAL) { roll = 0.f; pitch = -0.5f * DS_PI_F; yaw = 2.0f * (FAST ? ds_atan2_fast(x, -y) : atan2f(x, -y)); return true; }
  if (sarg >= DS_GIMBAL)  { roll = 0.f; pitch = 0.5f * DS_PI_F;  yaw = 2.0f * (FAST ? ds_atan2_fast(-x, y) : atan2f(-x, y)); return true; }
  float sqx = x * x, sqy = y * y, sqz = z * z, squ = w * w;
  if (FAST) {
    roll = ds_atan2_fast(2.0f * (y * z + w * x), squ - sqx - sqy + sqz);
    const float c2 = (1.0f - sarg) * (1.0f + sarg);
    pitch = ds_atan2_fast(sarg, c2 * ds_rsqrt(c2));
    yaw = ds_atan2_fast(2.0f * (x * y + w * z), squ + sqx - sqy - sqz);
  } else {
    roll = atan2f(2.0f * (y * z + w * x), squ - sqx - sqy + sqz);
    pitch = asinf(sarg);
    yaw = atan2f(2.0f * (x * y + w * z), squ + sqx - sqy - sqz);
  }
  return false;
}

// pybullet getQuaternionFromEuler.  FAST: the half-angle sines / cosines through MUFU.SIN / MUFU.COS (__sincosf, absolute
// error 2^-21.4 for |half angle| <= pi) - used for the controller's target quaternion only, where a 1e-6 quaternion error
// moves the command by < 4e-7 PWM (sensitivity = |pinv(G1/0.05)| rate_gain att_gain); the
// integrator path (DS_INTEG_RPY) keeps the libm version.
template <bool FAST = false>
__device__ __forceinline__ float4 ds_quat_from_euler(float r, float p, float y) {
  float sph, cph, sth, cth, sps, cps;
  if (FAST) {
    __sincosf(0.5f * r, &sph, &cph);
    __sincosf(0.5f * p, &sth, &cth);
    __sincosf(0.5f * y, &sps, &cps);
  } else {
    sincosf(0.5f * r, &sph, &cph);
    sincosf(0.5f * p, &sth, &cth);
    sincosf(0.5f * y, &sps, &cps);
  }
  float4 q;
  q.x = sph * cth * cps - cph * sth * sps;
  q.y = cph * sth * cps + sph * cth * sps;
  q.z = cph * cth * sps - sph * sth * cps;
  q.w = cph * cth * cps + sph * sth * sps;
  float n = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x *= n; q.y *= n; q.z *= n; q.w *= n;
  return q;
}

__device__ __forceinline__ float ds_clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// ---------------------------------------------------------------------------------------------
// counter-based noise: Philox-4x32-10 keyed by the seed, counter = (vehicle, substep, draw, 0); Box-Muller on 24-bit
// uniforms.  oracle/noise.py is the FP64 twin (same integers, same uniforms).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ds_philox4x32(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ void ds_box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincosf(6.28318530717958647692f * u2, &sn, &cs);
  n0 = r * cs; n1 = r * sn;
}
// 12 standard normals of (vehicle, substep): draws 0..2
__device__ __forceinline__ void ds_normals12(uint32_t veh, uint32_t substep, uint32_t k0, uint32_t k1, float n[12]) {
#pragma unroll
  for (uint32_t j = 0; j < 3; ++j) {
    const uint4 x = ds_philox4x32(make_uint4(veh, substep, j, 0u), k0, k1);
    ds_box_muller(x.x, x.y, n[4 * j], n[4 * j + 1]);
    ds_box_muller(x.z, x.w, n[4 * j + 2], n[4 * j + 3]);
  }
}
