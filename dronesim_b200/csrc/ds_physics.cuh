// Rigid-body dynamics + aerodynamic add-ons, K substeps in registers, one vehicle per thread.
//
// Restates (see oracle/dynamics.py for the FP64 twin and the list of repairs R1-R8):
//   motor map + rotor thrust/torque  dronesim/envs/BaseAviary.py:1487-1543 (quad), :1398-1457 (hexa)
//   _dynamics                        BaseAviary.py:1767-1828   (DS_INTEG_RPY)
//   _groundEffect                    BaseAviary.py:1648-1699
//   _drag                            BaseAviary.py:1705-1732
//   _downwash                        BaseAviary.py:1736-1763
// DS_INTEG_QUAT is Newton-Euler about the composite centre of mass with an exponential-map
// quaternion update (asked by north_star; beyond the reference).
//
// The kernel is bound by instruction issue (ncu: issue slots > 80 % busy, DRAM ~ 11 %), so the
// code below is written to minimise issued instructions per substep:
//  * the command is constant across the K substeps of a control step (BaseAviary.py:507-545), so
//    the rotor wrench sum_i T_i a_i, sum_i T_i m_i is hoisted out of the substep loop; only the terms
//    that depend on the moving state (ground effect, drag, downwash, gyroscopic torque) are per substep;
//  * reciprocal / exp2 / rsqrt are the single-instruction MUFU forms (ds_rcp, ds_ex2, ds_rsqrt);
//  * the downwash pair term is 19 instructions (see ds_downwash_pair): the Gaussian's 0.5 log2(e) is
//    folded into the per-type beta coefficients, DW_COEFF_1 is applied once after the neighbour loop;
//  * state is integrated in centre-of-mass coordinates; rotor sites are stored relative to the centre
//    of mass so the ground-effect heights need no base-frame conversion.
#pragma once
#include "ds_device.cuh"

struct PhysState {
  float px, py, pz;      // base-frame origin, world
  float qx, qy, qz, qw;
  float vx, vy, vz;      // velocity of the base-frame origin, world
  float wx, wy, wz;      // body rates
};

#define DS_DW_PAD 1       // shared-memory position rows are padded to D + 1 float4: envs of a warp hit disjoint banks

__device__ __forceinline__ void ds_quat_step(float& qx, float& qy, float& qz, float& qw, float wx, float wy, float wz,
                                             float dt) {
  // q <- normalize(q (x) exp(w dt)); polynomial sin/cos of the half angle (|half| < 0.5), libm beyond
  float tx = wx * dt, ty = wy * dt, tz = wz * dt;
  float h2 = 0.25f * (tx * tx + ty * ty + tz * tz);  // (angle/2)^2
  float k, c;
  if (h2 < 0.25f) {
    // Taylor in h2 = (angle/2)^2 < 0.25: the first dropped terms are 5e-9 (sine) and 3e-10 (cosine) relative at h2 = 0.25
    k = 0.5f + h2 * (-0.5f / 6.0f + h2 * (0.5f / 120.0f + h2 * (-0.5f / 5040.0f)));
    c = 1.0f + h2 * (-0.5f + h2 * (1.0f / 24.0f + h2 * (-1.0f / 720.0f + h2 * (1.0f / 40320.0f))));
  } else {
    // |w| dt >= 1 rad per substep (>= 240 rad/s): a blown-up state; MUFU sin / cos keep it finite and this branch free
    // of libm's argument-reduction slow path (two CALLs inside the substep loop otherwise)
    float half = sqrtf(h2), s;
    __sincosf(half, &s, &c);
    k = 0.5f * s * ds_rcp(half);
  }
  float dx = tx * k, dy = ty * k, dz = tz * k, dw = c;
  float nx = qw * dx + qx * dw + qy * dz - qz * dy;
  float ny = qw * dy - qx * dz + qy * dw + qz * dx;
  float nz = qw * dz + qx * dy - qy * dx + qz * dw;
  float nw = qw * dw - qx * dx - qy * dy - qz * dz;
  float n = ds_rsqrt(nx * nx + ny * ny + nz * nz + nw * nw);
  qx = nx * n; qy = ny * n; qz = nz * n; qw = nw * n;
}

// Downwash of every drone of the env on this one (BaseAviary.py:1747-1763), in units of DW_COEFF_1
// (PROP_RADIUS / 4)^2: sum_j [dz > 0, dxy < 10] exp(-0.5 (dxy / beta)^2) / dz^2, beta = DW2 dz + DW3.
// row: the env's position snapshot in shared memory; k2, k3 are pre-divided by sqrt(0.5 log2 e).
// 19 issued instructions per pair: LDS.128, 3 FADD, FMUL+FFMA (dxy^2), FFMA (beta), FMUL+MUFU.RCP (1/dz^2),
// MUFU.RCP (1/beta), 2 FMUL + MUFU.EX2 (Gaussian), FMUL, 2 FSETP (one predicate), predicated FADD.
// dz > 0 also drops the drone itself.  beta == 0 (dz = -DW3 / DW2 exactly) gives 1/beta = inf -> exp2(-inf) = 0,
// the reference's exp(-inf) (BaseAviary.py:1755), with no extra test.
__device__ __forceinline__ void ds_downwash_pair(float& acc, const float4 o, float px, float py, float pz, float k2,
                                                 float k3) {
  const float dz = o.z - pz, dx = o.x - px, dy = o.y - py;
  const float d2 = fmaf(dy, dy, dx * dx);
  const float ib = ds_rcp(fmaf(k2, dz, k3));   // 1 / beta'
  const float idz2 = ds_rcp(dz * dz);          // 1 / dz^2
  const float w = idz2 * ds_ex2((ib * ib) * -d2);
  asm("{\n\t.reg .pred p;\n\t"
      "setp.gt.f32 p, %1, 0f00000000;\n\t"
      "setp.lt.and.f32 p, %2, 0f42C80000, p;\n\t"  // dxy^2 < 100
      "@p add.f32 %0, %0, %3;\n\t}"
      : "+f"(acc)
      : "f"(dz), "f"(d2), "f"(w));
}

__device__ __forceinline__ float ds_downwash_sum(const float4* __restrict__ row, int D, float px, float py, float pz,
                                                 float k2, float k3) {
  float acc = 0.f;
#ifndef DS_PAIR_UNROLL
#define DS_PAIR_UNROLL 16
#endif
  constexpr int kUnroll = DS_PAIR_UNROLL;
  if (D == 16) {  // BASELINE configs[3]/[4]
#pragma unroll kUnroll
    for (int j = 0; j < 16; ++j) ds_downwash_pair(acc, row[j], px, py, pz, k2, k3);
  } else {
#pragma unroll 4
    for (int j = 0; j < D; ++j) ds_downwash_pair(acc, row[j], px, py, pz, k2, k3);
  }
  return acc;
}

// ---- symmetric variant (D == 16, every type shares DW_COEFF_2 / DW_COEFF_3 - true of all shipped URDFs) ----------
// Only the LOWER vehicle of a pair feels the pair's downwash (dz > 0 gate), and the pair term depends on the two
// positions and (k2, k3) alone, so each of the 120 unordered pairs of an env is evaluated ONCE: in round r = 1..7
// slot s evaluates the pair (s, s + r mod 16), keeps the term if it is the lower one, otherwise hands it to the
// partner by a 16-lane-wide SHFL.IDX; round 8 pairs (s, s + 8) are evaluated from both sides and kept locally.
// 7 x 22 + 17 issued instructions and 24 MUFU per substep instead of 16 x 19 and 48.  The env's snapshot rows are
// stored twice (rows s and s + 16 of a 32-row block) so the partner row s + r needs no wrap-around arithmetic.
#define DS_DW_SYM_ROWS 32
template <bool SEND>
__device__ __forceinline__ float ds_downwash_pair_sym(float& acc, const float4 o, float px, float py, float pz, float k2,
                                                      float k3) {
  const float dz = o.z - pz, dx = o.x - px, dy = o.y - py;
  const float d2 = fmaf(dy, dy, dx * dx);
  const float ib = ds_rcp(fmaf(k2, fabsf(dz), k3));  // 1 / beta' of the lower vehicle (its dz is |dz|)
  const float idz2 = ds_rcp(dz * dz);
  const float w = idz2 * ds_ex2((ib * ib) * -d2);
  float theirs = 0.f;
  if (SEND) {
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.lt.f32 q, %3, 0f42C80000;\n\t"          // dxy^2 < 100
        "setp.gt.and.f32 p, %2, 0f00000000, q;\n\t"   // partner above me: the term is mine
        "@p add.f32 %0, %0, %4;\n\t"
        "setp.lt.and.f32 p, %2, 0f00000000, q;\n\t"   // partner below me: the term is the partner's
        "selp.f32 %1, %4, 0f00000000, p;\n\t}"
        : "+f"(acc), "=f"(theirs)
        : "f"(dz), "f"(d2), "f"(w));
  } else {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.gt.f32 p, %1, 0f00000000;\n\t"
        "setp.lt.and.f32 p, %2, 0f42C80000, p;\n\t"
        "@p add.f32 %0, %0, %3;\n\t}"
        : "+f"(acc)
        : "f"(dz), "f"(d2), "f"(w));
  }
  return theirs;
}

// row: this thread's own row (slot) in the env's 32-row block; slot16 = slot + 16
__device__ __forceinline__ float ds_downwash_sum_sym16(const float4* __restrict__ row, int slot16, float px, float py,
                                                       float pz, float k2, float k3) {
  float acc = 0.f, recv = 0.f;
#pragma unroll
  for (int r = 1; r <= 7; ++r) {
    const float theirs = ds_downwash_pair_sym<true>(acc, row[r], px, py, pz, k2, k3);
    recv += __shfl_sync(0xffffffffu, theirs, slot16 - r, 16);  // from slot - r (mod 16) of my env
  }
  ds_downwash_pair_sym<false>(acc, row[8], px, py, pz, k2, k3);
  return acc + recv;
}

// act[] must already be clipped.  prev_rpm_sum: in = sum of rpm of the previously applied action
// (BaseAviary.py:532), out = sum of rpm of this action.
// FX >= 0: ground effect (bit 0) / drag (bit 1) resolved at compile time; FX < 0: run-time a.flags.
// Centre-of-mass offset rc (QUAT integrator only; DsTypeDev::has_rc: 0 none, 1 general, 2 along body z only): the
// state is integrated at the centre of mass, the add-ons that the reference applies at the base-frame
// origin (drag, downwash: R7 of oracle/dynamics.py) see p_base = c - R rc, v_base = u - R (w x rc) and add the
// torque (-rc) x f.
// EXT: first-order motor model (north_star; R9 of oracle/dynamics.py): rpm[] holds the actual rotor speeds (in / out),
// every substep moves them towards the commanded speed and rebuilds the rotor wrench, so nothing is hoisted.
template <int INTEG, int DW, bool NU6, bool WARPSYNC, int FX, bool EXT>
__device__ __forceinline__ void ds_physics(const DsArgs& a, const DsTypeDev& tp, int env_row0, int my_row, float4* sh_pos,
                                           const float* act, PhysState& s, float& prev_rpm_sum, float* rpm_state,
                                           uint32_t veh_id) {
  constexpr int NU = NU6 ? 6 : 4;
  const float dt = a.dt;
  const bool gnd = (FX >= 0) ? ((FX & 1) != 0) : ((a.flags & 1u) != 0);
  const bool drag = (FX >= 0) ? ((FX & 2) != 0) : ((a.flags & 2u) != 0);
  // a.rc_kind is the same for every lane (1 if some type of the swarm has a general offset, 2 if all offsets are along
  // z, 0 if none): types without an offset run the same arithmetic with rc = 0, which is exact (x + 0 = x), so a warp
  // that mixes airframes does not diverge here
  const int rc_kind = (INTEG == 0) ? a.rc_kind : 0;
  const bool has_rc = rc_kind != 0;
  const bool rc_gen = rc_kind == 1;

  const float rcx = tp.rc[0], rcy = tp.rc[1], rcz = tp.rc[2];
  const float rc2x = 2.0f * rcx, rc2y = 2.0f * rcy, rc2z = 2.0f * rcz;

  float roll = 0.f, pitch = 0.f, yaw = 0.f;
  if (INTEG == 1) ds_euler(s.qx, s.qy, s.qz, s.qw, roll, pitch, yaw);  // state cache rpy (BaseAviary.py:729)

  // R rc and R (w x rc) for the current attitude / rates
  auto rot_rc = [&](const Mat3& R, float& ox, float& oy, float& oz) {
    ox = R.m02 * rcz; oy = R.m12 * rcz; oz = R.m22 * rcz;
    if (rc_gen) {
      ox += R.m00 * rcx + R.m01 * rcy; oy += R.m10 * rcx + R.m11 * rcy; oz += R.m20 * rcx + R.m21 * rcy;
    }
  };
  auto rot_wxrc = [&](const Mat3& R, float& ox, float& oy, float& oz) {
    float kx = s.wy * rcz, ky = -s.wx * rcz;  // w x rc, rc along z
    if (rc_gen) {
      kx -= s.wz * rcy; ky += s.wz * rcx;
      const float kz = s.wx * rcy - s.wy * rcx;
      ox = R.m00 * kx + R.m01 * ky + R.m02 * kz; oy = R.m10 * kx + R.m11 * ky + R.m12 * kz;
      oz = R.m20 * kx + R.m21 * ky + R.m22 * kz;
    } else {
      ox = R.m00 * kx + R.m01 * ky; oy = R.m10 * kx + R.m11 * ky; oz = R.m20 * kx + R.m21 * ky;
    }
  };

  // QUAT: integrate centre-of-mass position / velocity
  float cx = s.px, cy = s.py, cz = s.pz, ux = s.vx, uy = s.vy, uz = s.vz;
  if (has_rc) {
    const Mat3 R = ds_rot(s.qx, s.qy, s.qz, s.qw, 2.0f);
    float ox, oy, oz;
    rot_rc(R, ox, oy, oz);
    cx += ox; cy += oy; cz += oz;
    rot_wxrc(R, ox, oy, oz);
    ux += ox; uy += oy; uz += oz;
  }
  // ---- rotor thrusts and the rotor part of the body wrench: once per control step (the command is constant across
  // the substeps), or once per substep when the motor model moves the rotor speeds
  float Tg[NU];  // T_i * GND_EFF_COEFF (PROP_RADIUS/4)^2: the only per-rotor value the substeps need
  float rpm_sum = 0.f, F0x = 0.f, F0y = 0.f, F0z = 0.f, t0x = 0.f, t0y = 0.f, t0z = 0.f;
  auto rotor_wrench = [&](int k) {
    rpm_sum = 0.f; F0x = 0.f; F0y = 0.f; F0z = 0.f; t0x = 0.f; t0y = 0.f; t0z = 0.f;
    // rotor noise (EXT; BaseAviary.py:1429-1432, 1518-1525): per substep N(0, sigma_f) on every thrust, N(0, sigma_m) on
    // every reaction torque; the quad model also puts (f_noise[0], f_noise[1]) on every rotor link laterally and
    // (m_noise[0], m_noise[1]) on the base (:1528-1543).  Ground effect keeps the noise-free thrust (:1680-1685).
    const bool noisy = EXT && (a.noise_f > 0.f || a.noise_m > 0.f) && tp.rotor_model != 2;  // no noise on the advanced branch
    float nz[12];
    if (noisy) {
      ds_normals12(veh_id, a.step0 + (uint32_t)k, a.seed_lo, a.seed_hi, nz);
      if (tp.rotor_model == 0) {
        const float l0 = a.noise_f * nz[0], l1 = a.noise_f * nz[1], nn = (float)tp.n_u;
        F0x = nn * l0; F0y = nn * l1;
        t0x = -tp.lat[2] * l1 + a.noise_m * nz[6]; t0y = tp.lat[2] * l0 + a.noise_m * nz[7];
        t0z = tp.lat[0] * l1 - tp.lat[1] * l0;
      }
    }
    // "advanced" quad types (EXT; BaseAviary.py:1493-1512 -> _get_prop_FMs :1570-1644 -> utils.py:149-202, 343-416, method 2):
    // oblique-flow propeller fit instead of KF rpm^2 / KM rpm^2; flow angles from the base velocity rotated by R (sic)
    const bool adv = EXT && tp.rotor_model == 2;
    float adv_vs = 0.f, adv_vc = 0.f, adv_cp = 1.f, adv_sp = 0.f;  // V sin(beta), V cos(beta), cos(psi), sin(psi)
    if (adv) {
      const Mat3 Ra = ds_rot(s.qx, s.qy, s.qz, s.qw, 2.0f);
      float vx = ux, vy = uy, vz = uz;
      if (has_rc) { float ox, oy, oz; rot_wxrc(Ra, ox, oy, oz); vx -= ox; vy -= oy; vz -= oz; }
      const float V = sqrtf(vx * vx + vy * vy + vz * vz);
      if (!(V > 0.1f)) { vx = 0.1f; vy = 0.f; vz = 0.f; }  // :1585-1589
      const float bx = Ra.m00 * vx + Ra.m01 * vy + Ra.m02 * vz, by = Ra.m10 * vx + Ra.m11 * vy + Ra.m12 * vz;
      const float bz = Ra.m20 * vx + Ra.m21 * vy + Ra.m22 * vz;
      const float cb = ds_clampf(bz * rsqrtf(bx * bx + by * by + bz * bz), -1.f, 1.f);  // cos(beta), beta = arccos (:1600)
      adv_vc = V * cb; adv_vs = V * sqrtf(fmaxf(1.f - cb * cb, 0.f));
      if (bx > 0.1f) {  // psi = arctan(by / bx) (:1603-1605): cos > 0
        const float t = by / bx, ic = rsqrtf(1.f + t * t);
        adv_cp = ic; adv_sp = t * ic;
      }
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) {  // rotors beyond n_u have scale = const = 0 -> T = 0
      const DsRotorDev& r = tp.rotor[i];
      float rpm = fmaf(r.scale, act[i], r.cnst);  // BaseAviary.py:1487-1490
      if (EXT) {
        rpm = (a.motor_a >= 1.f) ? rpm : fmaf(a.motor_a, rpm - rpm_state[i], rpm_state[i]);
        rpm_state[i] = rpm;
      }
      rpm_sum += rpm;
      float T = tp.kf * rpm * rpm;                // :1515
      Tg[i] = T * tp.gnd_k;
      if (adv) {
        if (i < tp.n_u) {
          const float* c = tp.adv;  // CstaticFT k1 k2 k3 k4 k5 CstaticMQ k6 k7 k8 k9 k10 k11 k12 | radius
          const float Rp = c[14];
          const float om = fmaxf(rpm * (6.28318530717958647692f / 60.0f), 10.0f);  // utils.py:176
          const float inv = 1.0f / (om * Rp);
          const float mu = adv_vs * inv, lc = adv_vc * inv;                          // utils.py:383-384
          const float cft = c[0] + c[1] * lc + c[2] * mu * mu + c[3] * lc * lc;      // eq. 95
          const float cfh = c[4] * mu + c[5] * lc * mu;                              // eq. 99
          const float cmr = c[10] * mu + c[11] * lc * mu;                            // eq. 101
          const float avg = 0.5f * 1.225f * (om * Rp) * (om * Rp) * (3.14159265358979323846f * Rp * Rp);
          const float fh = cfh * avg, ft = cft * avg, mz = cmr * avg * Rp * ((i & 1) ? 1.f : -1.f);  // direction (:1496)
          const float fx = adv_cp * fh, fy = adv_sp * fh;                            // R_z(psi) (:1633-1641)
          F0x += fx; F0y += fy; F0z += ft;
          t0x += r.ry * ft - r.rz * fy; t0y += r.rz * fx - r.rx * ft; t0z += r.rx * fy - r.ry * fx + mz;
        }
        continue;
      }
      F0x = fmaf(T, r.ax, F0x); F0y = fmaf(T, r.ay, F0y); F0z = fmaf(T, r.az, F0z);
      t0x = fmaf(T, r.mx, t0x); t0y = fmaf(T, r.my, t0y); t0z = fmaf(T, r.mz, t0z);
      if (noisy && i < tp.n_u) {
        const float nf = a.noise_f * nz[i], nm = a.noise_m * nz[6 + i] * tp.kf_over_km;
        F0x = fmaf(nf, r.ax, F0x); F0y = fmaf(nf, r.ay, F0y); F0z = fmaf(nf, r.az, F0z);
        t0x += nf * r.gx + nm * (r.mx - r.gx); t0y += nf * r.gy + nm * (r.my - r.gy); t0z += nf * r.gz + nm * (r.mz - r.gz);
      }
    }
  };
  if (!EXT) rotor_wrench(0);
  // drag coefficient x rotor speed sum: the first substep still sees the previously applied action (:532,:545)
  const float dk0 = -tp.drag_k[0], dk1 = -tp.drag_k[1], dk2 = -tp.drag_k[2];

  for (int k = 0; k < a.K; ++k) {
    if (EXT) {  // drag sees the rotor speeds before this substep's motor update
      if (k > 0) prev_rpm_sum = rpm_sum;
      rotor_wrench(k);
    }
    // ---- downwash first (BaseAviary.py:1747-1763): it needs the base-frame origin only, so the rotation matrix
    // does not have to stay live (or be rematerialised) across the unrolled pair loop
    float dw_fz = 0.f;
    if (DW) {  // every drone of the env reads the same position snapshot
      float px = cx, py = cy, pz = cz;  // base-frame origin
      if (has_rc) {  // R rc by the quaternion sandwich rc + w t + qv x t, t = 2 qv x rc: 15 operations, no matrix
        const float t0 = s.qy * rc2z - s.qz * rc2y, t1 = s.qz * rc2x - s.qx * rc2z, t2 = s.qx * rc2y - s.qy * rc2x;
        px -= fmaf(s.qw, t0, rcx) + (s.qy * t2 - s.qz * t1);
        py -= fmaf(s.qw, t1, rcy) + (s.qz * t0 - s.qx * t2);
        pz -= fmaf(s.qw, t2, rcz) + (s.qx * t1 - s.qy * t0);
      }
      float dsum;
      float4* buf = sh_pos + (k & 1) * DS_DW_BUF;
      if (DW == 2) {  // symmetric pairs: env_row0 / my_row index 32-row blocks, slot = my_row - env_row0
        const float4 me = make_float4(px, py, pz, 0.f);
        buf[my_row] = me;
        if (my_row - env_row0 < 8) buf[my_row + 16] = me;
        __syncwarp();
        dsum = ds_downwash_sum_sym16(buf + my_row, my_row - env_row0 + 16, px, py, pz, tp.dw_k2, tp.dw_k3);
      } else {
        buf[my_row] = make_float4(px, py, pz, 0.f);
        if (WARPSYNC) __syncwarp(); else __syncthreads();
        dsum = ds_downwash_sum(buf + env_row0, a.D, px, py, pz, tp.dw_k2, tp.dw_k3);
      }
      dw_fz = -tp.dw_k1 * dsum;
    }

    const Mat3 R = ds_rot(s.qx, s.qy, s.qz, s.qw, 2.0f);
    float Fx = F0x, Fy = F0y, Fz = F0z + dw_fz, tx = t0x, ty = t0y, tz = t0z;
    if (DW && rc_gen) { tx += -rcy * dw_fz; ty += rcx * dw_fz; }  // (-rc) x (0, 0, f)

    if (gnd) {  // BaseAviary.py:1672-1699; rotor sites are stored relative to the centre of mass
      bool gate;
      if (INTEG == 1) gate = (fabsf(roll) < 0.5f * DS_PI_F) && (fabsf(pitch) < 0.5f * DS_PI_F);
      else gate = (R.m22 > 0.f) && (fabsf(R.m20) < DS_GIMBAL);  // |roll| < pi/2 <=> cos(roll)cos(pitch) > 0
      const float gsel = gate ? 1.f : 0.f;  // branch-free: the gate is false only for an inverted vehicle
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        const DsRotorDev& r = tp.rotor[i];
        float h = fmaf(R.m20, r.rx, fmaf(R.m21, r.ry, fmaf(R.m22, r.rz, cz)));
        float ih = ds_rcp(fmaxf(h, tp.gnd_clip));
        float g = (Tg[i] * ih) * (ih * gsel);
        Fx = fmaf(g, r.ax, Fx); Fy = fmaf(g, r.ay, Fy); Fz = fmaf(g, r.az, Fz);
        tx = fmaf(g, r.gx, tx); ty = fmaf(g, r.gy, ty); tz = fmaf(g, r.gz, tz);
      }
    }
    if (drag) {  // BaseAviary.py:1719-1732
      float vx = ux, vy = uy, vz = uz;
      if (has_rc) {  // velocity of the base origin
        float ox, oy, oz;
        rot_wxrc(R, ox, oy, oz);
        vx -= ox; vy -= oy; vz -= oz;
      }
      const float sum = (EXT || k == 0) ? prev_rpm_sum : rpm_sum;
      float d0 = (dk0 * sum) * vx, d1 = (dk1 * sum) * vy, d2 = (dk2 * sum) * vz;
      float fx = R.m00 * d0 + R.m01 * d1 + R.m02 * d2;
      float fy = R.m10 * d0 + R.m11 * d1 + R.m12 * d2;
      float fz = R.m20 * d0 + R.m21 * d1 + R.m22 * d2;
      Fx += fx; Fy += fy; Fz += fz;
      if (has_rc) {  // (-rc) x f
        tx = fmaf(rcz, fy, tx); ty = fmaf(-rcz, fx, ty);
        if (rc_gen) { tx += -rcy * fz; ty += rcx * fz; tz += -rcx * fy + rcy * fx; }
      }
    }

    // ---- Newton-Euler (BaseAviary.py:1790-1807)
    float awx = (R.m00 * Fx + R.m01 * Fy + R.m02 * Fz) * tp.inv_mass;
    float awy = (R.m10 * Fx + R.m11 * Fy + R.m12 * Fz) * tp.inv_mass;
    float awz = (R.m20 * Fx + R.m21 * Fy + R.m22 * Fz) * tp.inv_mass - a.gravity;
    const float* J = tp.J;
    const float* Ji = tp.Jinv;
    float jx = J[0] * s.wx + J[1] * s.wy + J[2] * s.wz;
    float jy = J[3] * s.wx + J[4] * s.wy + J[5] * s.wz;
    float jz = J[6] * s.wx + J[7] * s.wy + J[8] * s.wz;
    float gx = tx - (s.wy * jz - s.wz * jy);
    float gy = ty - (s.wz * jx - s.wx * jz);
    float gz = tz - (s.wx * jy - s.wy * jx);
    float wdx = Ji[0] * gx + Ji[1] * gy + Ji[2] * gz;
    float wdy = Ji[3] * gx + Ji[4] * gy + Ji[5] * gz;
    float wdz = Ji[6] * gx + Ji[7] * gy + Ji[8] * gz;
    // ---- semi-implicit Euler (:1809-1812)
    ux = fmaf(dt, awx, ux); uy = fmaf(dt, awy, uy); uz = fmaf(dt, awz, uz);
    s.wx = fmaf(dt, wdx, s.wx); s.wy = fmaf(dt, wdy, s.wy); s.wz = fmaf(dt, wdz, s.wz);
    cx = fmaf(dt, ux, cx); cy = fmaf(dt, uy, cy); cz = fmaf(dt, uz, cz);
    if (INTEG == 1) {
      roll = fmaf(dt, s.wx, roll); pitch = fmaf(dt, s.wy, pitch); yaw = fmaf(dt, s.wz, yaw);
      float4 q = ds_quat_from_euler(roll, pitch, yaw);  // :1817
      s.qx = q.x; s.qy = q.y; s.qz = q.z; s.qw = q.w;
      ds_euler(s.qx, s.qy, s.qz, s.qw, roll, pitch, yaw);  // state refresh (:729)
    } else {
      ds_quat_step(s.qx, s.qy, s.qz, s.qw, s.wx, s.wy, s.wz, dt);
    }
  }
  // back to base-frame origin
  s.px = cx; s.py = cy; s.pz = cz; s.vx = ux; s.vy = uy; s.vz = uz;
  if (has_rc) {
    const Mat3 R = ds_rot(s.qx, s.qy, s.qz, s.qw, 2.0f);
    float ox, oy, oz;
    rot_rc(R, ox, oy, oz);
    s.px -= ox; s.py -= oy; s.pz -= oz;
    rot_wxrc(R, ox, oy, oz);
    s.vx -= ox; s.vy -= oy; s.vz -= oz;
  }
  prev_rpm_sum = rpm_sum;
}
