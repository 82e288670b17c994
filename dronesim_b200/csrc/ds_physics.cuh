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
// The kernel is bound by instruction issue (ncu: issue slots 81 % busy, FP32 FMA pipe 53 %, DRAM 19 %), and two thirds
// of its instructions are FP32.  sm_100 has packed FP32 instructions (FFMA2 / FMUL2 / FADD2: two operations per issue
// slot, same FMA-pipe cycles), so the substep is written on register PAIRS (f2, ds_lanes.cuh) wherever two operations
// share a shape:
//  * vectors are (x, y) pairs + a scalar z; the rotation matrix is three column pairs + its third row, so R v is
//    3 FFMA2 + 3 FFMA instead of 9 FFMA, and the same for J w and J^-1 g (constants stored in that layout);
//  * the body wrench lives in three pairs (Fx, Fy) (Fz, tx) (ty, tz): every rotor adds its ground-effect force and
//    torque with 3 FFMA2 (broadcast magnitude) instead of 6 FFMA; the heights of TWO rotors come from 3 FFMA2;
//  * two downwash partners are evaluated per packed instruction;
//  * every multiply-add is spelled as ds_fma: the r01 kernel left 34 % FMUL + 18 % FADD unfused.
// What has no packed form (MUFU, min / max, compares, selects) runs per half on the aliased registers.
// Unchanged from round 1:
//  * the command is constant across the K substeps of a control step (BaseAviary.py:507-545), so
//    the rotor wrench sum_i T_i a_i, sum_i T_i m_i is hoisted out of the substep loop; only the terms
//    that depend on the moving state (ground effect, drag, downwash, gyroscopic torque) are per substep;
//  * reciprocal / exp2 / rsqrt are the single-instruction MUFU forms (ds_rcp, ds_ex2, ds_rsqrt);
//  * the Gaussian's 0.5 log2(e) is folded into the per-type beta coefficients, DW_COEFF_1 is applied once after the
//    neighbour loop;
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


// Rotation matrix of a unit quaternion (btMatrix3x3::setRotation, s = 2) as column pairs + third row:
// c0 = (m00, m10), c1 = (m01, m11), c2 = (m02, m12), r = (m20, m21, m22).  16 FMA-pipe instructions.
struct RotP { f2 c0, c1, c2; float r0, r1, r2; };
__device__ __forceinline__ RotP ds_rotp(f2 qxy, f2 qzw) {
  const float x = f2_lo(qxy), y = f2_hi(qxy), z = f2_lo(qzw), w = f2_hi(qzw);
  const f2 d2 = qxy + qxy;                   // (2x, 2y)
  const float x2 = f2_lo(d2), y2 = f2_hi(d2), z2 = z + z;
  const f2 xz = qxy * z2;                    // (x 2z, y 2z)
  const f2 sq = qxy * d2;                    // (2xx, 2yy)
  const float wz2 = w * z2;
  const float tz = ds_fma(ds_neg(z), z2, 1.0f);
  RotP R;
  R.c0 = f2_make(tz - f2_hi(sq), ds_fma(x, y2, wz2));
  R.c1 = f2_make(ds_fma(x, y2, ds_neg(wz2)), tz - f2_lo(sq));
  R.c2 = f2_make(ds_fma(w, y2, f2_lo(xz)), ds_fma(ds_neg(w), x2, f2_hi(xz)));
  R.r0 = ds_fma(ds_neg(w), y2, f2_lo(xz));
  R.r1 = ds_fma(w, x2, f2_hi(xz));
  R.r2 = (1.0f - f2_lo(sq)) - f2_hi(sq);
  return R;
}
// acc + R v
__device__ __forceinline__ void ds_rot_fma(const RotP& R, float vx, float vy, float vz, f2& axy, float& az) {
  axy = ds_fma(R.c0, vx, ds_fma(R.c1, vy, ds_fma(R.c2, vz, axy)));
  az = ds_fma(R.r0, vx, ds_fma(R.r1, vy, ds_fma(R.r2, vz, az)));
}

// q <- normalize(q (x) exp(w dt)).  Taylor polynomials of 0.5 sin(h)/h and cos(h) in h^2 = (half angle)^2 < 0.25 (the
// first dropped terms are 5e-9 and 3e-10 relative at h^2 = 0.25); a.qk = TIMESTEP x the sine coefficients, a.qh = 0.25 dt^2.
// |w| dt >= 1 rad per substep (>= 240 rad/s): a blown-up state; MUFU sin / cos keep it finite.  Out of line on purpose:
// inlined, ptxas predicates these ten instructions into every substep instead of branching around them.
static __device__ __noinline__ float2 ds_quat_step_fast_spin(float h2, float dt) {  // -> (k, c)
  const float half = sqrtf(h2);
  float sn, cs;
  __sincosf(half, &sn, &cs);
  return make_float2(0.5f * dt * sn * ds_rcp(half), cs);
}

__device__ __forceinline__ void ds_quat_step(const DsArgs& a, f2& qxy, f2& qzw, f2 wxy, float wz) {
  const f2 w2 = wxy * wxy;
  const float h2 = ds_fma(wz, wz, f2_lo(w2) + f2_hi(w2)) * a.qh;
  float k = ds_fma(h2, ds_fma(h2, ds_fma(h2, a.qk[3], a.qk[2]), a.qk[1]), a.qk[0]);
  float c = ds_fma(h2, ds_fma(h2, ds_fma(h2, ds_fma(h2, 1.0f / 40320.0f, -1.0f / 720.0f), 1.0f / 24.0f), -0.5f), 1.0f);
  if (h2 >= 0.25f) { const float2 kc = ds_quat_step_fast_spin(h2, a.dt); k = kc.x; c = kc.y; }
  const f2 dxy = wxy * k;
  const float dx = f2_lo(dxy), dy = f2_hi(dxy), dz = wz * k;
  const float qx = f2_lo(qxy), qy = f2_hi(qxy), qz = f2_lo(qzw), qw = f2_hi(qzw);
  // (nx, ny) = qw (dx, dy) + c (qx, qy) + (qy dz - qz dy, qz dx - qx dz)
  f2 nxy = ds_fma(dxy, qw, qxy * c);
  nxy = f2_make(ds_fma(qy, dz, ds_fma(ds_neg(qz), dy, f2_lo(nxy))), ds_fma(qz, dx, ds_fma(ds_neg(qx), dz, f2_hi(nxy))));
  // (nz, nw) = c (qz, qw) + (qw dz + qx dy - qy dx, -qz dz - qx dx - qy dy)
  f2 nzw = qzw * c;
  nzw = f2_make(ds_fma(qw, dz, ds_fma(qx, dy, ds_fma(ds_neg(qy), dx, f2_lo(nzw)))),
                ds_fma(ds_neg(qz), dz, ds_fma(ds_neg(qx), dx, ds_fma(ds_neg(qy), dy, f2_hi(nzw)))));
  const f2 n2 = ds_fma(nzw, nzw, nxy * nxy);
  const float n = ds_rsqrt(f2_lo(n2) + f2_hi(n2));
  qxy = nxy * n;
  qzw = nzw * n;
}

// ---------------------------------------------------------------------------------------------------------------------
// Downwash of every drone of the env on this one (BaseAviary.py:1747-1763), in units of DW_COEFF_1
// (PROP_RADIUS / 4)^2: sum_j [dz > 0, dxy < 10] exp(-0.5 (dxy / beta)^2) / dz^2, beta = DW2 dz + DW3.
// row: the env's position snapshot in shared memory; k2, k3 are pre-divided by sqrt(0.5 log2 e).
// dz > 0 also drops the drone itself.  beta == 0 (dz = -DW3 / DW2 exactly) gives 1/beta = inf -> exp2(-inf) = 0,
// the reference's exp(-inf) (BaseAviary.py:1755), with no extra test.
// Two partners per evaluation: the six differences are scalar subtractions written into pairs, the rest is packed.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ds_dw_gate(float& acc, float dz, float d2, float w) {
  asm("{\n\t.reg .pred p;\n\t"
      "setp.gt.f32 p, %1, 0f00000000;\n\t"
      "setp.lt.and.f32 p, %2, 0f42C80000, p;\n\t"  // dxy^2 < 100
      "@p add.f32 %0, %0, %3;\n\t}"
      : "+f"(acc)
      : "f"(dz), "f"(d2), "f"(w));
}
// symmetric variant: the term goes to whichever vehicle of the pair is the lower one
__device__ __forceinline__ float ds_dw_gate_sym(float& acc, float dz, float d2, float w) {
  float theirs;
  asm("{\n\t.reg .pred p, q;\n\t"
      "setp.lt.f32 q, %3, 0f42C80000;\n\t"          // dxy^2 < 100
      "setp.gt.and.f32 p, %2, 0f00000000, q;\n\t"   // partner above me: the term is mine
      "@p add.f32 %0, %0, %4;\n\t"
      "setp.lt.and.f32 p, %2, 0f00000000, q;\n\t"   // partner below me: the term is the partner's
      "selp.f32 %1, %4, 0f00000000, p;\n\t}"
      : "+f"(acc), "=f"(theirs)
      : "f"(dz), "f"(d2), "f"(w));
  return theirs;
}

// the weights exp2(-(dxy / beta')^2) / dz^2 of two partners oa, ob (before the gates); SYM: beta of the LOWER vehicle
template <bool SYM>
__device__ __forceinline__ f2 ds_dw_weight2(const float4 oa, const float4 ob, float px, float py, float pz, float k2, float k3,
                                            f2& dz, f2& d2) {
  dz = f2_make(oa.z - pz, ob.z - pz);
  const f2 dx = f2_make(oa.x - px, ob.x - px), dy = f2_make(oa.y - py, ob.y - py);
  d2 = ds_fma(dy, dy, dx * dx);
  const f2 ib = ds_rcp(SYM ? ds_fma(ds_abs(dz), k2, k3) : ds_fma(dz, k2, k3));  // 1 / beta'
  const f2 idz2 = ds_rcp(dz * dz);                                             // 1 / dz^2
  return idz2 * ds_ex2(ds_neg((ib * ib) * d2));
}
__device__ __forceinline__ float ds_dw_weight1(const float4 o, float px, float py, float pz, float k2, float k3, float& dz,
                                               float& d2) {
  dz = o.z - pz;
  const float dx = o.x - px, dy = o.y - py;
  d2 = ds_fma(dy, dy, dx * dx);
  const float ib = ds_rcp(ds_fma(dz, k2, k3));
  const float idz2 = ds_rcp(dz * dz);
  return idz2 * ds_ex2(ds_neg((ib * ib) * d2));
}

__device__ __forceinline__ float ds_downwash_sum(const float4* __restrict__ row, int D, float px, float py, float pz,
                                                 float k2, float k3) {
  float acc0 = 0.f, acc1 = 0.f;
  auto two = [&](int j) {
    f2 dz, d2;
    const f2 w = ds_dw_weight2<false>(row[j], row[j + 1], px, py, pz, k2, k3, dz, d2);
    ds_dw_gate(acc0, f2_lo(dz), f2_lo(d2), f2_lo(w));
    ds_dw_gate(acc1, f2_hi(dz), f2_hi(d2), f2_hi(w));
  };
  if (D == 16) {  // BASELINE configs[3]/[4]
#pragma unroll
    for (int j = 0; j < 16; j += 2) two(j);
  } else {
    int j = 0;
#pragma unroll 2
    for (; j + 1 < D; j += 2) two(j);
    if (j < D) {
      float dz, d2;
      const float w = ds_dw_weight1(row[j], px, py, pz, k2, k3, dz, d2);
      ds_dw_gate(acc0, dz, d2, w);
    }
  }
  return acc0 + acc1;
}

// ---- symmetric variant (D == 16, every type shares DW_COEFF_2 / DW_COEFF_3 - true of all shipped URDFs) ----------
// Only the LOWER vehicle of a pair feels the pair's downwash (dz > 0 gate), and the pair term depends on the two
// positions and (k2, k3) alone, so each of the 120 unordered pairs of an env is evaluated ONCE: in round r = 1..7
// slot s evaluates the pair (s, s + r mod 16), keeps the term if it is the lower one, otherwise hands it to the
// partner by a 16-lane-wide SHFL.IDX; round 8 pairs (s, s + 8) are evaluated from both sides and kept locally.
// Rounds (1,2) (3,4) (5,6) (7,8) share their packed arithmetic.  The env's snapshot rows are stored twice (rows s and
// s + 16 of a 32-row block) so the partner row s + r needs no wrap-around arithmetic.
#define DS_DW_SYM_ROWS 32
// row: this thread's own row (slot) in the env's 32-row block; slot16 = slot + 16
__device__ __forceinline__ float ds_downwash_sum_sym16(const float4* __restrict__ row, int slot16, float px, float py,
                                                       float pz, float k2, float k3) {
  float acc0 = 0.f, acc1 = 0.f, recv0 = 0.f, recv1 = 0.f;
#pragma unroll
  for (int r = 1; r <= 7; r += 2) {
    f2 dz, d2;
    const f2 w = ds_dw_weight2<true>(row[r], row[r + 1], px, py, pz, k2, k3, dz, d2);
    const float t0 = ds_dw_gate_sym(acc0, f2_lo(dz), f2_lo(d2), f2_lo(w));
    recv0 += __shfl_sync(0xffffffffu, t0, slot16 - r, 16);  // from slot - r (mod 16) of my env
    if (r < 7) {
      const float t1 = ds_dw_gate_sym(acc1, f2_hi(dz), f2_hi(d2), f2_hi(w));
      recv1 += __shfl_sync(0xffffffffu, t1, slot16 - r - 1, 16);
    } else {
      ds_dw_gate(acc1, f2_hi(dz), f2_hi(d2), f2_hi(w));  // round 8: evaluated from both sides, kept locally
    }
  }
  return (acc0 + acc1) + (recv0 + recv1);
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
// RC: some type of the swarm has a centre-of-mass offset (compile time: the offset arithmetic is 25 instructions per
// substep that quad-only swarms do not need).
template <int INTEG, int DW, bool NU6, bool WARPSYNC, int FX, bool EXT, bool RC>
__device__ __forceinline__ void ds_physics(const DsArgs& a, const DsTypeDev& tp, int env_row0, int my_row, float4* sh_pos,
                                           const float* act, PhysState& s, float& prev_rpm_sum, float* rpm_state,
                                           uint32_t veh_id) {
  constexpr int NU = NU6 ? 6 : 4;
  const float dt = a.dt;
  const bool gnd = (FX >= 0) ? ((FX & 1) != 0) : ((a.flags & 1u) != 0);
  const bool drag = (FX >= 0) ? ((FX & 2) != 0) : ((a.flags & 2u) != 0);
  // RC is the same for every lane: types without an offset run the same arithmetic with rc = 0, which is exact
  // (x + 0 = x), so a warp that mixes airframes does not diverge here; offsets along body z only take the general path
  constexpr bool has_rc = RC && INTEG == 0;
  constexpr bool rc_gen = has_rc;

  const float rcx = tp.rc[0], rcy = tp.rc[1], rcz = tp.rc[2];

  float roll = 0.f, pitch = 0.f, yaw = 0.f;
  if (INTEG == 1) ds_euler(s.qx, s.qy, s.qz, s.qw, roll, pitch, yaw);  // state cache rpy (BaseAviary.py:729)

  // state as pairs: centre-of-mass position c, velocity u (QUAT: of the centre of mass), body rates w, quaternion
  f2 cxy = f2_make(s.px, s.py), uxy = f2_make(s.vx, s.vy), wxy = f2_make(s.wx, s.wy);
  f2 qxy = f2_make(s.qx, s.qy), qzw = f2_make(s.qz, s.qw);
  float cz = s.pz, uz = s.vz, wz = s.wz;

  // acc + sign R (w x rc): sign = -1 gives the velocity of the base origin from the centre-of-mass velocity
  auto add_rot_wxrc = [&](const RotP& R, float sign, f2& vxy, float& vz) {
    const float zc = sign * rcz;
    float kx = f2_hi(wxy) * zc, ky = f2_lo(wxy) * ds_neg(zc);  // sign (w x rc), rc along z
    if (rc_gen) {
      const float xc = sign * rcx, yc = sign * rcy;
      kx = ds_fma(wz, ds_neg(yc), kx); ky = ds_fma(wz, xc, ky);
      const float kz = ds_fma(f2_lo(wxy), yc, ds_neg(f2_hi(wxy) * xc));
      ds_rot_fma(R, kx, ky, kz, vxy, vz);
    } else {
      vxy = ds_fma(R.c0, kx, ds_fma(R.c1, ky, vxy));
      vz = ds_fma(R.r0, kx, ds_fma(R.r1, ky, vz));
    }
  };
  // acc + sign R rc
  auto add_rot_rc = [&](const RotP& R, float sign, f2& pxy, float& pz) {
    pxy = ds_fma(R.c2, sign * rcz, pxy); pz = ds_fma(R.r2, sign * rcz, pz);
    if (rc_gen) {
      pxy = ds_fma(R.c0, sign * rcx, ds_fma(R.c1, sign * rcy, pxy));
      pz = ds_fma(R.r0, sign * rcx, ds_fma(R.r1, sign * rcy, pz));
    }
  };

  if (has_rc) {
    const RotP R = ds_rotp(qxy, qzw);
    add_rot_rc(R, 1.0f, cxy, cz);
    add_rot_wxrc(R, 1.0f, uxy, uz);
  }
  // ---- rotor thrusts and the rotor part of the body wrench: once per control step (the command is constant across
  // the substeps), or once per substep when the motor model moves the rotor speeds.  The wrench is kept as the three
  // pairs (Fx, Fy) (Fz, tx) (ty, tz).
  f2 Tg[NU / 2];  // T_i * GND_EFF_COEFF (PROP_RADIUS/4)^2 of two rotors: the only per-rotor value the substeps need
  float rpm_sum = 0.f;
  f2 W0_0, W1_0, W2_0;
  auto rotor_wrench = [&](int k) {
    float F0x = 0.f, F0y = 0.f, F0z = 0.f, t0x = 0.f, t0y = 0.f, t0z = 0.f, tg[NU];
    rpm_sum = 0.f;
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
      const RotP Ra = ds_rotp(qxy, qzw);
      f2 vxy = uxy;
      float vz = uz;
      if (has_rc) add_rot_wxrc(Ra, -1.0f, vxy, vz);
      float vx = f2_lo(vxy), vy = f2_hi(vxy);
      const float V = sqrtf(vx * vx + vy * vy + vz * vz);
      if (!(V > 0.1f)) { vx = 0.1f; vy = 0.f; vz = 0.f; }  // :1585-1589
      f2 bxy = f2_make(0.f, 0.f);
      float bz = 0.f;
      ds_rot_fma(Ra, vx, vy, vz, bxy, bz);
      const float bx = f2_lo(bxy), by = f2_hi(bxy);
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
      tg[i] = T * tp.gnd_k;
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
#pragma unroll
    for (int p = 0; p < NU / 2; ++p) Tg[p] = f2_make(tg[2 * p], tg[2 * p + 1]);
    W0_0 = f2_make(F0x, F0y); W1_0 = f2_make(F0z, t0x); W2_0 = f2_make(t0y, t0z);
  };
  if (!EXT) rotor_wrench(0);
  // drag coefficient x rotor speed sum: the first substep still sees the previously applied action (:532,:545)
  const f2 ndk_xy = ld2(tp.ndk_xy);
  const float ndk_z = tp.ndk_z;

  for (int k = 0; k < a.K; ++k) {
    if (EXT) {  // drag sees the rotor speeds before this substep's motor update
      if (k > 0) prev_rpm_sum = rpm_sum;
      rotor_wrench(k);
    }
    const RotP R = ds_rotp(qxy, qzw);
    // ---- downwash (BaseAviary.py:1747-1763): every drone of the env reads the same position snapshot
    float dw_fz = 0.f;
    if (DW) {
      f2 pxy = cxy;  // base-frame origin p = c - R rc
      float pz = cz;
      if (has_rc) add_rot_rc(R, -1.0f, pxy, pz);
      const float px = f2_lo(pxy), py = f2_hi(pxy);
      float dsum;
      float4* buf = sh_pos + (k & 1) * DS_DW_BUF;
      // rows are float4 (x, y, z, -): stored as the (x, y) register pair + z, so no four-register tuple has to be assembled
      auto put = [&](int row) {
        *reinterpret_cast<float2*>(&buf[row]) = make_float2(px, py);
        reinterpret_cast<float*>(&buf[row])[2] = pz;
      };
      if (DW == 2) {  // symmetric pairs: env_row0 / my_row index 32-row blocks, slot = my_row - env_row0
        put(my_row);
        if (my_row - env_row0 < 8) put(my_row + 16);
        __syncwarp();
        dsum = ds_downwash_sum_sym16(buf + my_row, my_row - env_row0 + 16, px, py, pz, tp.dw_k2, tp.dw_k3);
      } else {
        put(my_row);
        if (WARPSYNC) __syncwarp(); else __syncthreads();
        dsum = ds_downwash_sum(buf + env_row0, a.D, px, py, pz, tp.dw_k2, tp.dw_k3);
      }
      dw_fz = dsum * tp.dw_k1n;
    }

    // body wrench (Fx, Fy) (Fz, tx) (ty, tz): rotors + downwash force at the base origin, torque (-rc) x (0, 0, f)
    f2 W0 = W0_0, W1 = W1_0, W2 = W2_0;
    if (DW) {
      if (rc_gen) { W1 = ds_fma(f2_make(1.0f, ds_neg(rcy)), dw_fz, W1); W2 = f2_make(ds_fma(dw_fz, rcx, f2_lo(W2)), f2_hi(W2)); }
      else W1 = f2_make(f2_lo(W1) + dw_fz, f2_hi(W1));
    }

    if (gnd) {  // BaseAviary.py:1672-1699; rotor sites are stored relative to the centre of mass
      // gate (|roll|, |pitch| < pi/2): a vehicle outside it gets "infinite" rotor heights -> 1/h = 0 (flushed), g = 0
      float czg;
      if (INTEG == 1) czg = ((fabsf(roll) < 0.5f * DS_PI_F) && (fabsf(pitch) < 0.5f * DS_PI_F)) ? cz : 3.0e38f;
      else czg = ((R.r2 > 0.f) && (ds_abs(R.r0) < DS_GIMBAL)) ? cz : 3.0e38f;  // |roll| < pi/2 <=> cos(roll)cos(pitch) > 0
#pragma unroll
      for (int p = 0; p < NU / 2; ++p) {
        const f2 h = ds_fma(ld2(tp.gh[p][0]), R.r0, ds_fma(ld2(tp.gh[p][1]), R.r1, ds_fma(ld2(tp.gh[p][2]), R.r2, czg)));
        const f2 ih = ds_rcp(ds_max(h, tp.gnd_clip));
        const f2 g = Tg[p] * (ih * ih);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float gi = j ? f2_hi(g) : f2_lo(g);
          const float2* c = tp.gw[2 * p + j];
          W0 = ds_fma(ld2(c[0]), gi, W0); W1 = ds_fma(ld2(c[1]), gi, W1); W2 = ds_fma(ld2(c[2]), gi, W2);
        }
      }
    }
    if (drag) {  // BaseAviary.py:1719-1732
      f2 vxy = uxy;
      float vz = uz;
      if (has_rc) add_rot_wxrc(R, -1.0f, vxy, vz);  // velocity of the base origin
      const float sum = (EXT || k == 0) ? prev_rpm_sum : rpm_sum;
      const f2 dxy = (ndk_xy * sum) * vxy;
      const float d2 = (ndk_z * sum) * vz;
      f2 fxy = R.c2 * d2;
      float fz = R.r2 * d2;
      fxy = ds_fma(R.c0, f2_lo(dxy), ds_fma(R.c1, f2_hi(dxy), fxy));
      fz = ds_fma(R.r0, f2_lo(dxy), ds_fma(R.r1, f2_hi(dxy), fz));
      W0 += fxy;
      float Fz = f2_lo(W1) + fz, tx = f2_hi(W1), ty = f2_lo(W2), tz = f2_hi(W2);
      if (has_rc) {  // (-rc) x f
        const float fx = f2_lo(fxy), fy = f2_hi(fxy);
        tx = ds_fma(fy, rcz, tx); ty = ds_fma(fx, ds_neg(rcz), ty);
        if (rc_gen) { tx = ds_fma(fz, ds_neg(rcy), tx); ty = ds_fma(fz, rcx, ty); tz = ds_fma(fx, rcy, ds_fma(fy, ds_neg(rcx), tz)); }
      }
      W1 = f2_make(Fz, tx); W2 = f2_make(ty, tz);
    }

    // ---- Newton-Euler + semi-implicit Euler (BaseAviary.py:1790-1812): u += (dt / m) R F - dt g, w += dt J^-1 (tau - w x J w)
    {
      f2 axy = f2_make(0.f, 0.f);
      float az = 0.f;
      const float Fx = f2_lo(W0), Fy = f2_hi(W0), Fz = f2_lo(W1);
      axy = ds_fma(R.c0, Fx, ds_fma(R.c1, Fy, R.c2 * Fz));
      az = ds_fma(R.r0, Fx, ds_fma(R.r1, Fy, R.r2 * Fz));
      uxy = ds_fma(axy, tp.dtm, uxy);
      uz = ds_fma(az, tp.dtm, uz) - a.dtg;
    }
    {
      const float wx = f2_lo(wxy), wy = f2_hi(wxy);
      const f2 jxy = ds_fma(ld2(tp.Jc[0]), wx, ds_fma(ld2(tp.Jc[1]), wy, ld2(tp.Jc[2]) * wz));
      const float jz = ds_fma(tp.Jr[0], wx, ds_fma(tp.Jr[1], wy, tp.Jr[2] * wz));
      const float jx = f2_lo(jxy), jy = f2_hi(jxy);
      const float gx = ds_fma(wz, jy, ds_fma(ds_neg(wy), jz, f2_hi(W1)));
      const float gy = ds_fma(wx, jz, ds_fma(ds_neg(wz), jx, f2_lo(W2)));
      const float gz = ds_fma(wy, jx, ds_fma(ds_neg(wx), jy, f2_hi(W2)));
      wxy = ds_fma(ld2(tp.Jdc[0]), gx, ds_fma(ld2(tp.Jdc[1]), gy, ds_fma(ld2(tp.Jdc[2]), gz, wxy)));
      wz = ds_fma(tp.Jdr[0], gx, ds_fma(tp.Jdr[1], gy, ds_fma(tp.Jdr[2], gz, wz)));
    }
    cxy = ds_fma(uxy, dt, cxy);
    cz = ds_fma(uz, dt, cz);
    if (FX < 0 && (a.flags & 64u)) {  // DS_FLAG_GROUND_PLANE: inelastic, frictionless stop (plane.urdf, BaseAviary.py:679-680)
      if (cz < a.floor_z) { cz = a.floor_z; uz = fmaxf(uz, 0.f); }
    }
    if (INTEG == 1) {
      roll = fmaf(dt, f2_lo(wxy), roll); pitch = fmaf(dt, f2_hi(wxy), pitch); yaw = fmaf(dt, wz, yaw);
      float4 q = ds_quat_from_euler(roll, pitch, yaw);  // :1817
      qxy = f2_make(q.x, q.y); qzw = f2_make(q.z, q.w);
      ds_euler(q.x, q.y, q.z, q.w, roll, pitch, yaw);  // state refresh (:729)
    } else {
      ds_quat_step(a, qxy, qzw, wxy, wz);
    }
  }
  // back to base-frame origin
  if (has_rc) {
    const RotP R = ds_rotp(qxy, qzw);
    add_rot_rc(R, -1.0f, cxy, cz);
    add_rot_wxrc(R, -1.0f, uxy, uz);
  }
  s.px = f2_lo(cxy); s.py = f2_hi(cxy); s.pz = cz;
  s.vx = f2_lo(uxy); s.vy = f2_hi(uxy); s.vz = uz;
  s.wx = f2_lo(wxy); s.wy = f2_hi(wxy); s.wz = wz;
  s.qx = f2_lo(qxy); s.qy = f2_hi(qxy); s.qz = f2_lo(qzw); s.qw = f2_hi(qzw);
  prev_rpm_sum = rpm_sum;
}
