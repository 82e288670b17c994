// INDI control laws, one vehicle per thread, FP32 (FP64 only inside the WLS slow path).
//
// Restates dronesim/control/INDIControl.py:232-490 (quad / 4 virtual controls) and
// dronesim/control/INDIControl_6DOF.py:341-634 (hexa / 6 virtual controls + WLS allocation).
//
// Algebraic shortcuts (results equal to the reference's within FP32 rounding):
//  * pinv(G) (INDIControl.py:319-339).  With r1,r2,r3 the columns of the ZYX rotation matrix the
//    reference's G is [-T r2 | T cos(roll) r1 | r3], so  G^-1 e = (-(r2.e)/T, (r1.e)/(T cos(roll)), r3.e):
//    three dot products instead of a 3x3 SVD.  (np.linalg.pinv only differs from the inverse below
//    a 1e-15 relative singular value, i.e. |cos(roll)| < 1e-15.)
//  * pinv(G1/0.05) (INDIControl.py:459) and the first-iteration WLS matrix (wls_alloc.py:190-259)
//    are per-type constants precomputed on the host in FP64 (DsTypeDev::alloc).
#pragma once
#include "ds_device.cuh"
#include "ds_wls.cuh"

#define DS_WLS_MARGIN 1.0e-3f

struct CtrlState {   // kinematic state the controller sees
  float px, py, pz;
  float qx, qy, qz, qw;
  float vx, vy, vz;
  float wx, wy, wz;  // BODY rates (the reference rotates the world rates first, INDIControl.py:428-430)
};
struct CtrlTarget { float x, y, z, yaw, vx, vy, vz, ax, ay, az; };
struct CtrlMem { float lvx, lvy, lvz, lrx, lry, lrz, lthrust; float cmd[6]; float afx, afy, afz; };
struct CtrlOut { float pex, pey, pez, yaw_err; int wls_iter; int sat; };

// cmd = clip(cmd + A nu) (INDIControl.py:459, 486-487), two rotors per packed instruction.  Rotors beyond n_u have zero
// rows and zero limits: their command stays 0, so the loop needs no per-rotor guard.
template <bool NU6, int NV>
__device__ __forceinline__ void ds_allocate(const DsTypeDev& tp, const float* nu, CtrlMem& m, CtrlOut& o, const f2* du_in = nullptr) {
  constexpr int NP = NU6 ? 3 : 2;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    f2 du;
    if (du_in) {
      du = du_in[p];
    } else {
      du = ld2(tp.alloc2[p][0]) * nu[0];
#pragma unroll
      for (int j = 1; j < NV; ++j) du = ds_fma(ld2(tp.alloc2[p][j]), nu[j], du);
    }
    const f2 c = f2_make(m.cmd[2 * p], m.cmd[2 * p + 1]) + du;
    const f2 cc = ds_min(ds_max(c, ld2(tp.plo[p])), ld2(tp.phi[p]));
    o.sat += (f2_lo(cc) != f2_lo(c)) + (f2_hi(cc) != f2_hi(c));
    m.cmd[2 * p] = f2_lo(cc); m.cmd[2 * p + 1] = f2_hi(cc);
  }
}
template <bool NU6>
__device__ __forceinline__ void ds_allocate_quad(const DsTypeDev& tp, const float nu[4], CtrlMem& m, CtrlOut& o) {
  ds_allocate<NU6, 4>(tp, nu, m, o);
}

// rate loop shared by both laws: returns nu[0..2] and updates last_rates (INDIControl.py:428-453).
// EXT: first-order low-pass on the angular-acceleration estimate (north_star; the reference's filter is a commented
// placeholder, INDIControl.py:432-439): a_f += b (a_raw - a_f), b >= 1 = off.
template <bool EXT>
__device__ __forceinline__ void ds_rate_loop(const DsTypeDev& tp, const CtrlState& s, float inv_dt, float acc_b, float rsp_p,
                                             float rsp_q, float rsp_r, CtrlMem& m, float nu[3]) {
  float aax = (s.wx - m.lrx) * inv_dt, aay = (s.wy - m.lry) * inv_dt, aaz = (s.wz - m.lrz) * inv_dt;
  m.lrx = s.wx; m.lry = s.wy; m.lrz = s.wz;
  if (EXT) {
    if (acc_b < 1.f) {
      aax = fmaf(acc_b, aax - m.afx, m.afx); aay = fmaf(acc_b, aay - m.afy, m.afy); aaz = fmaf(acc_b, aaz - m.afz, m.afz);
    }
    m.afx = aax; m.afy = aay; m.afz = aaz;
  }
  nu[0] = (rsp_p - s.wx) * tp.rate[0] - aax;
  nu[1] = (rsp_q - s.wy) * tp.rate[1] - aay;
  nu[2] = (rsp_r - s.wz) * tp.rate[2] - aaz;
}

// Where a lane whose first WLS iterate is infeasible (wls_alloc.py:264) sends its problem when the FP64 active-set loop
// does not run inside the calling kernel (DEFER): ds_wls_fixup_kernel solves the queued problems right after the step
// kernel.  Keeping that loop (a 2.5 KB stack frame, 400 DFMA, a call) out of the fused kernel is worth 2.4 % on the
// mixed swarm and 9 % on the hexa-only workload even though it never executes there (profiles/r01_notes.md).
struct WlsQueue {
  int* count;    // entries queued by this launch
  int* index;    // [n] vehicle ids
  float* nu;     // [n][6] virtual controls of the queued vehicles (indexed by vehicle id)
  int vehicle;   // this lane's vehicle id, < 0: the lane must not queue (tile padding)
};

template <bool NU6, bool EXT, bool DEFER = false>
__device__ __forceinline__ void ds_indi_control(const DsTypeDev& tp, const DsWlsDev* __restrict__ wls_tab, int type_id,
                                                const CtrlState& s, const CtrlTarget& t, float inv_dt, float acc_b, CtrlMem& m,
                                                CtrlOut& o, bool want_yaw_err, const WlsQueue* wq = nullptr) {
  // ---- position loop (INDIControl.py:278-296 / INDIControl_6DOF.py:390-413)
  o.pex = t.x - s.px; o.pey = t.y - s.py; o.pez = t.z - s.pz;
  float asx = (o.pex * tp.kp + t.vx - s.vx) * tp.kd;
  float asy = (o.pey * tp.kp + t.vy - s.vy) * tp.kd;
  float asz = (o.pez * tp.kp + t.vz - s.vz) * tp.kd;
  float cax = (s.vx - m.lvx) * inv_dt, cay = (s.vy - m.lvy) * inv_dt, caz = (s.vz - m.lvz) * inv_dt;
  m.lvx = s.vx; m.lvy = s.vy; m.lvz = s.vz;
  const bool six = NU6 && (tp.law == 1);  // (a 6-DOF type makes every kernel of the handle an NU6 one)
  float tax = six ? 0.f : t.ax, tay = six ? 0.f : t.ay, taz = six ? 0.f : t.az;  // 6DOF ignores target_acc (:410)
  float aex = ds_clampf(asx + tax - cax, -6.f, 6.f);
  float aey = ds_clampf(asy + tay - cay, -6.f, 6.f);
  float aez = ds_clampf(asz + taz - caz, -6.f, 6.f);

  // ---- attitude: Euler angles and the rotation matrix the reference's G is built from
  const float x = s.qx, y = s.qy, z = s.qz, w = s.qw;
  float d = x * x + y * y + z * z + w * w;
  Mat3 R = ds_rot(x, y, z, w, 2.0f * ds_rcp(d));  // |q|^2 = 1 up to rounding: the 1-ulp MUFU reciprocal is exact enough
  float sarg = -2.0f * (x * z - w * y);
  const bool gimbal = (sarg <= -DS_GIMBAL) || (sarg >= DS_GIMBAL);
  float A_r = 2.0f * (y * z + w * x), B_r = w * w - x * x - y * y + z * z;
  float A_y = 2.0f * (x * y + w * z), B_y = w * w + x * x - y * y - z * z;
  float cphi, cpsi, spsi;
  Mat3 Re = R;  // rotation matrix of the extracted Euler angles (== R unless gimbal branch)
  float phi = 0.f, theta = 0.f, psi = 0.f;
  const bool need_angles = (!six) || want_yaw_err || gimbal;
#ifdef DS_EXACT_CTRL_TRIG
  if (need_angles) ds_euler<false>(x, y, z, w, phi, theta, psi);
#else
  if (need_angles) ds_euler<true>(x, y, z, w, phi, theta, psi);
#endif
  if (!gimbal) {
    cphi = B_r * rsqrtf(A_r * A_r + B_r * B_r);
    float ny = rsqrtf(A_y * A_y + B_y * B_y);
    cpsi = B_y * ny; spsi = A_y * ny;
  } else {  // roll = 0, pitch = +-pi/2: rebuild the matrix from the branch's angles
    float sth, cth;  // MUFU sin / cos: this rare branch must not drag libm's argument-reduction slow path into the kernel
    __sincosf(theta, &sth, &cth);
    __sincosf(psi, &spsi, &cpsi);
    cphi = 1.f;
    Re.m00 = cth * cpsi; Re.m10 = cth * spsi; Re.m20 = -sth;
    Re.m01 = -spsi;      Re.m11 = cpsi;       Re.m21 = 0.f;
    Re.m02 = sth * cpsi; Re.m12 = sth * spsi; Re.m22 = cth;
  }
  // G^-1 accel_e
  const float T = 9.81f;  // INDIControl.py:314 (quirk Q2)
  float u0 = Re.m00 * aex + Re.m10 * aey + Re.m20 * aez;
  float u1 = Re.m01 * aex + Re.m11 * aey + Re.m21 * aez;
  float u2 = Re.m02 * aex + Re.m12 * aey + Re.m22 * aez;
  float dT = u2;
  float thrust = m.lthrust + dT;  // INDIControl.py:347 / INDIControl_6DOF.py:491
  o.wls_iter = 0;

  // rate set-points (attitude loop) and the translational virtual controls of the lane's law
  float nu[6];
  float rsp0, rsp1, rsp2;
  if (!six) {
    float dphi = -u1 * (1.0f / T);
    float dtheta = u0 * ds_rcp(T * cphi);
    float yaw_inc = ds_norm_ang(t.yaw - psi);  // :341
    float te_r = phi + dphi, te_p = theta + dtheta, te_y = psi + yaw_inc;
    o.yaw_err = te_y - psi;  // :227
    // ---- attitude loop (INDIControl.py:388-402)
#ifdef DS_EXACT_CTRL_TRIG
    float4 tq = ds_quat_from_euler<false>(te_r, te_p, te_y);
#else
    float4 tq = ds_quat_from_euler<true>(te_r, te_p, te_y);
#endif
    float ew = w * tq.w + x * tq.x + y * tq.y + z * tq.z;  // utils/math.py:23-31
    float ex = w * tq.x - x * tq.w - y * tq.z + z * tq.y;
    float ey = w * tq.y + x * tq.z - y * tq.w - z * tq.x;
    float ez = w * tq.z - x * tq.y + y * tq.x - z * tq.w;
    if (ew < 0.f) { ex = -ex; ey = -ey; ez = -ez; }  // quat_wrap_shortest, in place (quirk Q1)
    rsp0 = tp.att[0] * ex; rsp1 = tp.att[1] * ey; rsp2 = tp.att[2] * ez;
    nu[3] = dT;  // thrust - last_thrust (:454); dT avoids the FP32 cancellation of (lt + dT) - lt
    nu[4] = nu[5] = 0.f;
  } else {
    o.yaw_err = 0.f - psi;  // target_euler = 0 (:495)
    // attitude error: conj(q) (x) identity = (-x,-y,-z,w), no shortest wrap (:540-545)
    float e0 = -x, e1 = -y, e2 = -z;
    float r0 = cpsi * e0 + spsi * e1;  // inv(R_psi) (:551-557)
    float r1 = -spsi * e0 + cpsi * e1;
    rsp0 = tp.att[0] * r0; rsp1 = tp.att[1] * r1; rsp2 = tp.att[2] * e2;
    // accel_error_body = R^T accel_e (:589) - uses the quaternion's own matrix, not the Euler one
    nu[3] = R.m00 * aex + R.m10 * aey + R.m20 * aez;
    nu[4] = R.m01 * aex + R.m11 * aey + R.m21 * aez;
    nu[5] = R.m02 * aex + R.m12 * aey + R.m22 * aez;
  }
  m.lthrust = thrust;  // INDIControl.py:456 / INDIControl_6DOF.py:598
  // ---- rate loop and allocation, ONCE for both laws: a warp that mixes quads and hexas runs them together instead of
  // once per branch.  A quad's allocation matrix has zero columns 4, 5 and its nu[4] = nu[5] = 0, so the six-column
  // product below is its pinv(G1 / 0.05) nu (INDIControl.py:459) exactly.
  ds_rate_loop<EXT>(tp, s, inv_dt, acc_b, rsp0, rsp1, rsp2, m, nu);
  if constexpr (!NU6) {
    ds_allocate_quad<NU6>(tp, nu, m, o);
  } else {
    // 6-DOF law: first WLS iteration in closed form, du = M nu (rotor pairs on packed instructions)
    f2 du2[3];
    bool feasible = true;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      f2 d = ld2(tp.alloc2[p][0]) * nu[0];
#pragma unroll
      for (int j = 1; j < 6; ++j) d = ds_fma(ld2(tp.alloc2[p][j]), nu[j], d);
      du2[p] = d;
      // wls_alloc.py:264 decides u_opt >= umax + 1 or u_opt <= umin - 1 in FP64; the FP32 first iterate is trusted only
      // when it clears the thresholds by DS_WLS_MARGIN (>> its rounding error), anything closer goes to the FP64
      // active-set routine, whose own first pass repeats the reference's test exactly
      const f2 c = f2_make(m.cmd[2 * p], m.cmd[2 * p + 1]);
      const f2 hi = (ld2(tp.phi[p]) + (1.0f - DS_WLS_MARGIN)) - c, lo = (ld2(tp.plo[p]) - (1.0f - DS_WLS_MARGIN)) - c;
      feasible = feasible && (f2_lo(d) < f2_lo(hi)) && (f2_lo(d) > f2_lo(lo)) && (f2_hi(d) < f2_hi(hi)) && (f2_hi(d) > f2_hi(lo));
    }
    feasible = feasible || !six;  // the quad law clips, it has no feasibility notion (INDIControl.py:486-487)
    if (six) o.wls_iter = 1;
    if (DEFER) {
      if (!feasible) {  // rare: hold the command now, ds_wls_fixup_kernel applies the active-set solution
        o.wls_iter = 2;
        if (wq->vehicle >= 0) {
          const int qi = atomicAdd(wq->count, 1);
          wq->index[qi] = wq->vehicle;
          float* dst = wq->nu + (size_t)wq->vehicle * 6;
#pragma unroll
          for (int i = 0; i < 6; ++i) dst[i] = nu[i];
        }
#pragma unroll
        for (int p = 0; p < 3; ++p) du2[p] = f2_make(0.f, 0.f);
      }
    } else if (!feasible) {  // rare: run the active-set iterations in FP64
      const DsWlsDev* P = wls_tab + type_id;
      double v[6], umin[6], umax[6], u[6];
      for (int i = 0; i < 6; ++i) {
        v[i] = (double)nu[i];
        umin[i] = P->pmin[i] - (double)m.cmd[i];
        umax[i] = P->pmax[i] - (double)m.cmd[i];
        u[i] = 0.0;
      }
      int it = ds_wls_alloc(P, v, umin, umax, u);
      o.wls_iter = it;
      for (int p = 0; p < 3; ++p)  // non-convergence: hold the command
        du2[p] = (it > 0) ? f2_make((float)u[2 * p], (float)u[2 * p + 1]) : f2_make(0.f, 0.f);
    }
    ds_allocate<true, 6>(tp, nu, m, o, du2);  // cmd = clip(cmd + du) (:630-631 / INDIControl.py:486-487)
  }
}
