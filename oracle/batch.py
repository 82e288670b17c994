"""Vectorised (numpy, FP64) restatement of the same path ``oracle/sim.py`` steps vehicle by vehicle.

TEST / BENCH INFRASTRUCTURE (see ``oracle/__init__.py``): NOT the reference's own code - the reference loops over
drones in Python (BaseAviary.py:522) and has no batched form.  This module exists for two reasons:

* a fairer CPU bound next to the per-vehicle port (``bench.py`` ``cpu_baseline.vectorised``, SURVEY.md 8d ii);
* an oracle fast enough to check the GPU core on thousands of environments (``bench.py`` ``parity_check``,
  ``tests/test_gpu_parity.py``).

It follows ``oracle/dynamics.py`` (``body_wrench`` + ``substep_quat``: BaseAviary.py:1487-1543, 1398-1457, 1648-1763
with repairs R1-R8) and ``oracle/control.py`` (``QuadINDI`` = INDIControl.py:232-490, ``Hexa6DOFINDI`` =
INDIControl_6DOF.py:341-634) formula by formula, over arrays ``[E, D, ...]``; ``tests/test_oracle_batch.py`` pins it
to the per-vehicle oracle (1e-10).  Scope: the quaternion integrator with ground effect / drag / downwash and both
control laws; no extensions (motor model, filter, noise).  The 6-DOF allocation takes the closed form of the first
``wls_alloc`` iteration (the unconstrained minimiser, wls_alloc.py:190-259) wherever that iterate is feasible
(:262-266) and calls the scalar ``oracle.control.wls_alloc`` for the vehicles where it is not.
"""
import numpy as np

from . import control as oc
from . import dynamics as od

G = od.G


# ---------------------------------------------------------------- pybullet math, batched (oracle/pyb_math.py)
def rotmat(q):
    """[..., 4] xyzw -> [..., 3, 3] (getMatrixFromQuaternion)."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    s = 2.0 / (x * x + y * y + z * z + w * w)
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0], R[..., 0, 1], R[..., 0, 2] = 1.0 - (yy + zz), xy - wz, xz + wy
    R[..., 1, 0], R[..., 1, 1], R[..., 1, 2] = xy + wz, 1.0 - (xx + zz), yz - wx
    R[..., 2, 0], R[..., 2, 1], R[..., 2, 2] = xz - wy, yz + wx, 1.0 - (xx + yy)
    return R


def euler_from_quat(q):
    """[..., 4] -> [..., 3] (getEulerFromQuaternion, gimbal branch at |sin pitch| >= 0.99999)."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    sqx, sqy, sqz, squ = x * x, y * y, z * z, w * w
    sarg = -2.0 * (x * z - w * y)
    lo, hi = sarg <= -0.99999, sarg >= 0.99999
    roll = np.arctan2(2.0 * (y * z + w * x), squ - sqx - sqy + sqz)
    pitch = np.arcsin(np.clip(sarg, -1.0, 1.0))
    yaw = np.arctan2(2.0 * (x * y + w * z), squ + sqx - sqy - sqz)
    roll = np.where(lo | hi, 0.0, roll)
    pitch = np.where(lo, -0.5 * np.pi, np.where(hi, 0.5 * np.pi, pitch))
    yaw = np.where(lo, 2.0 * np.arctan2(x, -y), np.where(hi, 2.0 * np.arctan2(-x, y), yaw))
    return np.stack([roll, pitch, yaw], axis=-1)


def quat_from_euler(rpy):
    """[..., 3] -> [..., 4] normalised xyzw (getQuaternionFromEuler)."""
    h = 0.5 * rpy
    sph, cph = np.sin(h[..., 0]), np.cos(h[..., 0])
    sth, cth = np.sin(h[..., 1]), np.cos(h[..., 1])
    sps, cps = np.sin(h[..., 2]), np.cos(h[..., 2])
    q = np.stack([sph * cth * cps - cph * sth * sps, cph * sth * cps + sph * cth * sps,
                  cph * cth * sps - sph * sth * cps, cph * cth * cps + sph * sth * sps], axis=-1)
    return q / np.sqrt((q * q).sum(axis=-1, keepdims=True))


def quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def quat_exp(w, dt):
    th = w * dt
    ang = np.sqrt((th * th).sum(axis=-1))
    half = 0.5 * ang
    small = ang < 1e-8
    k = np.where(small, 0.5 * (1.0 - half * half / 6.0), np.sin(half) / np.where(small, 1.0, ang))
    return np.concatenate([th * k[..., None], np.cos(half)[..., None]], axis=-1)


def norm_ang(x):
    """utils/math.py:75-80."""
    two_pi = 2.0 * np.pi
    x = np.where(x > np.pi, x - two_pi * np.ceil((x - np.pi) / two_pi), x)
    return np.where(x < -np.pi, x + two_pi * np.ceil((-np.pi - x) / two_pi), x)


def _mv(M, v):
    return np.einsum("...ij,...j->...i", M, v)


def _mtv(M, v):
    return np.einsum("...ji,...j->...i", M, v)


class BatchOracle:
    """``n_envs`` x ``len(slot_types)`` vehicles, the example loop of ``oracle/sim.py::OracleSwarm`` over arrays."""

    def __init__(self, slot_types, n_envs, gnd=False, drag=False, dw=False, freq=240, aggregate_phy_steps=1, composite=True):
        self.E, self.D = int(n_envs), len(slot_types)
        self.types = list(slot_types)
        self.gnd, self.drag, self.dw = bool(gnd), bool(drag), bool(dw)
        self.TIMESTEP, self.K = 1.0 / freq, int(aggregate_phy_steps)
        D = self.D
        pp = [od.PhysParams(vt, composite) for vt in self.types]
        self.n_u = np.array([p.n_u for p in pp])
        z = np.zeros
        self.m = np.array([p.m for p in pp])
        self.J = np.stack([p.J for p in pp])
        self.J_inv = np.stack([p.J_inv for p in pp])
        self.rc = np.stack([p.r_com for p in pp])
        self.kf, self.km = np.array([p.kf for p in pp]), np.array([p.km for p in pp])
        self.scale, self.const, self.spin = z((D, 6)), z((D, 6)), z((D, 6))
        self.rpos, self.raxis, self.taxis = z((D, 6, 3)), z((D, 6, 3)), z((D, 6, 3))
        self.min_pwm, self.max_pwm = z((D, 6)), z((D, 6))
        for d, p in enumerate(pp):
            n = p.n_u
            self.scale[d, :n], self.const[d, :n], self.spin[d, :n] = p.scale, p.const, p.spin
            self.rpos[d, :n], self.raxis[d, :n], self.taxis[d, :n] = p.rotor_pos, p.rotor_axis, p.torque_axis
            self.min_pwm[d, :n], self.max_pwm[d, :n] = p.min_pwm, p.max_pwm
        self.rotor_on = (np.arange(6)[None, :] < self.n_u[:, None]).astype(float)  # [D, 6]
        self.arm = self.rpos - self.rc[:, None, :]  # rotor sites relative to the centre of mass
        self.gnd_coeff = np.array([p.gnd_coeff for p in pp])
        self.prop_radius = np.array([p.prop_radius for p in pp])
        self.gnd_h_clip = np.array([p.gnd_h_clip for p in pp])
        self.drag_coeff = np.stack([p.drag_coeff for p in pp])
        self.dwc = np.array([p.dw for p in pp])  # [D, 3]
        # controllers (oracle/control.py::_Base)
        self.six = np.array([vt.INDI_OUTPUT_NR == 6 for vt in self.types])
        self.kp = np.array([vt.guidance_indi_pos_gain for vt in self.types], float)
        self.kd = np.array([vt.guidance_indi_speed_gain for vt in self.types], float)
        self.att = np.array([vt.att_gain for vt in self.types], float)
        self.rate = np.array([vt.rate_gain for vt in self.types], float)
        self.alloc4 = z((D, 6, 4))  # pinv(G1 / 0.05) of the quad law (INDIControl.py:459)
        self.alloc6 = z((D, 6, 6))  # first-iteration WLS matrix of the 6-DOF law: u_opt = alloc6 v
        self.G1 = [np.array(vt.G1, float) for vt in self.types]
        for d, vt in enumerate(self.types):
            n = vt.INDI_ACTUATOR_NR
            if self.six[d]:
                B = self.G1[d] / 0.05
                gam, Wv = 100000.0, oc.Hexa6DOFINDI.WV.astype(float)  # wls_alloc defaults (:125), INDIControl_6DOF.py:618
                A = np.vstack([gam * Wv[:, None] * B, np.eye(n)])
                self.alloc6[d, :n, :] = np.linalg.pinv(A)[:, :6] * (gam * Wv)[None, :]
            else:
                self.alloc4[d, :n, :] = np.linalg.pinv(self.G1[d] / 0.05)
        self.wls_slow = 0
        self.wls_fail = 0

    # ------------------------------------------------------------------ state
    def reset(self, pos0, rpy0=None, vel0=None):
        E, D = self.E, self.D
        self.pos = np.array(pos0, float).reshape(E, D, 3).copy()
        rpy0 = np.zeros((E, D, 3)) if rpy0 is None else np.array(rpy0, float).reshape(E, D, 3)
        self.quat = quat_from_euler(rpy0)
        self.vel = np.zeros((E, D, 3)) if vel0 is None else np.array(vel0, float).reshape(E, D, 3).copy()
        self.rates = np.zeros((E, D, 3))
        self.last_clipped_action = np.zeros((E, D, 6))
        self.step_counter = 0
        six = self.six[None, :]
        self.last_vel = np.zeros((E, D, 3))
        self.last_rates = np.zeros((E, D, 3))
        self.last_thrust = np.where(six, 0.3, 0.0) * np.ones((E, D))          # INDIControl_6DOF.py:232 / INDIControl.py:127
        self.cmd = (np.where(six, 0.5, 0.0)[..., None] * self.rotor_on[None]) * np.ones((E, D, 6))  # :234 / :129

    # ------------------------------------------------------------------ physics (BaseAviary.step)
    def physics_step(self, action):
        E, D = self.E, self.D
        dt = self.TIMESTEP
        action = np.asarray(action, float).reshape(E, D, -1)
        clipped = np.clip(action[..., :6], self.min_pwm[None], self.max_pwm[None]) * self.rotor_on[None]  # CtrlAviary.py:258-263
        rpm = (self.scale[None] * clipped + self.const[None]) * self.rotor_on[None]
        T = self.kf[None, :, None] * rpm ** 2
        Q = self.km[None, :, None] * rpm ** 2
        f_rot = T[..., None] * self.raxis[None]                                               # [E, D, 6, 3]
        F0 = f_rot.sum(axis=2)
        tau0 = (np.cross(self.arm[None], f_rot) + (self.spin[None] * Q)[..., None] * self.taxis[None]).sum(axis=2)
        for k in range(self.K):
            R = rotmat(self.quat)
            F, tau = F0.copy(), tau0.copy()
            if self.gnd:  # _groundEffect :1672-1699
                rpy = euler_from_quat(self.quat)
                gate = (np.abs(rpy[..., 0]) < np.pi / 2) & (np.abs(rpy[..., 1]) < np.pi / 2)
                h = self.pos[..., 2:3] + np.einsum("edj,dij->edi", R[..., 2, :], self.rpos)
                h = np.maximum(h, self.gnd_h_clip[None, :, None])
                g = rpm ** 2 * (self.kf * self.gnd_coeff)[None, :, None] * (self.prop_radius[None, :, None] / (4 * h)) ** 2
                g = g * gate[..., None] * self.rotor_on[None]
                fg = g[..., None] * self.raxis[None]
                F += fg.sum(axis=2)
                tau += np.cross(self.arm[None], fg).sum(axis=2)
            if self.drag:  # _drag :1719-1732; the first substep still sees the previously applied action (:532, :545)
                prev = self.last_clipped_action if k == 0 else clipped
                prev_sum = ((self.scale[None] * prev + self.const[None]) * self.rotor_on[None]).sum(axis=-1)
                factors = -1 * self.drag_coeff[None] * prev_sum[..., None] * (2 * np.pi / 60)
                f = _mv(R, factors * self.vel)
                F += f
                tau += np.cross(-self.rc[None], f)
            if self.dw:  # _downwash :1747-1763
                dz = self.pos[:, None, :, 2] - self.pos[:, :, None, 2]                         # [E, i, j] = z_j - z_i
                dxy = np.linalg.norm(self.pos[:, None, :, 0:2] - self.pos[:, :, None, 0:2], axis=-1)
                on = (dz > 0) & (dxy < 10)
                dzs = np.where(on, dz, 1.0)
                alpha = self.dwc[None, :, None, 0] * (self.prop_radius[None, :, None] / (4 * dzs)) ** 2
                beta = self.dwc[None, :, None, 1] * dzs + self.dwc[None, :, None, 2]
                with np.errstate(divide="ignore", over="ignore"):
                    term = np.where(on, alpha * np.exp(-0.5 * (dxy / beta) ** 2), 0.0)
                fz = -term.sum(axis=2)
                f = np.zeros((E, D, 3))
                f[..., 2] = fz
                F += f
                tau += np.cross(-self.rc[None], f)
            # substep_quat (R8)
            rc, w = self.rc[None], self.rates
            c = self.pos + _mv(R, np.broadcast_to(rc, w.shape))
            vc = self.vel + _mv(R, np.cross(w, rc))
            acc = _mv(R, F) / self.m[None, :, None] - np.array([0.0, 0.0, G])
            wdot = _mv(self.J_inv[None], tau - np.cross(w, _mv(self.J[None], w)))
            vc = vc + dt * acc
            w = w + dt * wdot
            c = c + dt * vc
            q = quat_mul(self.quat, quat_exp(w, dt))
            q = q / np.sqrt((q * q).sum(axis=-1, keepdims=True))
            R2 = rotmat(q)
            self.pos = c - _mv(R2, np.broadcast_to(rc, w.shape))
            self.vel = vc - _mv(R2, np.cross(w, rc))
            self.quat, self.rates = q, w
        self.last_clipped_action = clipped
        self.step_counter += self.K

    # ------------------------------------------------------------------ control (computeControlFromState per vehicle)
    def control_step(self, tpos, tvel=None, tacc=None, tyaw=None):
        E, D = self.E, self.D
        dt = self.K * self.TIMESTEP
        tpos = np.asarray(tpos, float).reshape(E, D, 3)
        tvel = np.zeros((E, D, 3)) if tvel is None else np.asarray(tvel, float).reshape(E, D, 3)
        tacc = np.zeros((E, D, 3)) if tacc is None else np.asarray(tacc, float).reshape(E, D, 3)
        tyaw = np.zeros((E, D)) if tyaw is None else np.asarray(tyaw, float).reshape(E, D)
        six = np.broadcast_to(self.six[None, :], (E, D))
        q, v = self.quat, self.vel
        R = rotmat(q)
        # position loop (INDIControl.py:278-296 / INDIControl_6DOF.py:390-413)
        pos_e = tpos - self.pos
        accel_sp = (pos_e * self.kp[None, :, None] + tvel - v) * self.kd[None, :, None]
        cur_accel = (v - self.last_vel) / dt
        self.last_vel = v.copy()
        accel_e = np.clip(accel_sp + np.where(six[..., None], 0.0, tacc) - cur_accel, -6.0, 6.0)
        rpy = euler_from_quat(q)
        phi, theta, psi = rpy[..., 0], rpy[..., 1], rpy[..., 2]
        Tg = 9.81
        sph, sth, sps, cph, cth, cps = np.sin(phi), np.sin(theta), np.sin(psi), np.cos(phi), np.cos(theta), np.cos(psi)
        Gm = np.empty((E, D, 3, 3))
        Gm[..., 0, 0], Gm[..., 0, 1], Gm[..., 0, 2] = (cph * sps - sph * cps * sth) * Tg, (cph * cps * cth) * Tg, sph * sps + cph * cps * sth
        Gm[..., 1, 0], Gm[..., 1, 1], Gm[..., 1, 2] = (-sph * sps * sth - cps * cph) * Tg, (cph * sps * cth) * Tg, cph * sps * sth - cps * sph
        Gm[..., 2, 0], Gm[..., 2, 1], Gm[..., 2, 2] = -cth * sph * Tg, -sth * cph * Tg, cph * cth
        inc = _mv(np.linalg.pinv(Gm), accel_e)
        thrust = self.last_thrust + inc[..., 2]
        # attitude loop
        yaw_inc = norm_ang(tyaw - psi)
        te = rpy + np.stack([inc[..., 0], inc[..., 1], yaw_inc], axis=-1)
        te = np.where(six[..., None], 0.0, te)                                  # INDIControl_6DOF.py:495
        tq = quat_from_euler(te)
        x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
        ew = w * tq[..., 3] + x * tq[..., 0] + y * tq[..., 1] + z * tq[..., 2]   # utils/math.py:23-31
        ex = w * tq[..., 0] - x * tq[..., 3] - y * tq[..., 2] + z * tq[..., 1]
        ey = w * tq[..., 1] + x * tq[..., 2] - y * tq[..., 3] - z * tq[..., 0]
        ez = w * tq[..., 2] - x * tq[..., 1] + y * tq[..., 0] - z * tq[..., 3]
        flip = (~six) & (ew < 0)                                                # quat_wrap_shortest (quad law only, quirk Q1)
        att_err = np.stack([ex, ey, ez], axis=-1) * np.where(flip, -1.0, 1.0)[..., None]
        r0 = cps * att_err[..., 0] + sps * att_err[..., 1]                       # inv(R_psi), 6-DOF law (:551-557)
        r1 = -sps * att_err[..., 0] + cps * att_err[..., 1]
        att_err = np.where(six[..., None], np.stack([r0, r1, att_err[..., 2]], axis=-1), att_err)
        rate_sp = self.att[None] * att_err
        # rate loop (INDIControl.py:428-453); the state vector carries world rates, the law rotates them back
        wb = _mtv(R, _mv(R, self.rates))
        ang_acc = (wb - self.last_rates) / dt
        self.last_rates = wb
        nu3 = (rate_sp - wb) * self.rate[None] - ang_acc
        dthr = thrust - self.last_thrust
        self.last_thrust = thrust
        du4 = np.einsum("dij,edj->edi", self.alloc4, np.concatenate([nu3, dthr[..., None]], axis=-1))
        v6 = np.concatenate([nu3, _mtv(R, accel_e)], axis=-1)
        du6 = np.einsum("dij,edj->edi", self.alloc6, v6)
        umin, umax = self.min_pwm[None] - self.cmd, self.max_pwm[None] - self.cmd
        bad = six & (((du6 >= umax + 1.0) | (du6 <= umin - 1.0)) & (self.rotor_on[None] > 0)).any(axis=-1)
        for e, d in zip(*np.nonzero(bad)):  # wls_alloc.py:262-347: the active-set iterations, per vehicle
            n = self.n_u[d]
            u, it = oc.wls_alloc(v6[e, d], umin[e, d, :n], umax[e, d, :n], self.G1[d] / 0.05, None, None,
                                 oc.Hexa6DOFINDI.WV, np.ones(n), None)
            self.wls_slow += 1
            if u is None:
                self.wls_fail += 1
                u = np.zeros(n)
            du6[e, d, :n] = u
        du = np.where(six[..., None], du6, du4)
        self.cmd = np.clip(self.cmd + du, self.min_pwm[None], self.max_pwm[None]) * self.rotor_on[None]
        self.pos_e = pos_e
        return self.cmd.copy()
