"""FP64 restatement of the reference INDI controllers and WLS allocator.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  One vehicle at a time, deliberately literal:
numpy float64, ``np.linalg.pinv`` / ``np.linalg.lstsq`` exactly where the reference calls them.

Restated (paths relative to the reference checkout):
* ``dronesim/utils/math.py:23-31,46-51,75-80``     -> quat_inv_comp, quat_wrap_shortest, norm_ang
* ``dronesim/control/INDIControl.py:109-146,154-227,232-351,355-411,413-490`` -> QuadINDI
* ``dronesim/control/INDIControl_6DOF.py:214-251,259-336,341-496,500-634``    -> Hexa6DOFINDI
* ``dronesim/control/wls_alloc.py:125-350``        -> wls_alloc
* ``dronesim/control/BaseControl.py:61-103``       -> computeControlFromState slicing

Pinned in ``tests/test_oracle_control.py`` against fixtures produced by executing the
reference's own classes (``tests/golden/make_golden.py``) and the MATLAB ``lsqlin`` known
answer in ``wls_alloc.py:381-408``.

Documented deviation: the reference's ``Gains`` object is a class attribute shared
by every controller instance (utils/utils.py:21-24, INDIControl_6DOF.py:33-36), so the last
constructed controller's URDF gains win.  This oracle uses per-type gains.
"""
import numpy as np

from . import pyb_math as p

FLT_EPSILON = 1e-7  # wls_alloc.py:87
INFINITY = 1e32  # wls_alloc.py:88


# ---------------------------------------------------------------- utils/math.py
def quat_inv_comp(q1, q2):
    """utils/math.py:23-31 (xyzw): conj(q1) (x) q2."""
    i, x, y, z = 3, 0, 1, 2
    qerr = np.zeros(4)
    qerr[i] = q1[i] * q2[i] + q1[x] * q2[x] + q1[y] * q2[y] + q1[z] * q2[z]
    qerr[x] = q1[i] * q2[x] - q1[x] * q2[i] - q1[y] * q2[z] + q1[z] * q2[y]
    qerr[y] = q1[i] * q2[y] + q1[x] * q2[z] - q1[y] * q2[i] - q1[z] * q2[x]
    qerr[z] = q1[i] * q2[z] - q1[x] * q2[y] + q1[y] * q2[x] - q1[z] * q2[i]
    return qerr


def quat_wrap_shortest(q):
    """utils/math.py:46-51 - in place."""
    if q[3] < 0:
        for k in range(4):
            q[k] = -q[k]
    return q


def norm_ang(x):
    """utils/math.py:75-80."""
    while x > np.pi:
        x -= 2 * np.pi
    while x < -np.pi:
        x += 2 * np.pi
    return x


# ---------------------------------------------------------------- wls_alloc.py
def wls_alloc(v, umin, umax, B, u_guess, W_init, Wv, Wu, up, gamma_sq=100000, imax=100, return_W=False):
    """wls_alloc.py:125-350, same control flow and the same integer working-set bookkeeping.

    Returns ``(u, iterations)`` or ``(None, iterations)`` on non-convergence (wls_alloc.py:350);
    with ``return_W`` also the final working set ``W`` in {-1, 0, +1} (a local of the reference, :171).
    """
    n_u = len(umin)
    n_v = len(v)
    n_c = n_u + n_v
    A = np.zeros((n_c, n_u))
    A_free = np.zeros((n_c, n_u))
    b = np.zeros(n_c)
    d = np.zeros(n_c)
    free_index = np.zeros(n_u, dtype=int)
    n_free = 0
    free_chk = -1
    it = 0
    p_free = np.zeros(n_u)
    u = np.zeros(n_u)
    # :166-177
    if u_guess is None:
        for i in range(n_u):
            u[i] = (umax[i] + umin[i]) * 0.5
    else:
        u = np.array(u_guess, dtype=float).copy()
    W = np.array(W_init, dtype=float).copy() if W_init is not None else np.zeros(n_u)
    free_index_lookup = np.ones(n_u, dtype=int) * -1
    for i in range(n_u):  # :182-186
        if W[i] == 0:
            free_index_lookup[i] = n_free
            free_index[n_free] = i
            n_free += 1
    for i in range(n_v):  # :190-203
        b[i] = gamma_sq * Wv[i] * v[i] if Wv is not None else gamma_sq * v[i]
        d[i] = b[i]
        for j in range(n_u):
            A[i][j] = gamma_sq * Wv[i] * B[i][j] if Wv is not None else gamma_sq * B[i][j]
            d[i] -= A[i][j] * u[j]
    for i in range(n_v, n_c):  # :205-219
        A[i, :] = 0
        A[i][i - n_v] = Wu[i - n_v] if Wu is not None else 1.0
        if up is not None:
            b[i] = Wu[i - n_v] * up[i - n_v] if Wu is not None else up[i - n_v]
        else:
            b[i] = 0
        d[i] = b[i] - A[i][i - n_v] * u[i - n_v]

    while it < imax:  # :222
        it += 1
        pvec = np.zeros(n_u)
        u_opt = u.copy()
        if free_chk != n_free:  # :233-238
            for i in range(n_c):
                for j in range(n_free):
                    A_free[i][j] = A[i][free_index[j]]
            free_chk = n_free
        if n_free:  # :243-252
            p_free = np.linalg.lstsq(A_free[:n_c, :n_free], d, rcond=None)[0]
        for i in range(n_free):  # :257-259
            pvec[free_index[i]] = p_free[i]
            u_opt[free_index[i]] += p_free[i]
        n_infeasible = 0  # :262-266  (note the +-1.0 slack)
        for i in range(n_u):
            if u_opt[i] >= (umax[i] + 1.0) or u_opt[i] <= (umin[i] - 1.0):
                n_infeasible += 1
        if n_infeasible == 0:  # :269-298
            u = u_opt.copy()
            Lambda = np.zeros(n_u)
            for i in range(n_c):
                for k in range(n_free):
                    d[i] -= A_free[i][k] * p_free[k]
                for k in range(n_u):
                    Lambda[k] += A[i][k] * d[i]
            break_flag = True
            for i in range(n_u):
                Lambda[i] *= W[i]
                if Lambda[i] < -FLT_EPSILON:
                    break_flag = False
                    W[i] = 0
                    if free_index_lookup[i] < 0:
                        free_index_lookup[i] = n_free
                        free_index[n_free] = i
                        n_free += 1
            if break_flag:
                return (u, it, W.astype(int)) if return_W else (u, it)
            # NOTE: falls through with the *previous* alpha / id_alpha, exactly as the
            # reference does (its ``else`` at :299 only resets them on the infeasible branch).
        else:  # :299-302
            alpha = INFINITY
            alpha_tmp = 0.0
            id_alpha = 0
        for i in range(n_free):  # :305-317
            idx = free_index[i]
            if np.abs(pvec[idx]) > FLT_EPSILON:
                alpha_tmp = (umin[idx] - u[idx]) / pvec[idx] if pvec[idx] < 0 else (umax[idx] - u[idx]) / pvec[idx]
            else:
                alpha_tmp = INFINITY
            if alpha_tmp < alpha:
                alpha = alpha_tmp
                id_alpha = idx
        for i in range(n_u):  # :320-321
            u[i] += alpha * pvec[i]
        for i in range(n_c):  # :324-332
            k_len = min(n_free, len(p_free))
            for k in range(k_len):
                d[i] -= A_free[i][k] * alpha * p_free[k]
        W[id_alpha] = 1.0 if pvec[id_alpha] > 0 else -1.0  # :335-338
        n_free -= 1  # :342-347
        free_index[free_index_lookup[id_alpha]] = free_index[n_free]
        free_index_lookup[free_index[free_index_lookup[id_alpha]]] = free_index_lookup[id_alpha]
        free_index_lookup[id_alpha] = -1
    return (None, it, W.astype(int)) if return_W else (None, it)


# ---------------------------------------------------------------- shared position loop
def _G_matrix(phi, theta, psi, T=9.81):
    """INDIControl.py:304-333 == INDIControl_6DOF.py:424-453 (T = 9.81, quirk Q2)."""
    sph, sth, sps = np.sin(phi), np.sin(theta), np.sin(psi)
    cph, cth, cps = np.cos(phi), np.cos(theta), np.cos(psi)
    return np.array(
        [
            [(cph * sps - sph * cps * sth) * T, (cph * cps * cth) * T, sph * sps + cph * cps * sth],
            [(-sph * sps * sth - cps * cph) * T, (cph * sps * cth) * T, cph * sps * sth - cps * sph],
            [-cth * sph * T, -sth * cph * T, cph * cth],
        ]
    )


class _Base:
    def __init__(self, vt):
        """``vt``: a ``dronesim_b200.vehicles.VehicleType`` (the URDF-derived table)."""
        self.vt = vt
        self.n_u = vt.INDI_ACTUATOR_NR
        self.G1 = np.array(vt.G1, dtype=float)
        self.kp = vt.guidance_indi_pos_gain
        self.kd = vt.guidance_indi_speed_gain
        self.att = np.array(vt.att_gain, dtype=float)
        self.rate = np.array(vt.rate_gain, dtype=float)
        self.MIN_PWM = np.array(vt.MIN_PWM, dtype=float)
        self.MAX_PWM = np.array(vt.MAX_PWM, dtype=float)
        # extension beyond the reference (its filter is a commented placeholder, INDIControl.py:432-439):
        # first-order low-pass on the angular-acceleration estimate, coefficient b per control step; None = off
        self.acc_b = None
        self.ang_acc_filt = np.zeros(3)
        self.reset()

    def _filter_ang_acc(self, angular_accel):
        if self.acc_b is None:
            return angular_accel
        self.ang_acc_filt = self.ang_acc_filt + self.acc_b * (angular_accel - self.ang_acc_filt)
        return self.ang_acc_filt.copy()

    def computeControlFromState(self, control_timestep, state, target_pos, target_vel=np.zeros(3),
                                target_acc=np.zeros(3), target_rpy=np.zeros(3), target_rpy_rates=np.zeros(3)):
        """BaseControl.py:61-103."""
        return self.computeControl(
            control_timestep=control_timestep, cur_pos=state[0:3], cur_quat=state[3:7], cur_vel=state[10:13],
            cur_ang_vel=state[13:16], target_pos=target_pos, target_vel=target_vel, target_acc=target_acc,
            target_rpy=target_rpy, target_rpy_rates=target_rpy_rates)

    def memory(self):
        """Controller memory as one vector: last_vel3, last_rates3, last_thrust, cmd[n_u]."""
        return np.concatenate([self.last_vel, self.last_rates, [self.last_thrust], self.cmd])


class QuadINDI(_Base):
    """INDIControl.py (quad / 4-virtual-control law; also flies hexa_6DOF_simple)."""

    def reset(self):  # INDIControl.py:109-146
        self.control_counter = 0
        self.last_rates = np.zeros(3)
        self.last_thrust = 0.0
        self.cmd = np.ones(self.n_u) * 0.0
        self.last_vel = np.zeros(3)
        self.ang_acc_filt = np.zeros(3)

    def computeControl(self, control_timestep, cur_pos, cur_quat, cur_vel, cur_ang_vel, target_pos,
                       target_vel=np.zeros(3), target_acc=np.zeros(3), target_rpy=np.zeros(3),
                       target_rpy_rates=np.zeros(3)):
        self.control_counter += 1  # :203
        dt = control_timestep
        # ---- _INDIPositionControl :278-351
        pos_e = np.asarray(target_pos, float) - cur_pos
        speed_sp = pos_e * self.kp
        vel_e = speed_sp + target_vel - cur_vel
        accel_sp = vel_e * self.kd
        cur_accel = (cur_vel - self.last_vel) / dt
        self.last_vel = np.array(cur_vel, dtype=float)
        accel_e = np.clip(accel_sp + target_acc - cur_accel, -6.0, 6.0)
        cur_rpy = np.array(p.getEulerFromQuaternion(cur_quat))
        phi, theta, psi = cur_rpy
        G_inv = np.linalg.pinv(_G_matrix(phi, theta, psi))
        control_increment = G_inv.dot(accel_e)
        yaw_increment = norm_ang(target_rpy[2] - psi)
        target_euler = cur_rpy + np.array([control_increment[0], control_increment[1], yaw_increment])
        thrust = self.last_thrust + control_increment[2]
        # ---- _INDIAttitudeControl :388-402
        target_quat = np.array(p.getQuaternionFromEuler(target_euler))
        quat_err = quat_inv_comp(cur_quat, target_quat)
        quat_wrap_shortest(quat_err)  # in place (quirk Q1)
        att_err = np.array(quat_err[:3])
        rate_sp = self.att * att_err
        # ---- _INDIRateControl :428-490
        self.cmd = self.rate_control(dt, thrust, cur_quat, cur_ang_vel, rate_sp)
        cur_rpy2 = p.getEulerFromQuaternion(cur_quat)  # :225
        return self.cmd, pos_e, target_euler[2] - cur_rpy2[2]

    def rate_control(self, dt, thrust, cur_quat, cur_ang_vel, rate_sp):
        """INDIControl._INDIRateControl (INDIControl.py:413-490); also the RPYTAviary entry."""
        R = p.rotmat(cur_quat)
        w = R.T.dot(cur_ang_vel)
        angular_accel = self._filter_ang_acc((w - self.last_rates) / (1.0 * dt))
        self.last_rates = w
        indi_v = np.zeros(4)
        indi_v[0:3] = (np.asarray(rate_sp) - w) * self.rate - angular_accel
        indi_v[3] = thrust - self.last_thrust
        self.last_thrust = thrust
        indi_du = np.dot(np.linalg.pinv(self.G1 / 0.05), indi_v)
        cmd = self.cmd + indi_du
        return np.clip(cmd, self.MIN_PWM, self.MAX_PWM)


class Hexa6DOFINDI(_Base):
    """INDIControl_6DOF.py (6 virtual controls, WLS allocation)."""

    WV = np.array([1000, 1000, 0.1, 10, 10, 100])  # INDIControl_6DOF.py:618

    def reset(self):  # INDIControl_6DOF.py:214-251
        self.control_counter = 0
        self.last_rates = np.zeros(3)
        self.last_thrust = 0.3
        self.cmd = np.ones(self.n_u) * 0.5
        self.last_vel = np.zeros(3)
        self.ang_acc_filt = np.zeros(3)
        self.wls_fail = 0
        self.last_wls_iter = 0

    def computeControl(self, control_timestep, cur_pos, cur_quat, cur_vel, cur_ang_vel, target_pos,
                       target_vel=np.zeros(3), target_acc=np.zeros(3), target_rpy=np.zeros(3),
                       target_rpy_rates=np.zeros(3)):
        self.control_counter += 1
        dt = control_timestep
        # ---- _INDIPositionControl :390-496
        pos_e = np.asarray(target_pos, float) - cur_pos
        speed_sp = pos_e * self.kp
        vel_e = speed_sp + target_vel - cur_vel
        accel_sp = vel_e * self.kd
        cur_accel = (cur_vel - self.last_vel) / dt
        self.last_vel = np.array(cur_vel, dtype=float)
        accel_e = np.clip(accel_sp - cur_accel, -6.0, 6.0)  # target_acc ignored (:410)
        cur_rpy = np.array(p.getEulerFromQuaternion(cur_quat))
        phi, theta, psi = cur_rpy
        G_inv = np.linalg.pinv(_G_matrix(phi, theta, psi))
        control_increment = G_inv.dot(accel_e)
        thrust = self.last_thrust + control_increment[2]  # :491 (only feeds last_thrust, quirk Q7)
        target_euler = np.zeros(3)  # :495
        # ---- _INDIAttitudeControl :538-634
        target_quat = np.array(p.getQuaternionFromEuler(target_euler))
        quat_err = quat_inv_comp(cur_quat, target_quat)
        att_err = np.array(quat_err[:3])  # no shortest-wrap (:543)
        R_psi = np.array([[np.cos(psi), -np.sin(psi)], [np.sin(psi), np.cos(psi)]])
        R_psi = np.linalg.inv(R_psi)
        att_err[:2] = R_psi.dot(att_err[:2])
        rate_sp = self.att * att_err
        R = p.rotmat(cur_quat)
        w = R.T.dot(cur_ang_vel)
        angular_accel = self._filter_ang_acc((w - self.last_rates) / (1.0 * dt))
        self.last_rates = w
        indi_v = np.zeros(6)
        indi_v[0:3] = (rate_sp - w) * self.rate - angular_accel
        indi_v[3:6] = R.T.dot(accel_e)
        self.last_thrust = thrust
        umin = self.MIN_PWM - self.cmd
        umax = self.MAX_PWM - self.cmd
        indi_du, nit = wls_alloc(indi_v, umin, umax, self.G1 / 0.05, None, None, self.WV,
                                 np.ones(self.n_u), None)
        self.last_wls_iter = nit
        if indi_du is None:
            # the reference crashes here (``self.cmd += None``, INDIControl_6DOF.py:630).
            # Defined behaviour of the new core: hold the command, count the failure.
            self.wls_fail += 1
            indi_du = np.zeros(self.n_u)
        self.cmd = np.clip(self.cmd + indi_du, self.MIN_PWM, self.MAX_PWM)
        cur_rpy2 = p.getEulerFromQuaternion(cur_quat)
        return self.cmd, pos_e, target_euler[2] - cur_rpy2[2]


def make_controller(vt):
    """The controller class the reference examples pair with this vehicle type."""
    return Hexa6DOFINDI(vt) if vt.INDI_OUTPUT_NR == 6 else QuadINDI(vt)
