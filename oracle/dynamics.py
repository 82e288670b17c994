"""FP64 restatement of the reference's explicit dynamics and aerodynamic add-ons.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The functions restated here are dead code INSIDE
the reference fork (they read ``self.KF, self.M, self.J ...`` whose assignments are commented out,
BaseAviary.py:200-216,226-235, and the live path needs PyBullet), so the reference cannot step
them.  Their bodies do execute when called unbound on a stand-in ``self`` with a recording stand-in
for the module's ``p``: ``tests/golden/dyn_{robobee,tello}.npz`` hold the outputs of the reference's own
``_dynamics / _drag / _downwash / _groundEffect`` run that way (tests/golden/make_golden.py::
dynamics_fixture) and ``tests/test_oracle_dynamics.py`` pins this module to them to 1e-12.
**Pinned for the 4-rotor airframes; the generalisations that let the same formulas fly a
heterogeneous list and the tilted-rotor hexa (repairs R1-R7 below) and the ``"quat"`` integrator (R8,
beyond the reference) are restated, not pinned.**  Each repair is listed with the reference line it replaces.

Restated (dronesim/envs/BaseAviary.py):
* motor map            :1487-1490, :1398-1401   rpm = PWM2RPM_SCALE * cmd + PWM2RPM_CONST
* rotor thrust/torque  :1515-1543 (quad), :1402-1457 (hexa)   F = KF rpm^2, tau = KM rpm^2 (noise off)
* ``_dynamics``        :1767-1828   (integrator ``"rpy"``)
* ``_groundEffect``    :1648-1699, clip constant :235
* ``_drag``            :1705-1732  (uses the previous substep's action, :532,:545)
* ``_downwash``        :1736-1763
* state vector         :718-732, :764-790

Repairs (R#) - all needed because the reference code cannot run as written:
R1  ``self.X`` -> per-drone ``drones[i].X``                                   (:1788-1812 etc.)
R2  ``_dynamics`` receives PWM commands, not RPM: the motor map is applied first (:526 passes
    the clipped PWM action; :1788 squares it as if it were RPM).
R3  x/y torques from the URDF rotor geometry  sum_i r_i x (F_i a_i)  instead of the
    CF2X/CF2P switch (:1794-1803), which compares a list with an enum and never binds x_torque.
    The same body wrench is used by both integrators, so tilted hexa rotors produce lateral force.
R4  z torque uses every rotor's spin sign (-,+,-,+,...) instead of the first four only (:1793).
R5  world angular velocity reported as R(q) . rates instead of the [-1,-1,-1] placeholder (:1821-1826).
R6  ``_groundEffect`` loops over the vehicle's n_u rotor sites (links :1665 hard-codes 0..4).
R7  ``_drag`` / ``_downwash`` forces act at the base-frame origin along LINK_FRAME axes (the
    reference names link 4 = ``center_of_mass_link`` of the quads, :1727,:1758); downwash uses the
    RECEIVING drone's coefficients.
R8  integrator ``"quat"`` (beyond the reference, asked by north_star): Newton-Euler about the
    composite centre of mass, semi-implicit Euler, exponential-map quaternion update.
R10 "advanced" quad types (``"advanced" in TYPE``, :1493): the oblique-flow propeller fit of ``_get_prop_FMs`` replaces
    KF rpm^2 / KM rpm^2; pinned by ``tests/golden/rotor_tello_advanced.npz`` (the reference's own branch executed).
R9  first-order motor model (beyond the reference, asked by north_star; off by default): per substep
    rpm += (1 - exp(-dt / tau)) (rpm_cmd - rpm); thrust / torque / ground effect use the actual rpm, drag the
    actual rpm sum before the substep's update.  tau = 0 is the reference's static map (:1487-1490).
"""
import math

import numpy as np

from . import pyb_math as p

G = 9.8  # BaseAviary.py:182


class PhysParams:
    """Per-type constants of the dynamics (what the CUDA core keeps in shared memory)."""

    def __init__(self, vt, composite: bool):
        """``composite=False``: mass/inertia as parsed by the reference (first link only,
        BaseAviary.py:2055-2069) - the literal ``Physics.DYN`` numbers.
        ``composite=True``: whole-tree mass / inertia / centre of mass - what PyBullet simulates."""
        self.vt = vt
        self.n_u = vt.INDI_ACTUATOR_NR
        if composite:
            self.m, self.J, self.r_com = float(vt.M_TOTAL), np.array(vt.J_TOTAL, float), np.array(vt.COM, float)
        else:
            self.m, self.J, self.r_com = float(vt.M), np.array(vt.J, float), np.zeros(3)
        self.J_inv = np.linalg.inv(self.J)
        self.kf, self.km = float(vt.KF), float(vt.KM)
        self.scale = np.array(vt.PWM2RPM_SCALE, float)
        self.const = np.array(vt.PWM2RPM_CONST, float)
        self.rotor_pos = np.array(vt.rotor_pos, float)
        self.rotor_axis = np.array(vt.rotor_axis, float)
        self.torque_axis = np.array(vt.torque_axis, float)
        self.spin = np.array(vt.rotor_spin, float)
        self.gnd_coeff = float(vt.GND_EFF_COEFF)
        self.prop_radius = float(vt.PROP_RADIUS)
        self.gnd_h_clip = float(vt.GND_EFF_H_CLIP)
        self.drag_coeff = np.array(vt.DRAG_COEFF, float)
        self.dw = (float(vt.DW_COEFF_1), float(vt.DW_COEFF_2), float(vt.DW_COEFF_3))
        self.min_pwm = np.array(vt.MIN_PWM, float)
        self.max_pwm = np.array(vt.MAX_PWM, float)


def rpm_of_cmd(pp, cmd):
    return pp.scale * np.asarray(cmd, float) + pp.const


def quat_mul(a, b):
    """Hamilton product, xyzw."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ])


def quat_exp(w, dt):
    """Rotation by body rate ``w`` over ``dt`` as an xyzw quaternion (exact exponential map)."""
    th = np.asarray(w, float) * dt
    ang = math.sqrt(float(th.dot(th)))
    half = 0.5 * ang
    k = 0.5 * (1.0 - half * half / 6.0) if ang < 1e-8 else math.sin(half) / ang
    return np.array([th[0] * k, th[1] * k, th[2] * k, math.cos(half)])


def advanced_rotor_FMs(prop, quat, vel, rpm, rho=1.225):
    """``BaseAviary._get_prop_FMs`` (BaseAviary.py:1570-1644) with ``utils.calculate_propeller_forces_moments``
    method 2 (utils/utils.py:149-202, 343-416): per-rotor force and moment vectors in the body frame.
    ``prop``: {"coeff": 14 numbers of Data_section5_ObliqueFlow, "radius": m}."""
    (CsFT, k1, k2, k3, k4, k5, CsMQ, k6, k7, k8, k9, k10, k11, k12) = prop["coeff"]
    Rp = prop["radius"]
    R = p.rotmat(quat)
    vel = np.asarray(vel, float)
    V_i = vel if np.linalg.norm(vel) > 0.1 else np.array([0.1, 0.0, 0.0])  # :1585-1589
    V_b = R.dot(V_i)  # (sic) the reference rotates with R, not its transpose (:1590)
    V_b_normed = V_b / np.linalg.norm(V_b)
    beta = np.arccos(V_b_normed.dot(np.array([0.0, 0.0, 1.0])))  # :1600-1602
    psi = np.arctan(V_b[1] / V_b[0]) if V_b[0] > 0.1 else 0.0  # :1603-1605
    V = np.linalg.norm(vel)  # the un-substituted speed (:1623)
    F_b, M_b = [], []
    R_z = np.array([[np.cos(psi), -np.sin(psi), 0.0], [np.sin(psi), np.cos(psi), 0.0], [0.0, 0.0, 1.0]])
    for _rpm in rpm:
        omega = _rpm / 60.0 * 2 * np.pi
        omega = omega if omega > 10.0 else 10.0  # utils.py:176
        mu = V * math.sin(beta) / (omega * Rp)  # utils.py:383-384
        lambda_c = V * math.cos(beta) / (omega * Rp)
        cft = CsFT + k1 * lambda_c + k2 * mu**2 + k3 * lambda_c**2  # eq. 95
        cfh = k4 * mu + k5 * lambda_c * mu  # eq. 99
        cmq = CsMQ + k6 * lambda_c + k7 * mu**2 + k8 * lambda_c**2  # eq. 100
        cmr = k9 * mu + k10 * lambda_c * mu  # eq. 101
        cmp_ = k11 * mu + k12 * lambda_c * mu  # eq. 102
        avg = 0.5 * rho * (omega * Rp) ** 2 * math.pi * Rp**2  # utils.py:192-193
        FM = np.array([cfh * avg, 0.0, cft * avg, cmp_ * avg * Rp, cmq * avg * Rp, cmr * avg * Rp])
        F_b.append(R_z.dot(FM[:3]))
        M_b.append(R_z.dot(FM[3:]))
    return F_b, M_b


def body_wrench(pp, cmd, cmd_prev_rpm_sum, pos, quat, rpy, vel, others_pos, gnd, drag, dw, rpm=None, noise=None):
    """Body-frame force and torque about the (composite) centre of mass for one substep.
    ``rpm``: actual rotor speeds when the first-order motor model (extension, R9) is on; default = the static map.
    ``noise``: (f_noise[n_u], m_noise[n_u]) of this substep (:1429-1432 / :1518-1525) or None (noise off)."""
    R = p.rotmat(quat)
    rpm = rpm_of_cmd(pp, cmd) if rpm is None else np.asarray(rpm, float)
    T = pp.kf * rpm**2  # :1515 / :1402
    Q = pp.km * rpm**2  # :1516 / :1403
    F = np.zeros(3)
    tau = np.zeros(3)
    if noise is not None:
        f_noise, m_noise = noise
        T = T + f_noise  # forces += f_noise (:1433 / :1524)
        Q = Q + m_noise  # torques += m_noise (:1434 / :1525)
        if "morphing_hexa" not in pp.vt.TYPE:
            # quad model: every rotor link also gets (f_noise[0], f_noise[1]) laterally (:1528-1536) and the base
            # gets (m_noise[0], m_noise[1]) (:1537-1543); the quads' link frames are the body frame
            lat = np.array([f_noise[0], f_noise[1], 0.0])
            for i in range(pp.n_u):
                F += lat
                tau += np.cross(pp.rotor_pos[i] - pp.r_com, lat)
            tau += np.array([m_noise[0], m_noise[1], 0.0])
    if "advanced" in pp.vt.TYPE:
        # _quad_copter_physics, "advanced" branch (:1493-1512): force F_prop[i] on rotor link i, torque
        # [0, 0, M_prop[i][2] * direction[i]] on the same link, direction = [-1, 1, -1, 1]; no noise on this branch
        from dronesim_b200.vehicles import load_propeller
        F_b, M_b = advanced_rotor_FMs(load_propeller(), quat, vel, rpm)
        direction = [-1.0, 1.0, -1.0, 1.0]
        F = np.zeros(3)
        tau = np.zeros(3)
        for i in range(pp.n_u):
            F += F_b[i]
            tau += np.cross(pp.rotor_pos[i] - pp.r_com, F_b[i]) + np.array([0.0, 0.0, M_b[i][2] * direction[i]])
    else:
        for i in range(pp.n_u):
            f = T[i] * pp.rotor_axis[i]
            F += f
            tau += np.cross(pp.rotor_pos[i] - pp.r_com, f) + pp.spin[i] * Q[i] * pp.torque_axis[i]
    if gnd:  # _groundEffect :1672-1699
        if abs(rpy[0]) < np.pi / 2 and abs(rpy[1]) < np.pi / 2:
            for i in range(pp.n_u):
                h = pos[2] + R[2, :].dot(pp.rotor_pos[i])
                h = max(h, pp.gnd_h_clip)
                g = rpm[i] ** 2 * pp.kf * pp.gnd_coeff * (pp.prop_radius / (4 * h)) ** 2
                f = g * pp.rotor_axis[i]
                F += f
                tau += np.cross(pp.rotor_pos[i] - pp.r_com, f)
    if drag:  # _drag :1719-1732
        drag_factors = -1 * pp.drag_coeff * cmd_prev_rpm_sum * (2 * np.pi / 60)
        f = R.dot(drag_factors * np.asarray(vel, float))
        F += f
        tau += np.cross(-pp.r_com, f)
    if dw:  # _downwash :1747-1763
        for pj in others_pos:
            delta_z = pj[2] - pos[2]
            delta_xy = np.linalg.norm(np.array(pj[0:2]) - np.array(pos[0:2]))
            if delta_z > 0 and delta_xy < 10:
                alpha = pp.dw[0] * (pp.prop_radius / (4 * delta_z)) ** 2
                beta = pp.dw[1] * delta_z + pp.dw[2]
                f = np.array([0, 0, -alpha * np.exp(-0.5 * (delta_xy / beta) ** 2)])
                F += f
                tau += np.cross(-pp.r_com, f)
    return F, tau, R


def substep_rpy(pp, dt, pos, quat, rpy, vel, rates, F, tau, R):
    """``_dynamics`` :1788-1828 with the wrench from ``body_wrench`` (R2-R5)."""
    thrust_world_frame = R.dot(F)
    force_world_frame = thrust_world_frame - np.array([0, 0, G * pp.m])
    torques = tau - np.cross(rates, pp.J.dot(rates))
    rates_deriv = pp.J_inv.dot(torques)
    acc = force_world_frame / pp.m
    vel = vel + dt * acc
    rates = rates + dt * rates_deriv
    pos = pos + dt * vel
    rpy = rpy + dt * rates
    quat = np.array(p.getQuaternionFromEuler(rpy))  # :1817
    rpy = np.array(p.getEulerFromQuaternion(quat))  # :729 (state refresh)
    return pos, quat, rpy, vel, rates


def substep_quat(pp, dt, pos, quat, vel, w, F, tau, R, floor_z=None):
    """R8: rigid body about the composite centre of mass, semi-implicit Euler.
    ``floor_z``: the ground plane of the PyBullet world (plane.urdf, BaseAviary.py:679-680) as an inelastic frictionless
    stop for the centre of mass (a stand-in for Bullet's contact; off by default)."""
    rc_w = R.dot(pp.r_com)
    c = pos + rc_w
    vc = vel + R.dot(np.cross(w, pp.r_com))
    acc = R.dot(F) / pp.m - np.array([0.0, 0.0, G])
    wdot = pp.J_inv.dot(tau - np.cross(w, pp.J.dot(w)))
    vc = vc + dt * acc
    w = w + dt * wdot
    c = c + dt * vc
    if floor_z is not None and c[2] < floor_z:
        c[2] = floor_z
        vc[2] = max(vc[2], 0.0)
    quat = quat_mul(quat, quat_exp(w, dt))
    quat = quat / math.sqrt(float(quat.dot(quat)))
    R2 = p.rotmat(quat)
    pos = c - R2.dot(pp.r_com)
    vel = vc - R2.dot(np.cross(w, pp.r_com))
    return pos, quat, vel, w
