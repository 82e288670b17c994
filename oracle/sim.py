"""The reference's example loop, restated over envs x drones (FP64, per-vehicle Python loops).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Loop structure restated from ``examples/fly_INDI.py:213-245`` (identical in
``fly_INDI_TrajectoryTrack.py:222-258`` and ``fly_hexa_6DOF.py:203-236``) and
``BaseAviary.step`` (dronesim/envs/BaseAviary.py:507-555):

    obs = env.step(action)          # clip action (CtrlAviary.py:258-263), K substeps with it
    action = ctrl.computeControlFromState(obs.state, targets[wp])   # on the fresh state (quirk Q5)
    wp = wp + 1 if wp < NUM_WP - 1 else 0                           # fly_INDI.py:242-245

Integer work restated bit-exactly: waypoint counter (above), ``step_counter += AGGR_PHY_STEPS``
(BaseAviary.py:554), adjacency ``||p_i - p_j|| < R`` strict (BaseAviary.py:913-921), and the
batched ``done`` predicate built on ``fly_INDI_TrajectoryTrack.py:249-250``.
"""
import math

import numpy as np

from . import control as oc
from . import dynamics as od
from . import pyb_math as p

DONE_GOAL = 1  # slot reached the goal sphere  (fly_INDI_TrajectoryTrack.py:249-250)
DONE_FLOOR = 2  # below z_min                   (beyond the reference; off by default)
DONE_TIME = 4  # step_counter >= max_steps     (beyond the reference; off by default)


class OracleSwarm:
    def __init__(self, slot_types, n_envs, integrator="quat", composite=True, gnd=False, drag=False,
                 dw=False, freq=240, aggregate_phy_steps=1, neighbourhood_radius=np.inf, motor_tau=0.0,
                 acc_filter_hz=0.0, noise_force_sigma=0.0, noise_torque_sigma=0.0, noise_seed=0, env_offset=0):
        """``slot_types``: list of VehicleType, one per drone slot of an env (the reference's
        ``drone_model`` list, BaseAviary.py:131,219); every env has the same slot->type map."""
        self.E, self.D = int(n_envs), len(slot_types)
        self.types = list(slot_types)
        self.pp = [od.PhysParams(vt, composite) for vt in self.types]
        self.integrator = integrator
        self.gnd, self.drag, self.dw = bool(gnd), bool(drag), bool(dw)
        self.SIM_FREQ = freq
        self.TIMESTEP = 1.0 / freq
        self.K = int(aggregate_phy_steps)
        self.radius = neighbourhood_radius
        self.n_u = [vt.INDI_ACTUATOR_NR for vt in self.types]
        self.ctrl = [[oc.make_controller(vt) for vt in self.types] for _ in range(self.E)]
        # extensions beyond the reference (north_star), off by default: first-order motor lag (time constant, s)
        # and a first-order low-pass (cut-off, Hz) on the controllers' angular-acceleration estimate
        self.motor_tau = float(motor_tau)
        self.motor_a = 1.0 - math.exp(-self.TIMESTEP / self.motor_tau) if self.motor_tau > 0 else None
        if acc_filter_hz > 0:
            b = 1.0 - math.exp(-2.0 * math.pi * float(acc_filter_hz) * self.K * self.TIMESTEP)
            for row in self.ctrl:
                for c in row:
                    c.acc_b = b
        # rotor noise: the reference's N(0, 0.01) / N(0, 0.001) draws, from a counter-based source (oracle/noise.py)
        self.noise_f, self.noise_m = float(noise_force_sigma), float(noise_torque_sigma)
        self.noise_seed, self.env_offset = int(noise_seed), int(env_offset)
        self.floor_z = None  # ground plane (DS_FLAG_GROUND_PLANE), quaternion integrator only
        self.goal = None
        self.goal_radius = 0.3
        self.z_min = None
        self.max_steps = None

    # ------------------------------------------------------------------ state
    def reset(self, pos0, rpy0=None, vel0=None):
        E, D = self.E, self.D
        self.pos = np.array(pos0, float).reshape(E, D, 3).copy()
        rpy0 = np.zeros((E, D, 3)) if rpy0 is None else np.array(rpy0, float).reshape(E, D, 3)
        self.quat = np.zeros((E, D, 4))
        self.rpy = np.zeros((E, D, 3))
        for e in range(E):
            for d in range(D):
                self.quat[e, d] = p.getQuaternionFromEuler(rpy0[e, d])  # BaseAviary.py:688
                self.rpy[e, d] = p.getEulerFromQuaternion(self.quat[e, d])  # :729
        self.vel = np.zeros((E, D, 3)) if vel0 is None else np.array(vel0, float).reshape(E, D, 3).copy()
        self.rates = np.zeros((E, D, 3))  # body rates
        self.last_clipped_action = np.zeros((E, D, 6))  # BaseAviary.py:659-662
        self.rpm = np.zeros((E, D, 6))  # motor model state: starts at the rpm of the all-zero action
        for d in range(D):
            self.rpm[:, d, : self.n_u[d]] = od.rpm_of_cmd(self.pp[d], np.zeros(self.n_u[d]))
        self.step_counter = 0
        self.done_bits = np.zeros((E, D), dtype=np.uint8)
        for e in range(E):
            for d in range(D):
                self.ctrl[e][d].reset()

    def ang_v_world(self, e, d):
        return p.rotmat(self.quat[e, d]).dot(self.rates[e, d])

    def state_vector(self, e, d):
        """BaseAviary._getDroneStateVector (BaseAviary.py:780-790)."""
        n = self.n_u[d]
        return np.hstack([self.pos[e, d], self.quat[e, d], self.rpy[e, d], self.vel[e, d],
                          self.ang_v_world(e, d), self.last_clipped_action[e, d, :n]])

    def adjacency_bits(self, margin=None):
        """BaseAviary._getAdjacencyMatrix (BaseAviary.py:901-921) as one bitmask per vehicle.

        With ``margin``: returns ``(bits, sure)`` where ``sure`` masks the pairs whose distance is farther than
        ``margin`` from the radius (the decisions an FP32 evaluation of the same positions must reproduce)."""
        out = np.zeros((self.E, self.D), dtype=np.uint32)
        sure = np.zeros((self.E, self.D), dtype=np.uint32)
        for e in range(self.E):
            for i in range(self.D):
                bits, s = 1 << i, 1 << i
                for j in range(self.D):
                    if j == i:
                        continue
                    d = np.linalg.norm(self.pos[e, i] - self.pos[e, j])
                    if d < self.radius:
                        bits |= 1 << j
                    if margin is not None and abs(d - self.radius) > margin:
                        s |= 1 << j
                out[e, i], sure[e, i] = bits, s
        return out if margin is None else (out, sure)

    # ------------------------------------------------------------------ physics
    def physics_step(self, action):
        """``BaseAviary.step`` minus the bookkeeping: ``action[E, D, 6]`` PWM (padded)."""
        action = np.asarray(action, float).reshape(self.E, self.D, -1)
        dt = self.TIMESTEP
        for e in range(self.E):
            clipped = np.zeros((self.D, 6))
            for d in range(self.D):
                n, pp = self.n_u[d], self.pp[d]
                clipped[d, :n] = np.clip(action[e, d, :n], pp.min_pwm, pp.max_pwm)  # CtrlAviary.py:258-263
            for _k in range(self.K):
                snap_pos = self.pos[e].copy()  # the state cache all drones read (BaseAviary.py:513-520)
                new = []
                for d in range(self.D):
                    n, pp = self.n_u[d], self.pp[d]
                    prev_sum = float(np.sum(od.rpm_of_cmd(pp, self.last_clipped_action[e, d, :n])))
                    others = [snap_pos[j] for j in range(self.D) if j != d] if self.dw else []
                    rpm = None
                    if self.motor_a is not None:  # R9: first-order lag towards the commanded rpm
                        prev_sum = float(np.sum(self.rpm[e, d, :n]))
                        self.rpm[e, d, :n] += self.motor_a * (od.rpm_of_cmd(pp, clipped[d, :n]) - self.rpm[e, d, :n])
                        rpm = self.rpm[e, d, :n]
                    noise = None
                    if self.noise_f > 0 or self.noise_m > 0:
                        from . import noise as on
                        nz = on.normals12((self.env_offset + e) * self.D + d, self.step_counter + _k, self.noise_seed)
                        noise = (self.noise_f * nz[0:n], self.noise_m * nz[6:6 + n])
                    F, tau, R = od.body_wrench(pp, clipped[d, :n], prev_sum, snap_pos[d], self.quat[e, d],
                                               self.rpy[e, d], self.vel[e, d], others, self.gnd, self.drag, self.dw, rpm=rpm,
                                               noise=noise)
                    if self.integrator == "rpy":
                        new.append(od.substep_rpy(pp, dt, snap_pos[d], self.quat[e, d], self.rpy[e, d],
                                                  self.vel[e, d], self.rates[e, d], F, tau, R))
                    else:
                        ps, q, v, w = od.substep_quat(pp, dt, snap_pos[d], self.quat[e, d], self.vel[e, d],
                                                      self.rates[e, d], F, tau, R, floor_z=self.floor_z)
                        new.append((ps, q, np.array(p.getEulerFromQuaternion(q)), v, w))
                for d in range(self.D):
                    self.pos[e, d], self.quat[e, d], self.rpy[e, d], self.vel[e, d], self.rates[e, d] = new[d]
                self.last_clipped_action[e] = clipped  # BaseAviary.py:545
        self.step_counter += self.K  # BaseAviary.py:554
        self._update_done()

    def _update_done(self):
        for e in range(self.E):
            for d in range(self.D):
                b = 0
                if self.goal is not None and np.linalg.norm(self.pos[e, d] - self.goal) < self.goal_radius:
                    b |= DONE_GOAL
                if self.z_min is not None and self.pos[e, d, 2] < self.z_min:
                    b |= DONE_FLOOR
                if self.max_steps is not None and self.step_counter >= self.max_steps:
                    b |= DONE_TIME
                self.done_bits[e, d] |= b

    def env_done(self):
        """done of an env: slot 0 reached the goal (the example tests drone "0" only), or any
        slot below the floor, or the time limit."""
        d0 = (self.done_bits[:, 0] & DONE_GOAL) != 0
        anyf = ((self.done_bits & (DONE_FLOOR | DONE_TIME)) != 0).any(axis=1)
        return d0 | anyf

    # ------------------------------------------------------------------ control
    def control_step(self, tpos, tvel=None, tacc=None, tyaw=None):
        """One ``computeControlFromState`` per vehicle; returns the new action ``[E, D, 6]``."""
        E, D = self.E, self.D
        dt = self.K * self.TIMESTEP  # CTRL_EVERY_N_STEPS * env.TIMESTEP  (fly_INDI.py:231)
        z3 = np.zeros((E, D, 3))
        tpos = np.asarray(tpos, float).reshape(E, D, 3)
        tvel = z3 if tvel is None else np.asarray(tvel, float).reshape(E, D, 3)
        tacc = z3 if tacc is None else np.asarray(tacc, float).reshape(E, D, 3)
        tyaw = np.zeros((E, D)) if tyaw is None else np.asarray(tyaw, float).reshape(E, D)
        action = np.zeros((E, D, 6))
        self.pos_e = np.zeros((E, D, 3))
        self.yaw_err = np.zeros((E, D))
        for e in range(E):
            for d in range(D):
                cmd, pe, ye = self.ctrl[e][d].computeControlFromState(
                    control_timestep=dt, state=self.state_vector(e, d), target_pos=tpos[e, d],
                    target_vel=tvel[e, d], target_acc=tacc[e, d], target_rpy=np.array([0.0, 0.0, tyaw[e, d]]))
                action[e, d, : self.n_u[d]] = cmd
                self.pos_e[e, d], self.yaw_err[e, d] = pe, ye
        return action


    # ------------------------------------------------------------------ controller-embedding aviaries
    def velocity_preprocess(self, vel_action):
        """``VelocityAviary._preprocessAction`` (VelocityAviary.py:221-264): ``vel_action[E, D, 4]`` =
        direction xyz + fraction of the speed limit -> PWM action ``[E, D, 6]``.  ``SPEED_LIMIT`` =
        ``MAX_SPEED_KMH * 1000 / 3600`` (VelocityAviary.py:92-94).  The reference instantiates the quad-law
        ``INDIControl`` for every drone; a 6-DOF airframe is flown by its own law here (the reference would
        fail on the 6x6 ``G1``)."""
        E, D = self.E, self.D
        va = np.asarray(vel_action, float).reshape(E, D, 4)
        dt = self.K * self.TIMESTEP  # control_timestep = AGGR_PHY_STEPS * TIMESTEP (:243)
        action = np.zeros((E, D, 6))
        for e in range(E):
            for d in range(D):
                state = self.state_vector(e, d)
                v = va[e, d]
                n = np.linalg.norm(v[0:3])
                v_unit = v[0:3] / n if n != 0 else np.zeros(3)  # :239-242
                speed_limit = self.types[d].MAX_SPEED_KMH * (1000 / 3600)
                cmd, _, _ = self.ctrl[e][d].computeControl(
                    control_timestep=dt, cur_pos=state[0:3], cur_quat=state[3:7], cur_vel=state[10:13],
                    cur_ang_vel=state[13:16], target_pos=state[0:3], target_rpy=np.array([0, 0, state[9]]),
                    target_vel=speed_limit * np.abs(v[3]) * v_unit)
                action[e, d, : self.n_u[d]] = cmd
        return action

    def rate_preprocess(self, rate_thrust):
        """``RPYTAviary._preprocessAction`` (RPYTAviary.py:180-193): ``rate_thrust[E, D, 4]`` = p, q, r
        set-point + thrust -> ``INDIControl._INDIRateControl`` (INDIControl.py:413-490) -> PWM action."""
        E, D = self.E, self.D
        rt = np.asarray(rate_thrust, float).reshape(E, D, 4)
        dt = self.K * self.TIMESTEP
        action = np.zeros((E, D, 6))
        for e in range(E):
            for d in range(D):
                state = self.state_vector(e, d)
                c = self.ctrl[e][d]
                c.cmd = c.rate_control(dt, rt[e, d, 3], state[3:7], state[13:16], rt[e, d, 0:3])
                action[e, d, : self.n_u[d]] = c.cmd
        return action


def next_waypoint(wp, num_wp):
    """fly_INDI.py:242-245."""
    return np.where(wp < num_wp - 1, wp + 1, 0)


def initial_waypoints(n, num_wp):
    """fly_INDI.py:161-163: ``int((i * NUM_WP / 6) % NUM_WP)``."""
    return np.array([int((i * num_wp / 6) % num_wp) for i in range(n)], dtype=np.int32)
