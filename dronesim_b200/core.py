"""``SwarmCore``: thin Python owner of one ``ds_handle`` (one GPU's shard of environments).

PyTorch is used for plumbing only (device buffers for targets / actions / observations, the
current CUDA stream, zero-copy views of the handle's resident state); every computation of the
hot path happens in ``libdronesim_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .vehicles import LAW_6DOF, VehicleType, load_vehicle


class _CudaView:
    """Minimal ``__cuda_array_interface__`` carrier for a borrowed device pointer."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner  # keeps the handle alive while the view exists


def type_params(vt: VehicleType, composite: bool) -> L.ds_type_params:
    """Pack a ``VehicleType`` into the C table.  ``composite=False`` uses the mass / inertia the
    reference's parser reads (first link, BaseAviary.py:2055-2069: the literal ``Physics.DYN``
    numbers); ``composite=True`` the whole kinematic tree (what PyBullet simulates)."""
    p = L.ds_type_params()
    n_u, n_v = vt.INDI_ACTUATOR_NR, vt.INDI_OUTPUT_NR
    p.n_u, p.n_v, p.law = n_u, n_v, vt.law
    p.rotor_model = 1 if "morphing_hexa" in vt.TYPE else 0  # which live force model flies it (BaseAviary.py:935-944)
    if "advanced" in vt.TYPE and p.rotor_model == 0:  # BaseAviary.py:1493
        from .vehicles import load_propeller

        prop = load_propeller()
        p.rotor_model = 2
        for i in range(14):
            p.adv_coeff[i] = prop["coeff"][i]
        p.adv_radius = prop["radius"]
    if composite:
        mass, J, rc = vt.M_TOTAL, np.asarray(vt.J_TOTAL, float), np.asarray(vt.COM, float)
    else:
        mass, J, rc = vt.M, np.asarray(vt.J, float), np.zeros(3)
    p.mass = float(mass)
    for i in range(9):
        p.J[i] = float(J.reshape(-1)[i])
    for i in range(3):
        p.r_com[i] = float(rc[i])
        p.drag_coeff[i] = float(vt.DRAG_COEFF[i])
        p.att_gain[i] = float(vt.att_gain[i])
        p.rate_gain[i] = float(vt.rate_gain[i])
    p.kf, p.km = float(vt.KF), float(vt.KM)
    for i in range(n_u):
        for k in range(3):
            p.rotor_pos[i][k] = float(vt.rotor_pos[i][k])
            p.rotor_axis[i][k] = float(vt.rotor_axis[i][k])
            p.torque_axis[i][k] = float(vt.torque_axis[i][k])
        p.rotor_spin[i] = float(vt.rotor_spin[i])
        p.pwm2rpm_scale[i] = float(vt.PWM2RPM_SCALE[i])
        p.pwm2rpm_const[i] = float(vt.PWM2RPM_CONST[i])
        p.min_pwm[i] = float(vt.MIN_PWM[i])
        p.max_pwm[i] = float(vt.MAX_PWM[i])
    p.gnd_eff_coeff, p.prop_radius, p.gnd_eff_h_clip = float(vt.GND_EFF_COEFF), float(vt.PROP_RADIUS), float(vt.GND_EFF_H_CLIP)
    p.dw_coeff[0], p.dw_coeff[1], p.dw_coeff[2] = float(vt.DW_COEFF_1), float(vt.DW_COEFF_2), float(vt.DW_COEFF_3)
    p.kp_pos, p.kd_pos = float(vt.guidance_indi_pos_gain), float(vt.guidance_indi_speed_gain)
    alloc = vt.wls_unconstrained() if vt.law == LAW_6DOF else vt.pinv_alloc()  # [n_u][n_v]
    for i in range(n_v):
        for j in range(n_u):
            p.G1[i][j] = float(vt.G1[i][j])
    for i in range(n_u):
        for j in range(n_v):
            p.alloc[i][j] = float(alloc[i, j])
    for i in range(n_v):
        p.wls_wv[i] = float(vt.WLS_WV[i]) if n_v == 6 else float([1000.0, 1000.0, 0.1, 10.0][i])
    p.wls_gamma = float(vt.WLS_GAMMA)
    # controller reset values: INDIControl.py:127-129 vs INDIControl_6DOF.py:232-234
    p.init_cmd = 0.5 if vt.law == LAW_6DOF else 0.0
    p.init_thrust = 0.3 if vt.law == LAW_6DOF else 0.0
    p.max_speed_kmh = float(vt.MAX_SPEED_KMH)
    return p


class SwarmCore:
    """``n_envs`` environments x ``len(slot_models)`` drones resident on one GPU."""

    def __init__(self, slot_models: Sequence, n_envs: int, *, integrator: str = "quat", composite: Optional[bool] = None,
                 ground: bool = False, drag: bool = False, downwash: bool = False, stats: bool = False,
                 freq: float = 240.0, aggregate_phy_steps: int = 1, neighbourhood_radius: float = math.inf,
                 gravity: float = 9.8, goal=None, goal_radius: float = 0.3, z_min=None, max_steps: int = 0,
                 device: int = 0, env_offset: int = 0, assets_dir: Optional[str] = None, dw_ordered_pairs: bool = False,
                 types_in_smem: bool = False, ground_plane_z: Optional[float] = None, debug_redzones: bool = False,
                 motor_tau: float = 0.0, acc_filter_hz: float = 0.0, reward_mode: int = 0,
                 noise_force_sigma: float = 0.0, noise_torque_sigma: float = 0.0, noise_seed: int = 0):
        lib = L.lib()
        self.vehicle_types: List[VehicleType] = [m if isinstance(m, VehicleType) else load_vehicle(m, assets_dir)
                                                 for m in slot_models]
        self.D, self.E = len(self.vehicle_types), int(n_envs)
        self.N = self.D * self.E
        self.K = int(aggregate_phy_steps)
        self.SIM_FREQ = float(freq)
        self.integrator = integrator
        if composite is None:
            composite = integrator == "quat"
        self.composite = composite
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        # guard bands around every device buffer of the handle (check_redzones(); close() raises if one was overwritten);
        # DRONESIM_B200_REDZONES=1 turns them on for every core of the process (how the GPU test-suite is run once per change)
        self._redzones = bool(debug_redzones) or os.environ.get("DRONESIM_B200_REDZONES") == "1"
        cfg = L.ds_config()
        cfg.n_envs, cfg.drones_per_env, cfg.substeps = self.E, self.D, self.K
        cfg.integrator = L.DS_INTEG_RPY if integrator == "rpy" else L.DS_INTEG_QUAT
        cfg.flags = ((L.DS_FLAG_GROUND if ground else 0) | (L.DS_FLAG_DRAG if drag else 0)
                     | (L.DS_FLAG_DOWNWASH if downwash else 0) | (L.DS_FLAG_STATS if stats else 0)
                     | (L.DS_FLAG_DW_ORDERED_PAIRS if dw_ordered_pairs else 0)
                     | (L.DS_FLAG_TYPES_IN_SMEM if types_in_smem else 0)
                     | (L.DS_FLAG_GROUND_PLANE if ground_plane_z is not None else 0)
                     | (L.DS_FLAG_DEBUG_REDZONES if self._redzones else 0))
        cfg.ground_plane_z = float(ground_plane_z) if ground_plane_z is not None else 0.0
        cfg.device, cfg.sim_freq, cfg.gravity = self.device_index, float(freq), float(gravity)
        cfg.neighbourhood_radius = float(neighbourhood_radius)
        if goal is not None:
            cfg.done_goal_enable = 1
            cfg.goal[0], cfg.goal[1], cfg.goal[2] = [float(x) for x in goal]
        cfg.goal_radius = float(goal_radius)
        if z_min is not None:
            cfg.done_floor_enable, cfg.z_min = 1, float(z_min)
        cfg.max_steps, cfg.env_offset = int(max_steps), int(env_offset)
        # extensions beyond the reference (north_star): first-order motor lag, low-passed angular acceleration, tracking reward
        cfg.motor_tau, cfg.acc_filter_hz, cfg.reward_mode = float(motor_tau), float(acc_filter_hz), int(reward_mode)
        cfg.noise_force_sigma, cfg.noise_torque_sigma = float(noise_force_sigma), float(noise_torque_sigma)
        cfg.noise_seed = int(noise_seed) & 0xFFFFFFFFFFFFFFFF
        self._h = C.c_void_p()
        L.check(lib.ds_create(C.byref(cfg), C.byref(self._h)))
        # distinct types, in order of first appearance
        names, self.slot_type = [], []
        for vt in self.vehicle_types:
            if vt.name not in names:
                names.append(vt.name)
            self.slot_type.append(names.index(vt.name))
        self.type_names = names
        uniq = [next(v for v in self.vehicle_types if v.name == n) for n in names]
        arr = (L.ds_type_params * len(uniq))(*[type_params(v, composite) for v in uniq])
        st = (C.c_uint8 * self.D)(*self.slot_type)
        L.check(lib.ds_set_types(self._h, arr, len(uniq), st), self._h)
        self.n_u = [vt.INDI_ACTUATOR_NR for vt in self.vehicle_types]
        self._views = None
        self._keep = []  # device tensors referenced by in-flight target structs

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def check_redzones(self) -> int:
        """Guard-band bytes around the handle's device buffers that a kernel has overwritten (cores created with
        ``debug_redzones=True`` / DRONESIM_B200_REDZONES=1; synchronises the device)."""
        bad = C.c_int64(0)
        L.check(L.lib().ds_debug_check_redzones(self._h, C.byref(bad)), self._h)
        return int(bad.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            bad = self.check_redzones() if getattr(self, "_redzones", False) else 0
            L.lib().ds_destroy(self._h)
            self._h = C.c_void_p()
            self._views = None
            if bad:
                raise RuntimeError("dronesim_b200: %d guard-band bytes around the core's device buffers were overwritten" % bad)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _host_f32(a, shape):
        if a is None:
            return None
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(shape))
        return a

    def reset(self, pos0, rpy0=None, vel0=None, action0=None, wp0=None):
        """BaseAviary.reset + INDIControl.reset for every vehicle.  Host arrays, [E, D, ...] or [N, ...]."""
        N = self.N
        pos0 = self._host_f32(pos0, (N, 3))
        rpy0 = self._host_f32(rpy0, (N, 3))
        vel0 = self._host_f32(vel0, (N, 3))
        action0 = self._host_f32(action0, (N, 6))
        wp = None if wp0 is None else np.ascontiguousarray(np.asarray(wp0, dtype=np.int32).reshape(N))
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        with torch.cuda.device(self.device):
            L.check(L.lib().ds_reset(self._h, ptr(pos0), ptr(rpy0), ptr(vel0), ptr(action0), ptr(wp), self._stream()),
                    self._h)

    def reset_envs(self, mask, pos0, rpy0=None, vel0=None, action0=None, wp0=None):
        """``BaseAviary.reset`` for the envs with ``mask[e] != 0`` only, on the device, no host synchronisation
        (``ds_reset_envs``).  ``mask`` [E] uint8 / bool; ``pos0`` [N,3] (+ optional ``rpy0``, ``vel0`` [N,3], ``action0`` [N,6],
        ``wp0`` [N] int32): device tensors, or host arrays that are copied first; rows of unmasked envs are not read."""
        N = self.N

        def dev(a, shape, dtype):
            if a is None:
                return None
            t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
            return t.to(self.device, dtype).reshape(shape).contiguous()

        m = dev(mask, (self.E,), torch.uint8)
        p0, r0, v0 = dev(pos0, (N, 3), torch.float32), dev(rpy0, (N, 3), torch.float32), dev(vel0, (N, 3), torch.float32)
        a0, w0 = dev(action0, (N, 6), torch.float32), dev(wp0, (N,), torch.int32)
        L.check(L.lib().ds_reset_envs(self._h, self._p(m), self._p(p0), self._p(r0), self._p(v0), self._p(a0), self._p(w0),
                                      self._stream()), self._h)
        self._reset_keep = (m, p0, r0, v0, a0, w0)  # borrowed until the stream has consumed them

    # ------------------------------------------------------------------ targets
    def _dev4(self, a, name):
        """[N,4] float32 contiguous device tensor (numpy / torch in, padded from [N,3] if needed)."""
        if a is None:
            return None
        t = torch.as_tensor(a, dtype=torch.float32, device=self.device) if not (
            isinstance(a, torch.Tensor) and a.dtype == torch.float32 and a.device == self.device) else a
        t = t.reshape(self.N, -1)
        if t.shape[1] == 3:
            t = torch.nn.functional.pad(t, (0, 1))
        if t.shape[1] != 4:
            raise ValueError("%s must have 3 or 4 columns" % name)
        return t.contiguous()

    def targets_per_vehicle(self, pos_yaw, vel=None, acc=None) -> L.ds_targets:
        """mode 0: ``pos_yaw`` [N,4] (x,y,z,yaw), optional ``vel`` / ``acc`` [N,3|4]."""
        t = L.ds_targets()
        p, v, a = self._dev4(pos_yaw, "pos_yaw"), self._dev4(vel, "vel"), self._dev4(acc, "acc")
        t.mode = 0
        t.pos_yaw = p.data_ptr()
        t.vel = v.data_ptr() if v is not None else None
        t.acc = a.data_ptr() if a is not None else None
        t._keep = (p, v, a)
        return t

    def targets_velocity(self, vel_action) -> L.ds_targets:
        """mode 2: ``vel_action`` [N,4] = VelocityAviary actions (direction xyz, fraction of the speed limit)."""
        t = L.ds_targets()
        v = self._dev4(vel_action, "vel_action")
        t.mode = 2
        t.vel = v.data_ptr()
        t._keep = (v,)
        return t

    def targets_rate_thrust(self, rate_thrust) -> L.ds_targets:
        """mode 3: ``rate_thrust`` [N,4] = RPYTAviary actions (p, q, r set-point, thrust); quad-law types only."""
        t = L.ds_targets()
        v = self._dev4(rate_thrust, "rate_thrust")
        t.mode = 3
        t.vel = v.data_ptr()
        t._keep = (v,)
        return t

    def targets_table(self, table, offset=None, advance: bool = True, wp=None) -> L.ds_targets:
        """mode 1: ``table`` [num_wp, 10] = pos3, vel3, acc3, yaw (the layout of the reference examples'
        TARGET_POS/VEL/ACC/RPYS rows), shared by all vehicles; per-vehicle counters live in the state, unless ``wp``
        ([N] int32 device tensor) supplies this step's index of every vehicle (fly_INDI.py:230-245)."""
        tab = np.asarray(table, dtype=np.float64)
        rows = np.zeros((tab.shape[0], 12), dtype=np.float32)
        rows[:, 0:3], rows[:, 3] = tab[:, 0:3], tab[:, 9]
        rows[:, 4:7], rows[:, 8:11] = tab[:, 3:6], tab[:, 6:9]
        d = torch.from_numpy(rows).to(self.device)
        off = self._dev4(offset, "offset")
        t = L.ds_targets()
        t.mode, t.num_wp, t.advance_wp = 1, int(tab.shape[0]), 1 if advance else 0
        t.table = d.data_ptr()
        t.offset = off.data_ptr() if off is not None else None
        if wp is not None:
            wp = wp.to(self.device, torch.int32).reshape(self.N).contiguous()
            t.wp = wp.data_ptr()
        t._keep = (d, off, wp)
        return t

    # ------------------------------------------------------------------ stepping
    def step(self, targets: L.ds_targets, n_control_steps: int = 1, order: int = L.DS_ORDER_PHYSICS_THEN_CONTROL):
        """The fused hot path: n x (K physics substeps + one INDI evaluation)."""
        L.check(L.lib().ds_step(self._h, C.byref(targets), int(n_control_steps), int(order), self._stream()), self._h)

    def physics_step(self, action: torch.Tensor):
        """``BaseAviary.step`` with an external action, device tensor [N, 6] (PWM)."""
        a = action.reshape(self.N, 6)
        assert a.is_cuda and a.dtype == torch.float32 and a.is_contiguous()
        L.check(L.lib().ds_physics_step(self._h, C.c_void_p(a.data_ptr()), self._stream()), self._h)

    def _outs(self, want_cmd=True, want_aux=True):
        cmd = torch.empty((self.N, 6), dtype=torch.float32, device=self.device) if want_cmd else None
        pe = torch.empty((self.N, 3), dtype=torch.float32, device=self.device) if want_aux else None
        ye = torch.empty((self.N,), dtype=torch.float32, device=self.device) if want_aux else None
        return cmd, pe, ye

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def control_step(self, targets, control_timestep: float):
        cmd, pe, ye = self._outs()
        L.check(L.lib().ds_control_step(self._h, C.byref(targets), C.c_float(control_timestep), self._p(cmd), self._p(pe),
                                        self._p(ye), self._stream()), self._h)
        return cmd, pe, ye

    def control_from_state(self, state: torch.Tensor, targets, control_timestep: float):
        s = state.reshape(self.N, L.DS_OBS_STRIDE)
        assert s.is_cuda and s.dtype == torch.float32 and s.is_contiguous()
        cmd, pe, ye = self._outs()
        L.check(L.lib().ds_control_from_state(self._h, C.c_void_p(s.data_ptr()), C.byref(targets),
                                              C.c_float(control_timestep), self._p(cmd), self._p(pe), self._p(ye),
                                              self._stream()), self._h)
        return cmd, pe, ye

    def rate_control_step(self, rate_thrust: torch.Tensor, control_timestep: float):
        r = rate_thrust.reshape(self.N, 4)
        assert r.is_cuda and r.dtype == torch.float32 and r.is_contiguous()
        cmd, _, _ = self._outs(True, False)
        L.check(L.lib().ds_rate_control_step(self._h, C.c_void_p(r.data_ptr()), C.c_float(control_timestep), self._p(cmd),
                                             self._stream()), self._h)
        return cmd

    def get_obs(self, state=True, neighbors=True, done=True, reward=False):
        obs = torch.empty((self.N, L.DS_OBS_STRIDE), dtype=torch.float32, device=self.device) if state else None
        nb = torch.empty((self.N,), dtype=torch.int32, device=self.device) if neighbors else None
        dn = torch.empty((self.E,), dtype=torch.uint8, device=self.device) if done else None
        rw = torch.empty((self.E,), dtype=torch.float32, device=self.device) if reward else None
        L.check(L.lib().ds_get_obs(self._h, self._p(obs), self._p(nb), self._p(dn), self._p(rw), self._stream()), self._h)
        return obs, nb, dn, rw

    def set_env_outputs(self, done: bool = True, reward: bool = True):
        """Arm per-env ``done [E] uint8`` / ``reward [E] float32`` device tensors that every following ``step`` fills
        (inside the fused kernel, by warp shuffles, when drones_per_env divides 32).  Returns the two tensors."""
        self._env_done = torch.zeros((self.E,), dtype=torch.uint8, device=self.device) if done else None
        self._env_reward = torch.zeros((self.E,), dtype=torch.float32, device=self.device) if reward else None
        L.check(L.lib().ds_set_env_outputs(self._h, self._p(self._env_done), self._p(self._env_reward)), self._h)
        return self._env_done, self._env_reward

    def step_host(self, host_pos_yaw: torch.Tensor, host_obs: Optional[torch.Tensor], host_done: Optional[torch.Tensor]):
        """End-to-end control step with HOST (ideally pinned) buffers; synchronises."""
        L.check(L.lib().ds_step_host(self._h, C.c_void_p(host_pos_yaw.data_ptr()),
                                     self._p(host_obs), self._p(host_done), self._stream()), self._h)

    def rollout_host(self, host_pos_yaw: torch.Tensor, host_done: Optional[torch.Tensor] = None):
        """``host_pos_yaw`` [T, N, 4] pinned float32, ``host_done`` [T, E] pinned uint8 or None: T control steps with the
        H2D / D2H copies pipelined against the compute (``ds_rollout_host``); synchronises."""
        T = int(host_pos_yaw.shape[0])
        assert host_pos_yaw.dtype == torch.float32 and host_pos_yaw.is_contiguous() and host_pos_yaw.numel() == T * self.N * 4
        L.check(L.lib().ds_rollout_host(self._h, C.c_void_p(host_pos_yaw.data_ptr()), T, self._p(host_done), self._stream()),
                self._h)

    def rollout_host_table(self, targets: L.ds_targets, host_wp: torch.Tensor, host_done: Optional[torch.Tensor] = None):
        """``targets``: a ``targets_table`` (device-resident table + optional per-vehicle offsets); ``host_wp`` [T, N] pinned
        int32 = the waypoint index of every vehicle for each of T control steps (``ds_rollout_host_table``); synchronises."""
        T = int(host_wp.shape[0])
        assert host_wp.dtype == torch.int32 and host_wp.is_contiguous() and host_wp.numel() == T * self.N
        L.check(L.lib().ds_rollout_host_table(self._h, C.byref(targets), C.c_void_p(host_wp.data_ptr()), T, self._p(host_done),
                                              self._stream()), self._h)

    # ------------------------------------------------------------------ state access
    def views(self) -> dict:
        """Zero-copy torch views of the resident state (valid until ``close``)."""
        v = L.ds_state_views()
        L.check(L.lib().ds_views(self._h, C.byref(v)), self._h)
        npad, n = int(v.n_pad), int(v.n)

        def f4(ptr, cols=4):
            return torch.as_tensor(_CudaView(ptr, (npad, cols), "<f4", self), device=self.device)[:n]

        pos_t, quat, vel_r, om_w = f4(v.pos_thrust), f4(v.quat), f4(v.vel_rpm), f4(v.omega_wp)
        lv_d, lr_e, c0, c1 = f4(v.lastvel_done), f4(v.lastrates_err), f4(v.cmd0123), f4(v.cmd45, 2)
        ext = {}
        if v.rpm0123:
            ext = {"rpm": torch.cat([f4(v.rpm0123), f4(v.rpm45, 2)], dim=1), "ang_acc_filt": f4(v.ang_acc_filt)[:, :3]}
        return {
            **ext,
            "pos": pos_t[:, :3], "last_thrust": pos_t[:, 3], "quat": quat, "vel": vel_r[:, :3], "rpm_sum": vel_r[:, 3],
            "omega_body": om_w[:, :3], "wp_counter": om_w.view(torch.int32)[:, 3], "last_vel": lv_d[:, :3],
            "done_bits": lv_d.view(torch.int32)[:, 3] & 0x7FFFFFFF, "last_rates": lr_e[:, :3], "pos_err": lr_e[:, 3],
            "cmd0123": c0, "cmd45": c1, "step_counter": int(v.step_counter),
        }

    # ------------------------------------------------------------------ checkpoint / resume
    _STATE_KEYS = ("pos_thrust", "quat", "vel_rpm", "omega_wp", "lastvel_done", "lastrates_err", "cmd0123", "cmd45",
                   "rpm0123", "rpm45", "ang_acc_filt")

    def _raw_views(self):
        v = L.ds_state_views()
        L.check(L.lib().ds_views(self._h, C.byref(v)), self._h)
        npad = int(v.n_pad)
        out = {}
        for k in self._STATE_KEYS:
            ptr = getattr(v, k)
            if ptr:
                cols = 2 if k in ("cmd45", "rpm45") else 4
                out[k] = torch.as_tensor(_CudaView(ptr, (npad, cols), "<f4", self), device=self.device)
        return out, int(v.step_counter)

    def state_dict(self) -> dict:
        """The resident structure-of-arrays state (raw float32 words, integer fields included) + the step counter, on the
        host: everything a rollout needs to continue bit-exactly in another handle of the same configuration."""
        torch.cuda.current_stream(self.device).synchronize()
        raw, sc = self._raw_views()
        return {"step_counter": sc, **{k: t.cpu().clone() for k, t in raw.items()}}

    def load_state_dict(self, sd: dict):
        """Restore a ``state_dict`` into this handle (after ``reset``, which sizes everything)."""
        raw, _ = self._raw_views()
        for k, t in raw.items():
            t.copy_(sd[k].to(self.device))
        L.check(L.lib().ds_set_step_counter(self._h, int(sd["step_counter"])), self._h)

    def cmd(self) -> torch.Tensor:
        v = self.views()
        return torch.cat([v["cmd0123"], v["cmd45"]], dim=1)

    @property
    def step_counter(self) -> int:
        v = L.ds_state_views()
        L.check(L.lib().ds_views(self._h, C.byref(v)), self._h)
        return int(v.step_counter)

    def stats(self) -> dict:
        out = (C.c_double * L.DS_NUM_STATS)()
        L.check(L.lib().ds_stats(self._h, out, L.DS_NUM_STATS, self._stream()), self._h)
        keys = ["control_evals", "sum_pos_err_sq", "saturated_cmds", "wls_slow_path", "wls_non_converged", "non_finite",
                "min_altitude", "done_vehicles"]
        return {k: out[i] for i, k in enumerate(keys)}

    def stats_reset(self):
        L.check(L.lib().ds_stats_reset(self._h, self._stream()), self._h)

    def debug_wls(self, type_id: int, v: torch.Tensor, cmd: torch.Tensor, force_slow: bool = False, working_set: bool = False):
        """Diagnostic: the WLS allocator alone on n problems -> (du [n,6], iterations [n][, W [n,6]])."""
        v = v.to(self.device, torch.float32).contiguous()
        cmd = cmd.to(self.device, torch.float32).contiguous()
        n = v.shape[0]
        du = torch.empty((n, 6), dtype=torch.float32, device=self.device)
        it = torch.empty((n,), dtype=torch.int32, device=self.device)
        W = torch.zeros((n, 6), dtype=torch.int32, device=self.device)
        L.check(L.lib().ds_debug_wls(self._h, int(type_id), self._p(v), self._p(cmd), self._p(du), self._p(it), self._p(W), n,
                                     1 if force_slow else 0, self._stream()), self._h)
        return (du, it, W) if working_set else (du, it)

    # ------------------------------------------------------------------ trajectory capture (Logger layout)
    def log_attach(self, vehicles, capacity: int):
        """Record the state vector of ``vehicles`` (ids ``env * D + slot``) after every control / physics step."""
        ids = np.ascontiguousarray(np.asarray(vehicles, dtype=np.int32).reshape(-1))
        self._log_n, self._log_cap = int(ids.shape[0]), int(capacity)
        L.check(L.lib().ds_log_attach(self._h, ids.ctypes.data_as(C.c_void_p), self._log_n, self._log_cap), self._h)

    def log_read(self):
        """-> (timestamps [count], states [n_vehicles, 22, count]) as float64 numpy (Logger.timestamps / Logger.states)."""
        n, cap = getattr(self, "_log_n", 0), getattr(self, "_log_cap", 0)
        states = np.zeros((n, L.DS_OBS_STRIDE, cap), dtype=np.float32)
        ts = np.zeros((max(cap, 1),), dtype=np.float64)
        cnt = C.c_int32(0)
        L.check(L.lib().ds_log_read(self._h, states.ctypes.data_as(C.c_void_p) if n else None,
                                    ts.ctypes.data_as(C.c_void_p), C.byref(cnt), self._stream()), self._h)
        return ts[: cnt.value].copy(), states[:, :, : cnt.value].astype(np.float64)

    def launch_count(self) -> int:
        return int(L.lib().ds_launch_count(self._h))
