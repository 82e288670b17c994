"""``BaseControl`` facade (dronesim/control/BaseControl.py): constructor, ``reset`` and the
``computeControlFromState`` state slicing, executed by the CUDA core for ``num_envs`` vehicles of one
type at once (``num_envs=1`` reproduces the reference's shapes)."""
from __future__ import annotations

import numpy as np

from ..vehicles import LAW_6DOF, LAW_QUAD, load_vehicle


class BaseControl:
    LAW = None  # set by the subclasses

    def __init__(self, drone_model: str, g: float = 9.8, *, num_envs: int = 1, device: int = 0):
        from ..core import SwarmCore

        self.DRONE_MODEL = drone_model
        self.vehicle = load_vehicle(drone_model)
        v = self.vehicle
        if self.LAW is not None and v.law != self.LAW:
            print("[ERROR] in %s.__init__(), drone model '%s' has %d virtual controls: use dronesim_b200.control.%s"
                  % (type(self).__name__, drone_model, v.INDI_OUTPUT_NR, "INDIControl_6DOF" if v.law == LAW_6DOF else "INDIControl"))
            raise ValueError("controller / vehicle mismatch")
        #### the fields BaseControl / INDIControl expose (BaseControl.py:37-42, INDIControl.py:55-106)
        self.m, self.GRAVITY = v.M, g * v.M
        self.KF, self.KM = v.KF, v.KM
        self.G1 = np.array(v.G1)
        self.indi_actuator_nr, self.indi_output_nr = v.INDI_ACTUATOR_NR, v.INDI_OUTPUT_NR
        self.guidance_indi_pos_gain, self.guidance_indi_speed_gain = v.guidance_indi_pos_gain, v.guidance_indi_speed_gain
        self.PWM2RPM_SCALE, self.PWM2RPM_CONST = np.array(v.PWM2RPM_SCALE), np.array(v.PWM2RPM_CONST)
        self.MIN_PWM, self.MAX_PWM = np.array(v.MIN_PWM), np.array(v.MAX_PWM)
        self.NUM_ENVS = int(num_envs)
        self._core = SwarmCore([v], self.NUM_ENVS, device=device)
        self.reset()

    # ------------------------------------------------------------------
    def reset(self):
        """INDIControl.reset (INDIControl.py:109-146 / INDIControl_6DOF.py:214-251)."""
        self.control_counter = 0
        self._core.reset(np.zeros((self.NUM_ENVS, 3)))

    def close(self):
        self._core.close()

    def computeControlFromState(self, control_timestep, state, target_pos, target_vel=np.zeros(3), target_acc=np.zeros(3),
                                target_rpy=np.zeros(3), target_rpy_rates=np.zeros(3)):
        """BaseControl.computeControlFromState (BaseControl.py:61-103): ``state`` is the value of key
        "state" of the aviary obs, (16 + n_u,) for one vehicle or [num_envs, 16 + n_u] / [num_envs, 22]."""
        import torch

        E = self.NUM_ENVS
        dev = self._core.device
        s = torch.as_tensor(state, dtype=torch.float32, device=dev).reshape(E, -1)
        if s.shape[1] < 22:
            s = torch.nn.functional.pad(s, (0, 22 - s.shape[1]))
        return self._run(control_timestep, s.contiguous(), target_pos, target_vel, target_acc, target_rpy)

    def _compute(self, control_timestep, cur_pos, cur_quat, cur_vel, cur_ang_vel, target_pos, target_vel, target_acc, target_rpy):
        import torch

        E = self.NUM_ENVS
        dev = self._core.device
        s = torch.zeros((E, 22), dtype=torch.float32, device=dev)
        t = lambda a, n: torch.as_tensor(np.asarray(a, dtype=np.float32) if not torch.is_tensor(a) else a,  # noqa: E731
                                         dtype=torch.float32, device=dev).reshape(-1, n).expand(E, n)
        s[:, 0:3], s[:, 3:7], s[:, 10:13], s[:, 13:16] = t(cur_pos, 3), t(cur_quat, 4), t(cur_vel, 3), t(cur_ang_vel, 3)
        return self._run(control_timestep, s, target_pos, target_vel, target_acc, target_rpy)

    def _run(self, control_timestep, state_dev, target_pos, target_vel, target_acc, target_rpy):
        import torch

        E = self.NUM_ENVS
        dev = self._core.device

        def t(a, n):
            x = a if torch.is_tensor(a) else torch.from_numpy(np.asarray(a, dtype=np.float32))
            return x.to(device=dev, dtype=torch.float32).reshape(-1, n).expand(E, n)

        py = torch.cat([t(target_pos, 3), t(target_rpy, 3)[:, 2:3]], dim=1).contiguous()
        tg = self._core.targets_per_vehicle(py, vel=t(target_vel, 3).contiguous(), acc=t(target_acc, 3).contiguous())
        self.control_counter += 1
        cmd, pos_e, yaw_e = self._core.control_from_state(state_dev, tg, float(control_timestep))
        n_u = self.indi_actuator_nr
        if E == 1:  # the reference's return types: ndarray[n_u], ndarray[3], float
            return (cmd[0, :n_u].cpu().numpy().astype(np.float64), pos_e[0].cpu().numpy().astype(np.float64),
                    float(yaw_e[0].item()))
        return cmd[:, :n_u], pos_e, yaw_e

    # controller memory, as the reference attributes (first env for num_envs = 1)
    def _mem(self, key):
        v = self._core.views()[key].cpu().numpy().astype(np.float64)
        return v[0] if self.NUM_ENVS == 1 else v

    @property
    def last_vel(self):
        return self._mem("last_vel")

    @property
    def last_rates(self):
        return self._mem("last_rates")

    @property
    def last_thrust(self):
        return self._mem("last_thrust")

    @property
    def cmd(self):
        c = self._core.cmd().cpu().numpy().astype(np.float64)[:, : self.indi_actuator_nr]
        return c[0] if self.NUM_ENVS == 1 else c


__all__ = ["BaseControl", "LAW_QUAD", "LAW_6DOF"]
