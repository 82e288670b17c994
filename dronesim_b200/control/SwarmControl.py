"""``SwarmINDIControl``: the controllers of a whole (possibly mixed) drone list in ONE object and one launch per step.

The reference builds one controller object per drone and calls them in a Python loop
(``ctrl = [INDIControl(drone_model=d) for d in ARGS.drone]`` ... ``for j in range(num_drones): action[str(j)], _, _ =
ctrl[j].computeControlFromState(...)``, examples/fly_INDI.py:211, 229-240).  The per-drone facades of this package keep
that API - each owns a one-type core, so a D-drone loop costs D launches and D device-to-host copies per step.  This class
is the batched form of the same call: every drone of every env in one ``ds_control_from_state`` launch, the law
(``INDIControl`` or ``INDIControl_6DOF``) chosen per slot from the URDF, controller memory resident on the device.
"""
from __future__ import annotations

import numpy as np

from ..vehicles import load_vehicle


class SwarmINDIControl:
    def __init__(self, drone_model: list, g: float = 9.8, *, num_envs: int = 1, device: int = 0):
        from ..core import SwarmCore

        self.DRONE_MODEL = list(drone_model)
        self.vehicles = [load_vehicle(d) for d in self.DRONE_MODEL]
        self.NUM_DRONES, self.NUM_ENVS = len(self.vehicles), int(num_envs)
        self.indi_actuator_nr = [v.INDI_ACTUATOR_NR for v in self.vehicles]
        self._core = SwarmCore(self.vehicles, self.NUM_ENVS, device=device)
        self.reset()

    def reset(self):
        """``INDIControl.reset`` of every controller (INDIControl.py:109-146 / INDIControl_6DOF.py:214-251)."""
        self.control_counter = 0
        self._core.reset(np.zeros((self.NUM_ENVS, self.NUM_DRONES, 3)))

    def close(self):
        self._core.close()

    def computeControlFromState(self, control_timestep, state, target_pos, target_vel=None, target_acc=None, target_rpy=None):
        """``BaseControl.computeControlFromState`` (BaseControl.py:61-103) for all drones at once.

        ``state``: the aviary's state tensor ``[E, D, 22]`` (``env.state_tensor()`` / ``obs["state"]``), or the reference's obs
        dict ``{str(j): {"state": ...}}`` of one env; ``target_*``: ``[E, D, 3]``, ``[D, 3]`` or ``[3]`` (broadcast).
        Returns ``(cmd [E, D, 6], pos_e [E, D, 3], yaw_err [E, D])`` as device tensors; for one env and a dict input,
        the reference's action dict ``{str(j): ndarray[n_u_j]}`` and numpy arrays."""
        import torch

        E, D, dev = self.NUM_ENVS, self.NUM_DRONES, self._core.device
        as_dict = isinstance(state, dict)
        if as_dict:
            s = np.zeros((1, D, 22), dtype=np.float32)
            for j in range(D):
                v = np.asarray(state[str(j)]["state"], dtype=np.float32)
                s[0, j, : v.shape[0]] = v
            state = s
        st = torch.as_tensor(state, dtype=torch.float32, device=dev).reshape(E * D, -1)
        if st.shape[1] < 22:
            st = torch.nn.functional.pad(st, (0, 22 - st.shape[1]))

        def t(a, n):
            if a is None:
                return None
            x = a if torch.is_tensor(a) else torch.from_numpy(np.asarray(a, dtype=np.float32))
            x = x.to(device=dev, dtype=torch.float32)
            return x.reshape((-1, D, n) if x.dim() >= 2 and x.numel() >= D * n else (1, 1, n)).expand(E, D, n).reshape(E * D, n)

        rpy = t(target_rpy, 3)
        yaw = rpy[:, 2:3] if rpy is not None else torch.zeros((E * D, 1), dtype=torch.float32, device=dev)
        py = torch.cat([t(target_pos, 3), yaw], dim=1).contiguous()
        v, a = t(target_vel, 3), t(target_acc, 3)
        tg = self._core.targets_per_vehicle(py, vel=None if v is None else v.contiguous(), acc=None if a is None else a.contiguous())
        self.control_counter += 1
        cmd, pos_e, yaw_e = self._core.control_from_state(st.contiguous(), tg, float(control_timestep))
        cmd, pos_e, yaw_e = cmd.view(E, D, 6), pos_e.view(E, D, 3), yaw_e.view(E, D)
        if as_dict and E == 1:
            c = cmd[0].cpu().numpy().astype(np.float64)
            return ({str(j): c[j, : self.indi_actuator_nr[j]] for j in range(D)}, pos_e[0].cpu().numpy().astype(np.float64),
                    yaw_e[0].cpu().numpy().astype(np.float64))
        return cmd, pos_e, yaw_e
