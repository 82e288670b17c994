"""``INDIControl`` facade - the quad / 4-virtual-control INDI law (dronesim/control/INDIControl.py).

Position loop :232-351, attitude loop :355-411, rate loop + ``pinv(G1/0.05)`` allocation :413-490,
all evaluated by ``ds_control_from_state`` in ``libdronesim_b200.so``.
"""
from __future__ import annotations

import numpy as np

from .BaseControl import LAW_QUAD, BaseControl


class INDIControl(BaseControl):
    LAW = LAW_QUAD

    def computeControl(self, control_timestep, cur_pos, cur_quat, cur_vel, cur_ang_vel, target_pos, target_vel=np.zeros(3),
                       target_acc=np.zeros(3), target_rpy=np.zeros(3), target_rpy_rates=np.zeros(3)):
        """Argument order of INDIControl.py:154-166.  Returns ``(cmd[n_u] PWM, pos_e[3], yaw_err)``."""
        return self._compute(control_timestep, cur_pos, cur_quat, cur_vel, cur_ang_vel, target_pos, target_vel, target_acc,
                             target_rpy)

    def rateControl(self, control_timestep, rate_sp, thrust):
        """``_INDIRateControl`` alone (INDIControl.py:413-490) on the controller's resident state: the
        RPYTAviary entry (RPYTAviary.py:180-193).  ``rate_sp`` [3] or [E, 3], ``thrust`` scalar or [E]."""
        import torch

        E, dev = self.NUM_ENVS, self._core.device
        rt = torch.zeros((E, 4), dtype=torch.float32, device=dev)
        rt[:, :3] = torch.as_tensor(np.asarray(rate_sp, dtype=np.float32), device=dev).reshape(-1, 3)
        rt[:, 3] = torch.as_tensor(np.asarray(thrust, dtype=np.float32), device=dev).reshape(-1)
        cmd = self._core.rate_control_step(rt, float(control_timestep))[:, : self.indi_actuator_nr]
        return cmd[0].cpu().numpy().astype(np.float64) if E == 1 else cmd
