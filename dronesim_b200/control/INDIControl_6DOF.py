"""``INDIControl`` (6-DOF) facade - hexarotor law with 6 virtual controls and WLS allocation
(dronesim/control/INDIControl_6DOF.py:259-634, dronesim/control/wls_alloc.py:125-350).

Same class name as the reference module (``from ...INDIControl_6DOF import INDIControl``).
"""
from __future__ import annotations

import numpy as np

from .BaseControl import LAW_6DOF, BaseControl


class INDIControl(BaseControl):
    LAW = LAW_6DOF

    def computeControl(self, control_timestep, cur_pos, cur_quat, cur_vel, cur_ang_vel, target_pos, target_rpy=np.zeros(3),
                       target_vel=np.zeros(3), target_rpy_rates=np.zeros(3), target_acc=np.zeros(3)):
        """Argument order of INDIControl_6DOF.py:259-270 (differs from the quad module: call by keyword)."""
        return self._compute(control_timestep, cur_pos, cur_quat, cur_vel, cur_ang_vel, target_pos, target_vel, target_acc,
                             target_rpy)
