"""``VelocityAviary`` facade (dronesim/envs/VelocityAviary.py): velocity-command actions.

* action: ``{str(i): [vx, vy, vz, fraction of the speed limit]}`` (VelocityAviary.py:92-117);
* ``_preprocessAction`` (VelocityAviary.py:221-264): per drone, INDI ``computeControl`` with
  ``target_pos`` = the current position, ``target_rpy`` = (0, 0, current yaw) and
  ``target_vel = SPEED_LIMIT * |a[3]| * a[0:3] / ||a[0:3]||`` (zero direction -> zero velocity),
  ``control_timestep = AGGR_PHY_STEPS * TIMESTEP``; the resulting PWM command is what the
  AGGR_PHY_STEPS substeps apply (BaseAviary.py:507-545) and what the obs tail reports;
* obs / reward / done / info as CtrlAviary (VelocityAviary.py:121-216, 274-313).

Control and physics run in ONE fused kernel launch (``ds_step`` with target mode 2, order
control-then-physics); the controllers' memory (``last_vel``, ``last_rates``, ``last_thrust``, ``cmd``) lives
in the core's resident state, so ``env.ctrl`` is not a list of Python objects here.

Deviation: the reference builds an ``INDIControl`` (quad law) for every drone, which cannot run for a
6-rotor URDF (4-vector against a 6x6 ``G1``); here a 6-DOF airframe flies its own ``INDIControl_6DOF`` law.
"""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from .BaseAviary import Physics  # noqa: F401
from .CtrlAviary import CtrlAviary, _Box


class VelocityAviary(CtrlAviary):
    """Multi-drone environment class for high-level planning (velocity commands)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        #### Set a limit on the maximum target speed (VelocityAviary.py:92-94)
        self.SPEED_LIMIT = [d.MAX_SPEED_KMH * (1000 / 3600) for d in self.drones]

    def _actionSpace(self):
        """Per-drone Box([-1,-1,-1,0], [1,1,1,1]) (VelocityAviary.py:98-117)."""
        return {str(i): _Box(np.array([-1, -1, -1, 0]), np.array([1, 1, 1, 1])) for i in range(self.NUM_DRONES)}

    def _advance(self, action):
        a = self._pack_action(action, width=4)
        self._core.step(self._core.targets_velocity(a), 1, order=L.DS_ORDER_CONTROL_THEN_PHYSICS)
