"""``RPYTAviary`` facade (dronesim/envs/RPYTAviary.py): body-rate + thrust actions.

* action: ``{str(i): [p, q, r set-point, thrust]}`` (the order ``_preprocessAction`` reads:
  ``target_rpy_rates=v[:3]``, ``thrust=v[3]``, RPYTAviary.py:180-193);
* ``_preprocessAction``: per drone ``INDIControl._INDIRateControl(control_timestep=AGGR_PHY_STEPS*TIMESTEP,
  thrust, cur_quat, cur_ang_vel, target_rpy_rates)`` (INDIControl.py:413-490); the PWM command it returns
  is applied for the AGGR_PHY_STEPS substeps (BaseAviary.py:507-545);
* obs / reward / done / info as CtrlAviary.

One fused kernel launch per ``step`` (``ds_step`` with target mode 3, order control-then-physics).
Quad-law airframes only: ``INDIControl_6DOF`` has no rate/thrust entry (the core returns
``DS_ERR_UNSUPPORTED`` for a 6-DOF type).
"""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from .BaseAviary import Physics  # noqa: F401
from .CtrlAviary import CtrlAviary, _Box


class RPYTAviary(CtrlAviary):
    """Multi-drone environment class for rate / thrust control."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.SPEED_LIMIT = [d.MAX_SPEED_KMH * (1000 / 3600) for d in self.drones]  # RPYTAviary.py:79-81

    def _actionSpace(self):
        """Per-drone Box([0,-1,-1,-1], [THRUST2WEIGHT_RATIO,1,1,1]) (RPYTAviary.py:96-99)."""
        return {str(i): _Box(np.array([0.0, -1, -1, -1]), np.array([d.THRUST2WEIGHT_RATIO, 1, 1, 1]))
                for i, d in enumerate(self.drones)}

    def _advance(self, action):
        a = self._pack_action(action, width=4)
        self._core.step(self._core.targets_rate_thrust(a), 1, order=L.DS_ORDER_CONTROL_THEN_PHYSICS)
