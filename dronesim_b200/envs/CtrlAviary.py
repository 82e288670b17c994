"""``CtrlAviary`` facade (dronesim/envs/CtrlAviary.py): per-drone PWM actions, state + neighbours obs.

* action: ``{str(i): ndarray[n_u_i]}`` PWM, clipped to ``[MIN_PWM, MAX_PWM]`` per drone
  (CtrlAviary.py:236-263) - the clip happens inside the CUDA step;
* obs: ``{str(i): {"state": (16 + n_u_i,), "neighbors": (NUM_DRONES,)}}`` (CtrlAviary.py:212-232);
* reward -1, done False, info ``{"answer": 42}`` (CtrlAviary.py:267-310).

With ``num_envs > 1`` the same call returns batched device tensors instead of the per-drone dict:
``{"state": [E, D, 22], "neighbors": [E, D] bitmask}``, ``reward [E]``, ``done [E]``.
"""
from __future__ import annotations

import numpy as np

from .BaseAviary import BaseAviary, Physics  # noqa: F401


class _Box:
    """Minimal stand-in for ``gym.spaces.Box`` (gym is not a dependency of the core)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low, self.high, self.dtype = np.asarray(low, dtype), np.asarray(high, dtype), dtype
        self.shape = self.low.shape

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool((x >= self.low).all() and (x <= self.high).all())

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class CtrlAviary(BaseAviary):
    """Multi-drone environment class for control applications."""

    def _actionSpace(self):
        """Per-drone Box(MIN_PWM, MAX_PWM) (CtrlAviary.py:95-120)."""
        return {str(i): _Box(np.array(d.MIN_PWM), np.array(d.MAX_PWM)) for i, d in enumerate(self.drones)}

    def _observationSpace(self):
        """Per-drone {"state": Box(16 + n_u), "neighbors": MultiBinary(NUM_DRONES)} (CtrlAviary.py:124-208)."""
        out = {}
        for i, d in enumerate(self.drones):
            n = 16 + d.INDI_ACTUATOR_NR
            lo = np.full(n, -np.inf)
            hi = np.full(n, np.inf)
            lo[2] = 0.0
            lo[3:7], hi[3:7] = -1.0, 1.0
            lo[7:10], hi[7:10] = -np.pi, np.pi
            lo[16:], hi[16:] = np.array(d.MIN_PWM), np.array(d.MAX_PWM)
            out[str(i)] = {"state": _Box(lo, hi), "neighbors": _Box(np.zeros(self.NUM_DRONES), np.ones(self.NUM_DRONES), np.int8)}
        return out

    def _computeObs(self):
        if self.NUM_ENVS > 1:
            return {"state": self._state_dev, "neighbors": self._neigh_dev}
        adjacency_mat = self._getAdjacencyMatrix()
        return {str(i): {"state": self._getDroneStateVector(i), "neighbors": adjacency_mat[i, :]}
                for i in range(self.NUM_DRONES)}

    def _computeReward(self):
        return -1 if self.NUM_ENVS == 1 else self._reward_dev

    def _computeDone(self):
        if self.NUM_ENVS == 1:
            return bool(self._done_dev.item()) if self._core_has_done() else False
        return self._done_dev.bool()

    def _core_has_done(self):
        return True

    def _computeInfo(self):
        return {"answer": 42}
