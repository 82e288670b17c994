"""The vectorised oracle (oracle/batch.py) against the per-vehicle oracle (oracle/sim.py) it restates: CPU only."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from dronesim_b200.vehicles import load_vehicle  # noqa: E402
from dronesim_b200.workloads import hetero16  # noqa: E402
from oracle.batch import BatchOracle  # noqa: E402
from oracle.sim import OracleSwarm  # noqa: E402


def _run(models, E, K, flags, pos0, act0, tpos, tyaw, steps):
    vts = [load_vehicle(m) for m in models]
    a = OracleSwarm(vts, E, integrator="quat", composite=True, gnd=flags["ground"], drag=flags["drag"], dw=flags["downwash"],
                    aggregate_phy_steps=K)
    b = BatchOracle(vts, E, gnd=flags["ground"], drag=flags["drag"], dw=flags["downwash"], aggregate_phy_steps=K)
    a.reset(pos0)
    b.reset(pos0)
    act_a, act_b = act0.copy(), act0.copy()
    for _ in range(steps):
        a.physics_step(act_a)
        act_a = a.control_step(tpos, tyaw=tyaw)
        b.physics_step(act_b)
        act_b = b.control_step(tpos, tyaw=tyaw)
    return a, b, act_a, act_b


def test_batch_oracle_equals_per_vehicle_oracle_hetero16():
    E = 2
    models, K, flags, pos0, act0, tgt = hetero16(E, seed=3)
    tpos = tgt[:, :3].reshape(E, 16, 3)
    a, b, act_a, act_b = _run(models, E, K, flags, pos0, act0, tpos, None, steps=20)
    np.testing.assert_allclose(b.pos, a.pos, atol=1e-10)
    np.testing.assert_allclose(b.quat, a.quat, atol=1e-10)
    np.testing.assert_allclose(b.vel, a.vel, atol=1e-9)
    np.testing.assert_allclose(b.rates, a.rates, atol=1e-8)
    np.testing.assert_allclose(act_b, act_a, atol=1e-9)


@pytest.mark.parametrize("model", ["robobee", "hexa_6DOF", "hexa_6DOF_simple"])
def test_batch_oracle_single_types_with_yaw_targets_and_slow_wls(model):
    E = 6
    rng = np.random.default_rng(5)
    pos0 = np.array([0.0, 0.0, 1.0]) + rng.uniform(-0.05, 0.05, (E, 1, 3))
    act0 = np.zeros((E, 1, 6))
    act0[:, 0, : (6 if "hexa" in model else 4)] = 0.4
    tpos = pos0 + rng.uniform(-1.5, 1.5, (E, 1, 3))  # far set-points: saturating commands, WLS active set on the 6-DOF law
    tyaw = rng.uniform(-3.0, 3.0, (E, 1))
    flags = dict(ground=True, drag=True, downwash=False)
    a, b, act_a, act_b = _run([model], E, 4, flags, pos0, act0, tpos, tyaw, steps=30)
    np.testing.assert_allclose(b.pos, a.pos, atol=1e-9)
    np.testing.assert_allclose(b.quat, a.quat, atol=1e-9)
    np.testing.assert_allclose(act_b, act_a, atol=1e-8)
