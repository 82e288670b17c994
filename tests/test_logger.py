"""The rollout log format: ``dronesim_b200.utils.Logger`` against the reference's own ``Logger``
executed here (behind stand-ins for its plotting imports), and the on-device capture against per-step observations."""
import os
import sys
import types

import numpy as np
import pytest

from dronesim_b200.utils import Logger
from oracle import ref_shims


def _reference_logger_class():
    """dronesim/utils/Logger.py imports matplotlib / cycler for its plot() method only (Logger.py:5-7)."""
    ref_shims.install()
    for name in ("matplotlib", "matplotlib.pyplot", "cycler"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.cycler = lambda *a, **k: None
                sys.modules[name] = m
    if "matplotlib.pyplot" in sys.modules and "matplotlib" in sys.modules:
        setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])
    from dronesim.utils.Logger import Logger as RefLogger

    return RefLogger


def _drive(lg, rng, n_drones, steps, state_len):
    for t in range(steps):
        for j in range(n_drones):
            lg.log(drone=j, timestamp=t / 48.0, state=rng.normal(size=state_len), control=rng.normal(size=12))


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference checkout not present")
@pytest.mark.parametrize("duration", [0, 1])
def test_logger_matches_reference_logger(duration, tmp_path):
    Ref = _reference_logger_class()
    a = Logger(logging_freq_hz=48, state_length=20, num_drones=2, duration_sec=duration)
    b = Ref(logging_freq_hz=48, state_length=20, num_drones=2, duration_sec=duration)
    _drive(a, np.random.default_rng(4), 2, 60, 20)  # 60 samples: overflows the 48 preallocated columns
    _drive(b, np.random.default_rng(4), 2, 60, 20)
    for k in ("timestamps", "states", "controls", "counters"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k))
    pa = a.save(file_path=str(tmp_path) + os.sep, file_name="ours")
    b.save(file_path=str(tmp_path) + os.sep, file_name="ref")
    za, zb = np.load(pa), np.load(os.path.join(str(tmp_path), "ref.npy"))
    assert sorted(za.files) == sorted(zb.files) == ["controls", "states", "timestamps"]
    for k in za.files:
        np.testing.assert_array_equal(za[k], zb[k])


def test_logger_shapes_and_growth():
    lg = Logger(logging_freq_hz=10, state_length=22, num_drones=3, duration_sec=2)
    assert lg.states.shape == (3, 22, 20) and lg.controls.shape == (3, 12, 20) and lg.timestamps.shape == (3, 20)
    _drive(lg, np.random.default_rng(0), 3, 25, 22)
    assert lg.states.shape == (3, 22, 25) and (lg.counters == 25).all()


@pytest.mark.gpu
def test_device_log_equals_per_step_observations():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dronesim_b200.envs import CtrlAviary, Physics

    init = np.array([[0.0, 0.0, 1.0], [0.4, 0.0, 1.5]])
    env = CtrlAviary(drone_model=["robobee", "hexa_6DOF"], num_drones=2, initial_xyzs=init, physics=Physics.PYB_GND_DRAG_DW,
                     aggregate_phy_steps=5, num_envs=4)
    lg = Logger(logging_freq_hz=48, state_length=22, num_drones=2, duration_sec=1)
    lg.attach(env, env_index=2)
    env.reset()
    rng = np.random.default_rng(8)
    seen = []
    for t in range(30):
        a = np.zeros((4, 2, 6), dtype=np.float32)
        a[:, 0, :4] = 0.45 + rng.uniform(-0.05, 0.05, (4, 4))
        a[:, 1, :] = 0.40 + rng.uniform(-0.05, 0.05, (4, 6))
        obs, _, _, _ = env.step(a)
        seen.append(obs["state"][2].cpu().numpy().astype(np.float64))  # [2, 22] of env 2
    T = lg.collect()
    assert T == 30 and lg.states.shape == (2, 22, 30)
    # same device function in two kernels (ds_log_kernel / ds_obs_kernel): equal up to FMA-contraction choices
    np.testing.assert_allclose(lg.states, np.stack(seen, axis=2), rtol=5e-7, atol=2e-7)
    np.testing.assert_array_equal(lg.states[:, 0:7], np.stack(seen, axis=2)[:, 0:7])  # copied fields are bit-exact
    np.testing.assert_allclose(lg.timestamps[0], (np.arange(30) + 1) * 5 / 240.0)
    # a second reset rewinds the device log
    env.reset()
    assert lg.collect() == 0
    env.close()
