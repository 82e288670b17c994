"""Synthetic swarms of the BASELINE.json configurations (SURVEY.md section 8 d), seeded numpy only.

Shared by ``bench.py`` (both arms), the tests and ``__graft_entry__.smoke()`` so that the CUDA core
and the CPU oracle always see the same inputs.  Nothing here touches the GPU.

``hetero16`` is configs[3]/[4] of BASELINE.json: 16 drones per env, 8 quads (robobee / tello
alternating) + 8 hexa_6DOF, slot -> type fixed per env, ground effect + drag + downwash, K = 8
substeps per control step, hover at the initial position.

Layout: 4 x 4 lateral grid with 1.0 m pitch, altitude 2.0 + 0.25 * slot.  The reference's downwash
model (BaseAviary.py:1753-1755) has alpha ~ 1 / dz^2, i.e. it is singular when a vehicle passes
through the altitude of a laterally close neighbour: a 1e-16 m rounding difference between two
"same altitude" vehicles 0.5 m apart produces a 1e26 N force in the FP64 oracle, and on the GPU a
0.5 m grid lost ~0.1 % of 4 Mi vehicles to such crossings during the start-up transient (the quad
law restarts from cmd = 0, INDIControl.py:129, and drops ~1.2 m before it recovers, robobee and
tello by different amounts).  SURVEY's "0.5 m pitch, z = 1.0 + 0.25 (slot mod 4)" layout therefore
cannot be flown by the reference formula at all.  At 1.0 m lateral pitch the Gaussian factor of a
crossing pair is exp(-0.5 (1.0 / 0.11)^2) ~ 1e-18, which keeps the singularity harmless; the
downwash arithmetic executed per pair is the same whatever the geometry (no early-out).
"""
from __future__ import annotations

import numpy as np

HETERO16_MODELS = ["robobee", "tello"] * 4 + ["hexa_6DOF"] * 8


def hetero16(n_envs: int, seed: int = 0, env_offset: int = 0, dtype=np.float64):
    """(models, K, flags, pos0[E,16,3], action0[E,16,6], targets[E*16,4]) of the heterogeneous swarm.

    The per-env random offsets are a pure function of (seed, global env index), so any sharding of
    the envs over ranks reproduces the same swarm.  ``dtype=np.float32`` halves the host memory of the 64 Mi-vehicle
    sweep point (the core converts to float32 anyway)."""
    D = 16
    slot = np.arange(D)
    base = np.stack([1.0 * (slot % 4), 1.0 * (slot // 4), 2.0 + 0.25 * slot], axis=1)  # [16,3]
    # counter-based noise: hash of the global (env, slot, axis) index -> U(-0.02, 0.02)
    e = (np.arange(n_envs, dtype=np.uint64) + np.uint64(env_offset))[:, None, None]
    idx = (e * np.uint64(D) + slot.astype(np.uint64)[None, :, None]) * np.uint64(3) + np.arange(3, dtype=np.uint64)[None, None, :]
    noise = (_hash01(idx, seed) - 0.5) * 0.04
    pos0 = (base[None, :, :] + noise).astype(dtype)
    del noise, idx
    action0 = np.zeros((n_envs, D, 6), dtype=dtype)
    action0[:, :8, :4] = 0.45
    action0[:, 8:, :] = 0.45
    tgt = np.concatenate([pos0.reshape(-1, 3), np.zeros((n_envs * D, 1), dtype=dtype)], axis=1)
    flags = dict(ground=True, drag=True, downwash=True)
    return HETERO16_MODELS, 8, flags, pos0, action0, tgt


def _hash01(idx: np.ndarray, seed: int) -> np.ndarray:
    """splitmix64 of (idx, seed) -> float64 in [0, 1)."""
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(seed + 1)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


# algorithmic bytes per vehicle per control step (SURVEY.md section 8 d, FP32 SoA, no padding)
def algorithmic_bytes_per_control_step(n_u: int, per_vehicle_targets: bool = True) -> int:
    phys = 2 * 4 * 13  # pos3 quat4 vel3 omega3, read + write
    ctrl = 2 * 4 * (3 + 3 + 1 + n_u)  # last_vel3 last_rates3 last_thrust cmd[n_u], read + write
    tgt = 4 * 10 if per_vehicle_targets else 8  # pos3 vel3 acc3 yaw | waypoint counter r+w
    return phys + ctrl + tgt + 1  # + type id


def hetero16_bytes_per_control_step() -> float:
    """Mean over the 8 quads + 8 hexas of an env: (233 + 249) / 2 = 241 B."""
    return 0.5 * (algorithmic_bytes_per_control_step(4) + algorithmic_bytes_per_control_step(6))


# --------------------------------------------------------------------------------------------
# single-type swarms of BASELINE configs[1] / configs[2] at scale (one drone per env)
# --------------------------------------------------------------------------------------------
def circle_table(num_wp: int = 1440, radius: float = 1.2, z: float = 0.6) -> np.ndarray:
    """examples/fly_hexa_6DOF.py:157-166,224-226: circle of ``radius`` flown twice over NUM_WP waypoints,
    rows = pos3 vel3 acc3 yaw (vel / acc / yaw zero, as the script passes none)."""
    i = np.arange(num_wp)
    tab = np.zeros((num_wp, 10))
    tab[:, 0] = radius * np.cos((i / num_wp) * (4 * np.pi) + np.pi / 2)
    tab[:, 1] = radius * np.sin((i / num_wp) * (4 * np.pi) + np.pi / 2) - radius
    tab[:, 2] = z
    return tab


def reference_trajectory_table() -> np.ndarray:
    """[1200, 10] = pos3 vel3 acc3 yaw: the table the reference's ``trajGenerator`` returns for the three gates of
    fly_INDI_TrajectoryTrack.py:133-160, generated by executing the reference (tests/golden/make_golden.py::traj_fixture)."""
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "traj_3gates.npz")
    return np.array(np.load(path)["table"], dtype=np.float64)


def single_type(name: str, n_envs: int, seed: int = 0, env_offset: int = 0):
    """(models, K, flags, pos0[E,1,3], action0[E,1,6], table[num_wp,10], wp0[E]) of

    ``"traj_quad"``  configs[1]: robobee tracking the reference's own 3-gate trajectory (the 1200-row table its
                     trajGenerator produces for fly_INDI_TrajectoryTrack.py:133-160, frozen in
                     tests/golden/traj_3gates.npz), K = 2 (96 Hz control), no add-ons;
    ``"hexa_circle"`` configs[2]: hexa_6DOF on the fly_hexa_6DOF.py circle, K = 2, ground effect + drag;
    ``"quad_k8"``    robobee hover-table, K = 8, no add-ons (the plain dynamics + INDI path of configs[4]).
    """
    e = (np.arange(n_envs, dtype=np.uint64) + np.uint64(env_offset))[:, None]
    idx = e * np.uint64(3) + np.arange(3, dtype=np.uint64)[None, :]
    noise = (_hash01(idx, seed) - 0.5) * 0.1  # U(-0.05, 0.05)
    if name == "traj_quad":
        models, K, flags = ["robobee"], 2, dict(ground=False, drag=False, downwash=False)
        tab = reference_trajectory_table()
        base, cmd0 = np.array([-3.0, 0.0, 2.0]), 0.4
    elif name == "hexa_circle":
        models, K, flags = ["hexa_6DOF"], 2, dict(ground=True, drag=True, downwash=False)
        tab = circle_table()
        base, cmd0 = np.array([0.0, 0.0, 0.6]), 0.1
    elif name == "quad_k8":
        models, K, flags = ["robobee"], 8, dict(ground=False, drag=False, downwash=False)
        tab = np.zeros((720, 10))
        tab[:, 2] = 0.5
        tab[:, 9] = 0.4 + np.arange(720) / 200.0  # fly_INDI.py:165-167
        base, cmd0 = np.array([0.0, 0.0, 0.5]), 0.4
    else:
        raise KeyError(name)
    # every env starts ON the trajectory at its own waypoint (env index staggers the table rows read per step)
    wp0 = ((np.arange(n_envs, dtype=np.int64) + env_offset) % tab.shape[0]).astype(np.int32)
    del base
    pos0 = (tab[wp0, 0:3] + noise).reshape(n_envs, 1, 3)
    act0 = np.zeros((n_envs, 1, 6))
    act0[:, 0, : (6 if "hexa" in models[0] else 4)] = cmd0
    return models, K, flags, pos0, act0, tab, wp0
