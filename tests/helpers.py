"""Shared helpers for the parity tests: drive the CUDA core and the FP64 oracle side by side."""
import numpy as np

from oracle import pyb_math
from oracle.sim import OracleSwarm


def rpy_of_quat(q):
    return np.array([pyb_math.getEulerFromQuaternion(x) for x in np.asarray(q, float).reshape(-1, 4)])


def angle_between(q1, q2):
    """Rotation angle (rad) between two xyzw unit quaternions, elementwise."""
    q1 = np.asarray(q1, float).reshape(-1, 4)
    q2 = np.asarray(q2, float).reshape(-1, 4)
    d = np.abs(np.sum(q1 * q2, axis=1) / (np.linalg.norm(q1, axis=1) * np.linalg.norm(q2, axis=1)))
    return 2.0 * np.arccos(np.clip(d, -1.0, 1.0))


def make_pair(models, n_envs, integrator, K, gnd=False, drag=False, dw=False, radius=np.inf, motor_tau=0.0, acc_filter_hz=0.0,
              **kw):
    """(SwarmCore, OracleSwarm) with identical configuration."""
    from dronesim_b200.core import SwarmCore
    from dronesim_b200.vehicles import load_vehicle

    vts = [load_vehicle(m) for m in models]
    composite = integrator == "quat"
    core = SwarmCore(vts, n_envs, integrator=integrator, composite=composite, ground=gnd, drag=drag, downwash=dw,
                     aggregate_phy_steps=K, neighbourhood_radius=radius, motor_tau=motor_tau, acc_filter_hz=acc_filter_hz, **kw)
    orc = OracleSwarm(vts, n_envs, integrator=integrator, composite=composite, gnd=gnd, drag=drag, dw=dw,
                      aggregate_phy_steps=K, neighbourhood_radius=radius, motor_tau=motor_tau, acc_filter_hz=acc_filter_hz)
    return core, orc


def core_state(core):
    v = core.views()
    return {k: (v[k].detach().cpu().numpy().astype(np.float64) if hasattr(v[k], "cpu") else v[k]) for k in v}


def adjacency_bits_f32(pos, radius):
    """The adjacency bitmask exactly as ``ds_obs_kernel`` evaluates it: ``|p_i - p_j|^2 < r^2`` in float32, every operation
    correctly rounded, in the order ((dx*dx + dy*dy) + dz*dz) against fl(r * r).  ``pos`` [E, D, 3] (float32-representable)."""
    p = np.asarray(pos).astype(np.float32)
    E, D, _ = p.shape
    d = p[:, :, None, :] - p[:, None, :, :]                      # [E, i, j, 3], float32
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    r2 = np.float32(radius) * np.float32(radius)
    near = d2 < r2
    near[:, np.arange(D), np.arange(D)] = True
    return (near.astype(np.uint64) << np.arange(D, dtype=np.uint64)[None, None, :]).sum(axis=2).astype(np.uint32)
