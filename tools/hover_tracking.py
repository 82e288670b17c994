"""Hover-settle and tracking-error record of the example scripts on the GPU core, for offline comparison with the
reference's PyBullet runs (`python examples/<script>.py --physics pyb` in an environment that has pybullet):

    python tools/hover_tracking.py [--out profiles/hover_tracking_r02.json]      (needs a GPU)

BASELINE configs[0..2] + the velocity example, each flown with the flags the reference script defaults to and with the
ground plane of its PyBullet world; quantities: settle time and steady-state error of the hover (fly_INDI.py), time to
the final gate and RMS tracking error (fly_INDI_TrajectoryTrack.py), RMS circle-tracking error and level-flight bound
(fly_hexa_6DOF.py), velocity-command error (fly_INDI_velocity.py).  SURVEY.md 8(c): the PyBullet path itself cannot run
in this image, so these numbers are what a PyBullet run is to be held against."""
import argparse
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "examples", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def hover_record(physics):
    """fly_INDI.py (robobee, target [0, 0, 0.5] from [0, 1, 0.5], 48 Hz control), 10 s: settle time = first time after
    which |pos - target| stays below 5 cm; steady-state error = mean |pos - target| over the last second."""
    from dronesim_b200.control.INDIControl import INDIControl
    from dronesim_b200.envs.BaseAviary import Physics
    from dronesim_b200.envs.CtrlAviary import CtrlAviary
    from dronesim_b200.utils.Logger import Logger

    SIM, CTRL, DUR = 240, 48, 10
    AGGR = SIM // CTRL
    env = CtrlAviary(drone_model=["robobee"], num_drones=1, initial_xyzs=np.array([[0.0, 1.0, 0.5]]), physics=Physics(physics),
                     neighbourhood_radius=10, freq=SIM, aggregate_phy_steps=AGGR, ground_plane=(physics != "dyn"))
    ctrl = INDIControl(drone_model="robobee")
    logger = Logger(logging_freq_hz=CTRL, num_drones=1, duration_sec=DUR)
    logger.attach(env)
    NUM_WP = CTRL * 15
    yaw = [0.4 + i / 200 for i in range(NUM_WP)]
    action = {"0": np.array([0.4, 0.4, 0.4, 0.4])}
    env.reset()
    wp = 0
    for i in range(0, DUR * SIM, AGGR):
        obs, _, _, _ = env.step(action)
        action["0"], _, _ = ctrl.computeControlFromState(control_timestep=AGGR / SIM, state=obs["0"]["state"],
                                                         target_pos=np.array([0.0, 0.0, 0.5]), target_rpy=np.array([0, 0, yaw[wp]]))
        wp = wp + 1 if wp < NUM_WP - 1 else 0
    T = logger.collect()
    err = np.linalg.norm(logger.states[0, 0:3, :T] - np.array([[0.0], [0.0], [0.5]]), axis=0)
    t = logger.timestamps[0, :T]
    outside = np.flatnonzero(err > 0.05)
    settle = float(t[outside[-1] + 1]) if outside.size and outside[-1] + 1 < T else (0.0 if not outside.size else None)
    env.close()
    ctrl.close()
    return {"script": "fly_INDI", "physics": physics, "duration_s": DUR, "settle_time_s_5cm": settle,
            "steady_state_error_m_last_second": float(err[t >= DUR - 1].mean()), "max_error_m": float(err.max()),
            "min_altitude_m": float(logger.states[0, 2, :T].min())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "hover_tracking_r02.json"))
    a = ap.parse_args()
    rec = {"note": "GPU core (FP32), ground plane on for the PyBullet physics names; compare with the reference's "
                   "`--physics pyb` runs of the same scripts (PyBullet is absent from this image)",
           "runs": []}
    for ph in ("dyn", "pyb"):
        rec["runs"].append(hover_record(ph))
    rec["runs"].append(load("fly_INDI_TrajectoryTrack").main(["--physics", "pyb"]))
    rec["runs"].append(load("fly_INDI_TrajectoryTrack").main(["--physics", "pyb", "--num_envs", "4096"]))
    rec["runs"].append(load("fly_hexa_6DOF").main(["--physics", "pyb_gnd_drag_dw"]))
    rec["runs"].append(load("fly_INDI_velocity").main([]))
    with open(a.out, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
