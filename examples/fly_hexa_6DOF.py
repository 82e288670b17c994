"""The reference's ``examples/fly_hexa_6DOF.py`` loop (:134-262) on the B200 core: only the imports differ.

    python examples/fly_hexa_6DOF.py [--drone hexa_6DOF] [--physics pyb_gnd_drag_dw] [--num_envs 65536]

The tilted-rotor hexarotor holds its attitude level with the 6-DOF INDI law + WLS allocation while its set-point runs
twice around a circle of radius R (:156-166); BASELINE.json configs[2] is this script at 65k envs with ground effect
and drag.  GUI, video, camera following and plotting of the reference script are not available.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dronesim_b200.control.INDIControl_6DOF import INDIControl  # noqa: E402  (reference: dronesim.control.INDIControl_6DOF)
from dronesim_b200.envs.BaseAviary import Physics  # noqa: E402
from dronesim_b200.envs.CtrlAviary import CtrlAviary  # noqa: E402
from dronesim_b200.utils.Logger import Logger  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description="hexarotor 6-DOF INDI control (fly_hexa_6DOF.py of the reference)")
    ap.add_argument("--drone", default="hexa_6DOF")
    ap.add_argument("--physics", default="pyb_gnd_drag_dw", choices=[p.value for p in Physics])
    ap.add_argument("--simulation_freq_hz", default=240, type=int)
    ap.add_argument("--control_freq_hz", default=96, type=int)
    ap.add_argument("--duration_sec", default=10, type=int)
    ap.add_argument("--num_envs", default=1, type=int)
    ARGS = ap.parse_args(argv)
    E = ARGS.num_envs

    R = 1.2                                                               # fly_hexa_6DOF.py:134-136
    AGGR_PHY_STEPS = int(ARGS.simulation_freq_hz / ARGS.control_freq_hz)  # :142-144
    INIT_XYZS = np.array([[0.0, 0.0, 0.6]])                               # :151-154
    INIT_RPYS = np.array([[0.0, 0.0, 0.0]])
    INIT_VELS = np.array([[0.0, 0.0, 0.0]])
    PERIOD = 15                                                           # :157-166
    NUM_WP = ARGS.control_freq_hz * PERIOD
    TARGET_POS = np.zeros((NUM_WP, 3))
    for i in range(NUM_WP):
        TARGET_POS[i, :] = (R * np.cos((i / NUM_WP) * (4 * np.pi) + np.pi / 2) + INIT_XYZS[0, 0],
                            R * np.sin((i / NUM_WP) * (4 * np.pi) + np.pi / 2) - R + INIT_XYZS[0, 1], 0)
    wp_counter = 0
    TARGET_RPYS = np.zeros((NUM_WP, 3))                                   # :171-173
    env = CtrlAviary(drone_model=[ARGS.drone], num_drones=1, initial_xyzs=INIT_XYZS, initial_vels=INIT_VELS,
                     initial_rpys=INIT_RPYS, physics=Physics(ARGS.physics), neighbourhood_radius=10,
                     freq=ARGS.simulation_freq_hz, aggregate_phy_steps=AGGR_PHY_STEPS, num_envs=E,
                     ground_plane=(ARGS.physics != "dyn"))  # the PyBullet world has a floor (BaseAviary.py:679-680)
    logger = Logger(logging_freq_hz=int(ARGS.simulation_freq_hz / AGGR_PHY_STEPS), num_drones=1, duration_sec=ARGS.duration_sec)
    logger.attach(env)
    ctrl = [INDIControl(drone_model=ARGS.drone, num_envs=E)]              # :201
    CTRL_EVERY_N_STEPS = int(np.floor(env.SIM_FREQ / ARGS.control_freq_hz))   # :204
    action = {"0": np.array([0.1, 0.1, 0.1, 0.1, 0.1, 0.1])}              # :205-207
    env.reset()
    START = time.time()
    err2, n_err = 0.0, 0
    for i in range(0, int(ARGS.duration_sec * env.SIM_FREQ), AGGR_PHY_STEPS):   # :210
        obs, reward, done, info = env.step(action)                        # :213
        if i % CTRL_EVERY_N_STEPS == 0:                                   # :216
            state = obs["0"]["state"] if E == 1 else obs["state"][:, 0, :]
            cmd, pos_e, _ = ctrl[0].computeControlFromState(
                control_timestep=CTRL_EVERY_N_STEPS * env.TIMESTEP, state=state,
                target_pos=np.hstack([TARGET_POS[wp_counter, 0:2], INIT_XYZS[0, 2]]), target_rpy=TARGET_RPYS[wp_counter])
            action = {"0": cmd} if E == 1 else cmd.reshape(E, 1, -1)
            pe = np.asarray(pos_e.cpu() if hasattr(pos_e, "cpu") else pos_e, dtype=float).reshape(-1, 3)
            if i >= 2 * env.SIM_FREQ:  # tracking error after the 2 s take-off transient
                err2 += float((pe ** 2).sum(axis=1).mean())
                n_err += 1
            wp_counter = wp_counter + 1 if wp_counter < (NUM_WP - 1) else 0   # :233-236
        if i % env.SIM_FREQ == 0:
            env.render()
    T = logger.collect()
    rp = np.abs(logger.states[0, 7:9, :T])  # roll, pitch: the 6-DOF law flies level
    out = {"script": "fly_hexa_6DOF", "drone": ARGS.drone, "physics": ARGS.physics, "num_envs": E,
           "rms_tracking_error_m_after_2s": float(np.sqrt(err2 / max(n_err, 1))),
           "max_abs_roll_pitch_rad_after_2s": float(rp[:, min(T - 1, 2 * ARGS.control_freq_hz):].max()) if T > 2 * ARGS.control_freq_hz else None,
           "final_position": [float(x) for x in logger.states[0, 0:3, T - 1]], "samples": int(T),
           "wall_clock_s": time.time() - START}
    print("[INFO] %s" % out)
    env.close()
    ctrl[0].close()
    return out


if __name__ == "__main__":
    main()
