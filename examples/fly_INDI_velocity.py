"""The reference's ``examples/fly_INDI_velocity.py`` loop (:122-215) on the B200 core: only the imports differ.

    python examples/fly_INDI_velocity.py [--drone tello] [--num_drones 5] [--num_envs 1]

Five quadrotors on a circle each receive the constant velocity command of the reference script - direction (0.2, 0.2,
0.2), 2 % of the airframe's speed limit (:177-182) - through ``VelocityAviary`` (whose ``_preprocessAction`` runs the
INDI law, VelocityAviary.py:221-264; here fused with the physics in one launch).  GUI / video are not available.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dronesim_b200.envs.BaseAviary import Physics  # noqa: E402           (reference: dronesim.envs.BaseAviary)
from dronesim_b200.envs.VelocityAviary import VelocityAviary  # noqa: E402  (reference: dronesim.envs.VelocityAviary)


def main(argv=None):
    ap = argparse.ArgumentParser(description="velocity-command flight (fly_INDI_velocity.py of the reference)")
    ap.add_argument("--drone", default="tello")
    ap.add_argument("--num_drones", default=5, type=int)
    ap.add_argument("--simulation_freq_hz", default=240, type=int)
    ap.add_argument("--control_freq_hz", default=96, type=int)
    ap.add_argument("--duration_sec", default=20, type=int)
    ap.add_argument("--num_envs", default=1, type=int)
    ARGS = ap.parse_args(argv)
    E, ND = ARGS.num_envs, ARGS.num_drones

    H, H_STEP, R = 0.50, 0.05, 1.5                                        # fly_INDI_velocity.py:122-125
    AGGR_PHY_STEPS = int(ARGS.simulation_freq_hz / ARGS.control_freq_hz)  # :127-129
    INIT_XYZS = np.array([[R * np.cos((i / 6) * 2 * np.pi + np.pi / 2), R * np.sin((i / 6) * 2 * np.pi + np.pi / 2),
                           H + i * H_STEP] for i in range(ND)])           # :130-139
    INIT_RPYS = np.zeros((ND, 3))
    env = VelocityAviary(drone_model=ND * [ARGS.drone], num_drones=ND, initial_xyzs=INIT_XYZS, initial_rpys=INIT_RPYS,
                         physics=Physics.PYB, neighbourhood_radius=10, freq=ARGS.simulation_freq_hz,
                         aggregate_phy_steps=AGGR_PHY_STEPS, num_envs=E, ground_plane=True)   # :143-156
    CTRL_EVERY_N_STEPS = int(np.floor(env.SIM_FREQ / ARGS.control_freq_hz))   # :171
    # the reference's first action is a 4-vector of 0.4 (:172), which VelocityAviary reads as a velocity command too
    action = {str(i): np.array([0.4, 0.4, 0.4, 0.4]) for i in range(ND)}
    obs = env.reset()
    START = time.time()
    for i in range(0, int(ARGS.duration_sec * env.SIM_FREQ), AGGR_PHY_STEPS):   # :174
        obs, reward, done, info = env.step(action if E == 1 else np.tile(np.stack([action[str(j)] for j in range(ND)]), (E, 1, 1)))
        if i % CTRL_EVERY_N_STEPS == 0:                                   # :180-194
            for j in range(ND):
                V_des_unit = np.ones(3) * 0.2
                magnitude = 0.02
                action[str(j)] = np.array([V_des_unit[0], V_des_unit[1], V_des_unit[2], magnitude])
        if i % env.SIM_FREQ == 0:
            env.render()
    st = np.stack([obs[str(j)]["state"] for j in range(ND)]) if E == 1 else np.asarray(obs["state"][0].cpu())
    speed_limit = env.SPEED_LIMIT[0]
    v_cmd = speed_limit * 0.02 * np.ones(3) / np.sqrt(3.0)
    out = {"script": "fly_INDI_velocity", "drone": ARGS.drone, "num_drones": ND, "num_envs": E,
           "commanded_velocity": [float(x) for x in v_cmd],
           "final_velocity_mean": [float(x) for x in st[:, 10:13].mean(axis=0)],
           "final_velocity_error_max": float(np.abs(st[:, 10:13] - v_cmd).max()),
           "displacement_mean": [float(x) for x in (st[:, 0:3] - INIT_XYZS).mean(axis=0)], "wall_clock_s": time.time() - START}
    print("[INFO] %s" % out)
    env.close()
    return out


if __name__ == "__main__":
    main()
