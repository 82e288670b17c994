"""The reference's ``examples/fly_INDI.py`` loop (:139-262) on the B200 core: only the imports differ.

    python examples/fly_INDI.py [--drone robobee] [--physics dyn] [--duration_sec 2] [--num_envs 1]

With ``--num_envs E`` the same script flies E independent copies of the aviary in one launch per
``env.step`` / ``ctrl.computeControlFromState`` (observations and commands are device tensors then).
GUI, video and plotting flags of the reference script are not available (host-side visualisation is out of scope).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dronesim_b200.control.INDIControl import INDIControl  # noqa: E402  (reference: dronesim.control.INDIControl)
from dronesim_b200.envs.BaseAviary import Physics  # noqa: E402          (reference: dronesim.envs.BaseAviary)
from dronesim_b200.envs.CtrlAviary import CtrlAviary  # noqa: E402       (reference: dronesim.envs.CtrlAviary)
from dronesim_b200.utils.Logger import Logger  # noqa: E402             (reference: dronesim.utils.Logger)


def main(argv=None):
    ap = argparse.ArgumentParser(description="INDI hover / yaw sweep (fly_INDI.py of the reference)")
    ap.add_argument("--drone", default="robobee")
    ap.add_argument("--physics", default="dyn", choices=[p.value for p in Physics])
    ap.add_argument("--simulation_freq_hz", default=240, type=int)
    ap.add_argument("--control_freq_hz", default=48, type=int)
    ap.add_argument("--duration_sec", default=2, type=int)
    ap.add_argument("--num_envs", default=1, type=int)
    ARGS = ap.parse_args(argv)

    AGGR_PHY_STEPS = int(ARGS.simulation_freq_hz / ARGS.control_freq_hz)  # fly_INDI.py:139-141
    INIT_XYZS = np.array([[0.0, 1.0, 0.5]])                               # :147
    INIT_RPYS = np.array([[0.0, 0.0, 0.0]])
    env = CtrlAviary(drone_model=[ARGS.drone], num_drones=1, initial_xyzs=INIT_XYZS, initial_rpys=INIT_RPYS,
                     physics=Physics(ARGS.physics), neighbourhood_radius=10, freq=ARGS.simulation_freq_hz,
                     aggregate_phy_steps=AGGR_PHY_STEPS, num_envs=ARGS.num_envs,
                     ground_plane=(ARGS.physics != "dyn"))  # the PyBullet world has a floor (BaseAviary.py:679-680)
    ctrl = [INDIControl(drone_model=ARGS.drone, num_envs=ARGS.num_envs)]
    NUM_WP = ARGS.control_freq_hz * 15                                    # :151-167
    TARGET_RPYS = np.array([[0, 0, 0.4 + i / 200] for i in range(NUM_WP)])
    wp_counter = 0
    logger = Logger(logging_freq_hz=int(ARGS.simulation_freq_hz / AGGR_PHY_STEPS), num_drones=1,
                    duration_sec=ARGS.duration_sec)
    logger.attach(env)  # recorded on the device, collected once at the end
    CTRL_EVERY_N_STEPS = int(np.floor(env.SIM_FREQ / ARGS.control_freq_hz))  # :213
    action = {"0": np.array([0.4, 0.4, 0.4, 0.4])}                        # :214
    env.reset()
    START = time.time()
    for i in range(0, int(ARGS.duration_sec * env.SIM_FREQ), AGGR_PHY_STEPS):   # :217
        obs, reward, done, info = env.step(action)                        # :223
        if i % CTRL_EVERY_N_STEPS == 0:                                   # :226
            state = obs["0"]["state"] if ARGS.num_envs == 1 else obs["state"][:, 0, :]
            cmd, _, _ = ctrl[0].computeControlFromState(control_timestep=CTRL_EVERY_N_STEPS * env.TIMESTEP, state=state,
                                                        target_pos=np.array([0.0, 0.0, 0.5]), target_rpy=TARGET_RPYS[wp_counter])
            action = {"0": cmd} if ARGS.num_envs == 1 else cmd.reshape(ARGS.num_envs, 1, -1)
            wp_counter = wp_counter + 1 if wp_counter < NUM_WP - 1 else 0  # :242-245
        if i % env.SIM_FREQ == 0:
            env.render()                                                  # :265-266
    T = logger.collect()
    pos = logger.states[0, 0:3, -1]
    print("[INFO] %d samples logged, %.2f s wall clock, final position %s" % (T, time.time() - START, np.round(pos, 4)))
    env.close()
    ctrl[0].close()
    return pos


if __name__ == "__main__":
    main()
