"""GPU: the drop-in facades (``dronesim_b200.envs.CtrlAviary``, ``dronesim_b200.control.INDIControl*``)
driven exactly like the reference's example scripts, against the FP64 oracle running the same loop."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from dronesim_b200.vehicles import load_vehicle  # noqa: E402
from helpers import angle_between  # noqa: E402
from oracle.sim import OracleSwarm  # noqa: E402


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_fly_indi_loop_through_the_facades():
    """examples/fly_INDI.py:139-262 with --physics dyn semantics: robobee, 240 Hz sim, 48 Hz control, 1 s."""
    _need_gpu()
    from dronesim_b200.control.INDIControl import INDIControl
    from dronesim_b200.envs.CtrlAviary import CtrlAviary, Physics

    SIM, CTRL = 240, 48
    AGGR = int(SIM / CTRL)
    INIT_XYZS = np.array([[0.0, 1.0, 0.5]])
    env = CtrlAviary(drone_model=["robobee"], num_drones=1, initial_xyzs=INIT_XYZS, initial_rpys=np.zeros((1, 3)),
                     physics=Physics.DYN, neighbourhood_radius=10, freq=SIM, aggregate_phy_steps=AGGR, gui=False)
    assert env.AGGR_PHY_STEPS == 5 and env.TIMESTEP == 1 / 240 and env.NUM_DRONES == 1
    assert env.drones[0].INDI_ACTUATOR_NR == 4 and env.drones[0].M == 0.75
    ctrl = [INDIControl(drone_model="robobee")]
    vt = load_vehicle("robobee")
    orc = OracleSwarm([vt], 1, integrator="rpy", composite=False, aggregate_phy_steps=AGGR)
    orc.reset(INIT_XYZS)
    NUM_WP = CTRL * 15
    TARGET_RPYS = np.array([[0, 0, 0.4 + i / 200] for i in range(NUM_WP)])
    wp = 0
    CTRL_EVERY_N_STEPS = int(np.floor(env.SIM_FREQ / CTRL))
    action = {"0": np.array([0.4, 0.4, 0.4, 0.4])}
    obs0 = env.reset()
    assert set(obs0) == {"0"} and obs0["0"]["state"].shape == (20,) and obs0["0"]["neighbors"].shape == (1,)
    act_o = np.zeros((1, 1, 6))
    act_o[0, 0, :4] = 0.4
    for i in range(0, 1 * SIM, AGGR):
        obs, reward, done, info = env.step(action)
        orc.physics_step(act_o)
        assert reward == -1 and done is False and info == {"answer": 42}
        if i % CTRL_EVERY_N_STEPS == 0:
            action["0"], pos_e, yaw_e = ctrl[0].computeControlFromState(
                control_timestep=CTRL_EVERY_N_STEPS * env.TIMESTEP, state=obs["0"]["state"], target_pos=np.array([0, 0, 0.5]),
                target_rpy=TARGET_RPYS[wp])
            act_o = orc.control_step(np.array([0, 0, 0.5]).reshape(1, 1, 3), tyaw=TARGET_RPYS[wp, 2].reshape(1, 1))
            np.testing.assert_allclose(action["0"], act_o[0, 0, :4], atol=5e-5)
            np.testing.assert_allclose(pos_e, orc.pos_e[0, 0], atol=1e-4)
            wp = wp + 1 if wp < NUM_WP - 1 else 0
    assert env.step_counter == orc.step_counter == 240
    st = obs["0"]["state"]
    ref = orc.state_vector(0, 0)
    assert np.abs(st[0:3] - ref[0:3]).max() <= 1e-4
    assert angle_between(st[3:7], ref[3:7]).max() <= 1e-4
    np.testing.assert_allclose(st[7:10], ref[7:10], atol=1e-4)
    np.testing.assert_allclose(st[16:20], ref[16:20], atol=5e-5)  # obs tail = last clipped action (PWM)
    np.testing.assert_allclose(env.pos[0], ref[0:3], atol=1e-4)
    np.testing.assert_allclose(ctrl[0].last_vel, orc.ctrl[0][0].last_vel, atol=1e-4)
    env.close()
    ctrl[0].close()


def test_hetero_pair_hexa_and_quad_facade():
    """A 6-rotor and a 4-rotor vehicle in one aviary (the reason the reference made actions dicts),
    Physics.PYB_GND_DRAG_DW, each flown by its own controller class."""
    _need_gpu()
    from dronesim_b200.control.INDIControl import INDIControl
    from dronesim_b200.control.INDIControl_6DOF import INDIControl as INDIControl6
    from dronesim_b200.envs import CtrlAviary, Physics

    models = ["hexa_6DOF", "tello"]
    init = np.array([[0.0, 0.0, 1.0], [0.0, 0.1, 2.2]])
    env = CtrlAviary(drone_model=models, num_drones=2, initial_xyzs=init, physics=Physics.PYB_GND_DRAG_DW, freq=240,
                     aggregate_phy_steps=2, neighbourhood_radius=1.5)
    ctrl = [INDIControl6(drone_model="hexa_6DOF"), INDIControl(drone_model="tello")]
    with pytest.raises(ValueError):
        INDIControl(drone_model="hexa_6DOF")  # 6 virtual controls need the 6-DOF class
    vts = [load_vehicle(m) for m in models]
    orc = OracleSwarm(vts, 1, integrator="quat", composite=True, gnd=True, drag=True, dw=True, aggregate_phy_steps=2,
                      neighbourhood_radius=1.5)
    orc.reset(init)
    action = {"0": np.full(6, 0.1), "1": np.full(4, 0.4)}
    act_o = np.zeros((1, 2, 6))
    act_o[0, 0, :], act_o[0, 1, :4] = 0.1, 0.4
    env.reset()
    tpos = init.copy()
    for i in range(48):
        obs, _, _, _ = env.step(action)
        orc.physics_step(act_o)
        for j in range(2):
            action[str(j)], _, _ = ctrl[j].computeControlFromState(control_timestep=2 / 240, state=obs[str(j)]["state"],
                                                                   target_pos=tpos[j], target_rpy=np.zeros(3))
        act_o = orc.control_step(tpos.reshape(1, 2, 3))
        np.testing.assert_allclose(action["0"], act_o[0, 0, :6], atol=1e-4)
        np.testing.assert_allclose(action["1"], act_o[0, 1, :4], atol=1e-4)
    assert obs["0"]["state"].shape == (22,) and obs["1"]["state"].shape == (20,)
    for j in range(2):
        ref = orc.state_vector(0, j)
        assert np.abs(obs[str(j)]["state"][0:3] - ref[0:3]).max() <= 1e-4
        assert angle_between(obs[str(j)]["state"][3:7], ref[3:7]).max() <= 1e-4
    adj = np.array([obs["0"]["neighbors"], obs["1"]["neighbors"]])
    exp = orc.adjacency_bits()[0]
    np.testing.assert_array_equal(adj, [[(exp[i] >> j) & 1 for j in range(2)] for i in range(2)])
    env.close()
    for c in ctrl:
        c.close()


def test_batched_facade_matches_single_env():
    """num_envs > 1 returns device tensors; every env evolves exactly like the single-env facade."""
    _need_gpu()
    from dronesim_b200.envs import CtrlAviary, Physics

    init = np.array([[0.0, 0.0, 1.0], [0.3, 0.0, 1.6]])
    kw = dict(drone_model=["robobee", "tello"], num_drones=2, initial_xyzs=init, physics=Physics.PYB_DW, aggregate_phy_steps=4)
    e1 = CtrlAviary(**kw)
    e8 = CtrlAviary(num_envs=8, **kw)
    rng = np.random.default_rng(5)
    for _ in range(6):
        a = {"0": 0.47 + rng.uniform(-0.05, 0.05, 4), "1": 0.49 + rng.uniform(-0.05, 0.05, 4)}
        o1, r1, d1, _ = e1.step(a)
        o8, r8, d8, _ = e8.step(a)
    assert o8["state"].shape == (8, 2, 22) and o8["state"].is_cuda and r8.shape == (8,) and d8.shape == (8,)
    s8 = o8["state"].cpu().numpy()
    for e in range(8):
        np.testing.assert_array_equal(s8[e, 0, :20], o1["0"]["state"].astype(np.float32))
        np.testing.assert_array_equal(s8[e, 1, :20], o1["1"]["state"].astype(np.float32))
    assert (r8.cpu().numpy() == -1).all() and not d8.cpu().numpy().any()
    e1.close()
    e8.close()


def test_velocity_aviary_closed_loop():
    """examples/fly_INDI_velocity.py-style loop: VelocityAviary with a quad pair + a hexa, 1 s, against the oracle
    (VelocityAviary._preprocessAction restated, then the AGGR_PHY_STEPS substeps)."""
    _need_gpu()
    from dronesim_b200.envs import Physics, VelocityAviary

    models = ["robobee", "tello", "hexa_6DOF"]
    init = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 1.2], [2.0, 0.5, 1.5]])
    AGGR = 5
    env = VelocityAviary(drone_model=models, num_drones=3, initial_xyzs=init, physics=Physics.PYB_DRAG, freq=240,
                         aggregate_phy_steps=AGGR)
    assert env.action_space["0"].shape == (4,) and abs(env.SPEED_LIMIT[0] - 30 / 3.6) < 1e-12
    vts = [load_vehicle(m) for m in models]
    orc = OracleSwarm(vts, 1, integrator="quat", composite=True, drag=True, aggregate_phy_steps=AGGR)
    orc.reset(init)
    env.reset()
    rng = np.random.default_rng(11)
    for i in range(48):
        va = np.concatenate([rng.normal(0, 1, (3, 3)), rng.uniform(0.0, 0.08, (3, 1))], axis=1)
        if i % 9 == 4:
            va[1, 0:3] = 0.0  # zero direction -> zero target velocity (VelocityAviary.py:239-242)
        obs, reward, done, info = env.step({str(j): va[j] for j in range(3)})
        act = orc.velocity_preprocess(va.reshape(1, 3, 4))
        orc.physics_step(act)
        assert reward == -1 and done is False and info == {"answer": 42}
    assert env.step_counter == orc.step_counter == 48 * AGGR
    for j in range(3):
        ref = orc.state_vector(0, j)
        st = obs[str(j)]["state"]
        assert np.abs(st[0:3] - ref[0:3]).max() <= 1e-4, (j, st[0:3], ref[0:3])
        assert angle_between(st[3:7], ref[3:7]).max() <= 1e-4
        np.testing.assert_allclose(st[16:], ref[16:], atol=1e-4)  # obs tail = the applied (preprocessed) PWM command
    env.close()


def test_rpyt_aviary_closed_loop():
    """RPYTAviary: body-rate + thrust actions through INDIControl._INDIRateControl, two quads, 0.5 s."""
    _need_gpu()
    from dronesim_b200.envs import Physics, RPYTAviary

    models = ["robobee", "tello"]
    init = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 1.2]])
    AGGR = 2
    env = RPYTAviary(drone_model=models, num_drones=2, initial_xyzs=init, physics=Physics.PYB, freq=240, aggregate_phy_steps=AGGR)
    vts = [load_vehicle(m) for m in models]
    orc = OracleSwarm(vts, 1, integrator="quat", composite=True, aggregate_phy_steps=AGGR)
    orc.reset(init)
    env.reset()
    rng = np.random.default_rng(12)
    thrust = np.array([0.2, 0.2])
    for i in range(60):
        thrust = np.clip(thrust + rng.normal(0, 0.01, 2), 0.0, 1.0)
        rt = np.concatenate([rng.normal(0, 0.2, (2, 3)), thrust[:, None]], axis=1)
        obs, _, _, _ = env.step({str(j): rt[j] for j in range(2)})
        orc.physics_step(orc.rate_preprocess(rt.reshape(1, 2, 4)))
    for j in range(2):
        ref = orc.state_vector(0, j)
        st = obs[str(j)]["state"]
        assert np.abs(st[0:3] - ref[0:3]).max() <= 1e-4
        assert angle_between(st[3:7], ref[3:7]).max() <= 1e-4
        np.testing.assert_allclose(st[16:20], ref[16:20], atol=1e-4)
    with pytest.raises(Exception):  # the 6-DOF law has no rate / thrust entry
        e2 = RPYTAviary(drone_model=["hexa_6DOF"], num_drones=1, initial_xyzs=np.array([[0, 0, 1.0]]))
        e2.step({"0": np.array([0, 0, 0, 0.3])})
    env.close()


def test_example_script_runs_like_the_reference_example():
    """examples/fly_INDI.py (the reference script's loop with the imports swapped): 1 s, robobee settles towards the
    hover target [0, 0, 0.5] from [0, 1, 0.5]; the batched variant returns the same flight for every env."""
    _need_gpu()
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "examples", "fly_INDI.py")
    spec = importlib.util.spec_from_file_location("fly_INDI_example", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    p1 = mod.main(["--duration_sec", "1"])
    # the law restarts from cmd = 0 (INDIControl.py:129) and the explicit dynamics have no ground contact: after 1 s the
    # vehicle has sagged ~0.6 m and is moving from y = 1 towards the target at y = 0 (the oracle flies the same, test_cfg1)
    assert np.isfinite(p1).all() and -0.5 < p1[2] < 0.5 and 0.2 < p1[1] < 0.9
    p4 = mod.main(["--duration_sec", "1", "--num_envs", "4"])
    np.testing.assert_allclose(p4, p1, atol=1e-6)


def _load_example(name):
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "examples", name + ".py")
    spec = importlib.util.spec_from_file_location(name + "_example", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_example_trajectory_track_reaches_the_final_gate():
    """examples/fly_INDI_TrajectoryTrack.py (BASELINE configs[1]): the robobee follows the reference trajGenerator's table
    through the three gates and stops within 0.3 m of the last one (:249-250); 4 envs fly the same flight."""
    _need_gpu()
    mod = _load_example("fly_INDI_TrajectoryTrack")
    o1 = mod.main([])
    # the law leads the 12.5 s table (velocity / acceleration feed-forward): the 0.3 m stop fires after ~8.5 s
    assert o1["reached_final_gate_at_s"] is not None and 6.0 < o1["reached_final_gate_at_s"] < 12.5
    assert o1["rms_tracking_error_m"] < 0.4
    assert np.linalg.norm(np.array(o1["final_position"]) - np.array([3.0, 0.0, 2.0])) < 0.35
    o4 = mod.main(["--num_envs", "4"])
    assert o4["reached_final_gate_at_s"] == o1["reached_final_gate_at_s"]
    np.testing.assert_allclose(o4["final_position"], o1["final_position"], atol=1e-6)


def test_example_hexa_6dof_circle():
    """examples/fly_hexa_6DOF.py (BASELINE configs[2]): 6-DOF law + WLS, ground effect + drag + downwash flags, level flight."""
    _need_gpu()
    mod = _load_example("fly_hexa_6DOF")
    o = mod.main(["--duration_sec", "6"])
    # the reference gains trail the 1 m/s circle set-point by most of a metre (recorded in profiles/hover_tracking_r02.json)
    assert o["rms_tracking_error_m_after_2s"] < 1.2
    assert o["max_abs_roll_pitch_rad_after_2s"] < 0.08   # the tilted-rotor hexa translates without banking
    assert abs(o["final_position"][2] - 0.6) < 0.1
    o3 = mod.main(["--duration_sec", "3", "--num_envs", "3"])
    assert np.isfinite(o3["final_position"]).all()


def test_example_velocity_commands():
    """examples/fly_INDI_velocity.py: five tellos reach the commanded velocity (2 % of the speed limit along (1,1,1))."""
    _need_gpu()
    mod = _load_example("fly_INDI_velocity")
    o = mod.main(["--duration_sec", "6"])
    assert o["final_velocity_error_max"] < 0.02
    assert all(d > 0.2 for d in o["displacement_mean"])


def test_ground_plane_and_auto_reset():
    """DS_FLAG_GROUND_PLANE against the oracle's floor, and the batched auto-reset of finished envs."""
    _need_gpu()
    from dronesim_b200.core import SwarmCore
    from dronesim_b200.envs.CtrlAviary import CtrlAviary, Physics

    # (a) a robobee dropped from 0.4 m with the controller restarting from cmd = 0 touches down, rests, lifts off
    vt = load_vehicle("robobee")
    core = SwarmCore([vt], 2, aggregate_phy_steps=5, ground_plane_z=0.0)
    orc = OracleSwarm([vt], 2, integrator="quat", composite=True, aggregate_phy_steps=5)
    orc.floor_z = 0.0
    pos0 = np.array([[0.0, 0.0, 0.4], [0.3, 0.0, 0.2]])
    core.reset(pos0)
    orc.reset(pos0)
    tgt = core.targets_per_vehicle(np.array([[0.0, 0.0, 0.5, 0.0], [0.3, 0.0, 0.5, 0.0]]))
    act = np.zeros((2, 1, 6))
    zmin = 1.0
    for step in range(96):  # 2 s
        core.step(tgt, 1)
        orc.physics_step(act)
        act = orc.control_step(np.array([[0.0, 0.0, 0.5], [0.3, 0.0, 0.5]]).reshape(2, 1, 3))
        zmin = min(zmin, float(core.views()["pos"][:, 2].min()))
    assert zmin >= -1e-6 and zmin < 1e-3                       # it did reach the floor and never went through it
    np.testing.assert_allclose(core.views()["pos"].cpu().numpy(), orc.pos.reshape(2, 3), atol=2e-4)
    assert core.views()["pos"][:, 2].min() > 0.3               # and took off again towards the set-point
    core.close()
    # (b) auto-reset: envs whose done fires are put back to their initial pose on the device, the others keep flying
    E = 6
    xyz = np.zeros((E, 1, 3))
    xyz[:, 0, 2] = [1.0, 1.0, 0.05, 1.0, 0.05, 1.0]
    env = CtrlAviary(drone_model=["robobee"], num_drones=1, initial_xyzs=xyz, physics=Physics.PYB, aggregate_phy_steps=4,
                     num_envs=E, z_min=0.02, auto_reset=True)
    env.reset()
    a = np.zeros((E, 1, 6), dtype=np.float32)
    saw_done = np.zeros(E, dtype=bool)
    for _ in range(12):
        obs, rew, done, info = env.step(a)
        saw_done |= done.cpu().numpy()
    assert saw_done[[2, 4]].all() and not saw_done[[0, 1, 3, 5]].any()
    z = env.state_tensor()[:, 0, 2].cpu().numpy()
    assert (z[[2, 4]] > 0.0).all() and (z[[2, 4]] <= 0.05 + 1e-6).all()    # restarted from 0.05 m, still falling from there
    assert (z[[0, 1, 3, 5]] < 0.95).all()                                 # never reset: 48 substeps of free fall from 1 m
    env.close()


def test_swarm_controller_equals_the_per_drone_controllers():
    """``SwarmINDIControl`` (one launch for a mixed drone list) returns what the reference-style per-drone controller
    objects return, call for call, over a closed loop with the aviary."""
    _need_gpu()
    from dronesim_b200.control.INDIControl import INDIControl
    from dronesim_b200.control.INDIControl_6DOF import INDIControl as INDIControl6
    from dronesim_b200.control.SwarmControl import SwarmINDIControl
    from dronesim_b200.envs.CtrlAviary import CtrlAviary, Physics

    models = ["robobee", "hexa_6DOF", "tello"]
    xyz = np.array([[0.0, 0.0, 1.0], [1.5, 0.0, 1.2], [3.0, 0.0, 1.4]])
    tgt = xyz + np.array([0.2, -0.1, 0.1])
    envs = [CtrlAviary(drone_model=models, num_drones=3, initial_xyzs=xyz, physics=Physics.PYB_GND_DRAG_DW, aggregate_phy_steps=4)
            for _ in range(2)]
    per = [INDIControl(drone_model="robobee"), INDIControl6(drone_model="hexa_6DOF"), INDIControl(drone_model="tello")]
    swarm = SwarmINDIControl(models)
    acts = [{"0": np.full(4, 0.4), "1": np.full(6, 0.45), "2": np.full(4, 0.4)} for _ in range(2)]
    for e in envs:
        e.reset()
    for step in range(40):
        obs_a, _, _, _ = envs[0].step(acts[0])
        obs_b, _, _, _ = envs[1].step(acts[1])
        for j in range(3):
            acts[0][str(j)], _, _ = per[j].computeControlFromState(control_timestep=4 / 240, state=obs_a[str(j)]["state"],
                                                                   target_pos=tgt[j], target_rpy=np.array([0, 0, 0.2]))
        acts[1], pos_e, yaw_e = swarm.computeControlFromState(4 / 240, obs_b, tgt, target_rpy=np.array([0, 0, 0.2]))
        for j in range(3):
            np.testing.assert_array_equal(acts[1][str(j)], acts[0][str(j)])
    np.testing.assert_array_equal(envs[0].pos, envs[1].pos)
    for c in per + [swarm] + envs:
        c.close()
