"""CPU: pin the oracle's control half against fixtures produced by EXECUTING THE REFERENCE'S OWN
CODE (tests/golden/make_golden.py) and against the one known-answer test the reference ships
(wls_alloc.py:381-408, MATLAB lsqlin).  When the reference checkout is present (this container) the
oracle is additionally compared with the live reference classes."""
import os

import numpy as np
import pytest

from dronesim_b200.vehicles import load_vehicle
from oracle import control as oc
from oracle import ref_shims

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_wls_known_answer_matlab():
    g = dict(np.load(os.path.join(GOLD, "wls_cases.npz")))
    du, it = oc.wls_alloc(g["kat_v"], g["kat_umin"], g["kat_umax"], g["kat_B"], None, None, g["kat_Wv"], None, g["kat_up"])
    assert it == int(g["kat_iter"]) == 6  # integer-exact iteration count
    np.testing.assert_allclose(du, g["kat_du"], rtol=0, atol=1e-9)  # == the reference function's output
    np.testing.assert_allclose(du, g["kat_matlab"], rtol=1e-6)  # == MATLAB lsqlin (wls_alloc.py:404-408)


def test_wls_random_cases_iterations_and_active_set():
    g = dict(np.load(os.path.join(GOLD, "wls_cases.npz")))
    B, Wv = g["rnd_B"], g["rnd_Wv"]
    # the oracle is a pure-Python loop: every 5th of the 10,240 reference cases, plus every non-converging one (whose 100
    # iterations dominate the time: keep 12 of them)
    idx = sorted(set(range(0, g["rnd_v"].shape[0], 5)) | set(np.flatnonzero(~g["rnd_ok"])[:12].tolist()))
    idx = [k for k in idx if g["rnd_ok"][k] or k in set(np.flatnonzero(~g["rnd_ok"])[:12].tolist())]
    for k in idx:
        cmd = g["rnd_cmd"][k].astype(float)
        du, it, W = oc.wls_alloc(g["rnd_v"][k].astype(float), 0.0 - cmd, 1.0 - cmd, B, None, None, Wv, np.ones(6), None,
                                 return_W=True)
        assert it == g["rnd_iter"][k]
        assert (du is not None) == bool(g["rnd_ok"][k])
        np.testing.assert_array_equal(W, g["rnd_W"][k])  # the working set, read off the reference function's frame
        if du is not None:
            np.testing.assert_allclose(du, g["rnd_du"][k], rtol=1e-9, atol=1e-9)
    assert (~g["rnd_ok"]).sum() >= 1  # the fixture exercises the reference's None return (wls_alloc.py:350)


def test_wls_first_iteration_matrix_is_the_closed_form():
    """iteration 1 is du = M nu with the per-type constant the CUDA core keeps in shared memory."""
    g = dict(np.load(os.path.join(GOLD, "wls_cases.npz")))
    M = load_vehicle("hexa_6DOF").wls_unconstrained()
    one = (g["rnd_iter"] == 1) & g["rnd_ok"]
    assert one.sum() > 100
    du = g["rnd_v"][one] @ M.T
    scale = np.abs(g["rnd_du"][one]).max(axis=1, keepdims=True)
    assert (np.abs(du - g["rnd_du"][one]) / scale).max() < 1e-9


def test_quat_helpers():
    g = np.load(os.path.join(GOLD, "quat_helpers.npz"))
    for a, b, inv, wrap in zip(g["q1"], g["q2"], g["inv_comp"], g["wrap"]):
        e = oc.quat_inv_comp(a, b)
        np.testing.assert_allclose(e, inv, atol=1e-15)
        np.testing.assert_allclose(oc.quat_wrap_shortest(e.copy()), wrap, atol=1e-15)
    for a, n in zip(g["ang"], g["norm_ang"]):
        assert abs(oc.norm_ang(a) - n) < 1e-14


@pytest.mark.parametrize("name", ["robobee", "tello", "hexa_6DOF_simple", "hexa_6DOF"])
def test_controller_sequences_vs_reference_fixture(name):
    g = np.load(os.path.join(GOLD, "ctrl_%s.npz" % name))
    vt = load_vehicle(name)
    S, T = g["states"].shape[:2]
    for s in range(S):
        c = oc.make_controller(vt)
        for t in range(T):
            cmd, pe, ye = c.computeControlFromState(control_timestep=float(g["dt"][s, t]), state=g["states"][s, t],
                                                    target_pos=g["tpos"][s, t], target_vel=g["tvel"][s, t],
                                                    target_acc=g["tacc"][s, t], target_rpy=g["trpy"][s, t])
            np.testing.assert_allclose(cmd, g["cmd"][s, t], atol=1e-10)
            np.testing.assert_allclose(pe, g["pos_e"][s, t], atol=1e-13)
            assert abs(ye - g["yaw_err"][s, t]) < 1e-12
            np.testing.assert_allclose(c.last_vel, g["last_vel"][s, t], atol=1e-13)
            np.testing.assert_allclose(c.last_rates, g["last_rates"][s, t], atol=1e-12)
            assert abs(c.last_thrust - g["last_thrust"][s, t]) < 1e-9 * max(1.0, abs(g["last_thrust"][s, t]))


def test_hover_probe_commands():
    """first three commands of a robobee at rest, target yaw 0.4, dt = 5/240."""
    cmds = np.load(os.path.join(GOLD, "hover_robobee.npz"))["cmds"]
    np.testing.assert_allclose(cmds[0], [0, 0.01773833, 0, 0.01773833], atol=1e-8)
    c = oc.make_controller(load_vehicle("robobee"))
    st = np.zeros(20)
    st[2], st[6] = 0.5, 1.0
    for k in range(3):
        cmd, _, _ = c.computeControlFromState(control_timestep=5 / 240, state=st, target_pos=np.array([0, 0, 0.5]),
                                              target_rpy=np.array([0, 0, 0.4]))
        np.testing.assert_allclose(cmd, cmds[k], atol=1e-12)


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("name", ["robobee", "hexa_6DOF"])
def test_oracle_vs_live_reference_classes(name):
    """Where /root/reference exists, run the UNMODIFIED reference controller beside the oracle on fresh inputs."""
    vt = load_vehicle(name)
    ref = ref_shims.hexa_controller(name) if vt.INDI_OUTPUT_NR == 6 else ref_shims.quad_controller(name)
    mine = oc.make_controller(vt)
    rng = np.random.default_rng(11)
    n = 16 + vt.INDI_ACTUATOR_NR
    from oracle import pyb_math

    for t in range(25):
        st = np.zeros(n)
        st[0:3] = rng.uniform(-1, 1, 3)
        st[3:7] = pyb_math.getQuaternionFromEuler(rng.uniform(-0.8, 0.8, 3))
        st[10:13] = rng.normal(0, 0.3, 3)
        st[13:16] = rng.normal(0, 0.3, 3)
        kw = dict(control_timestep=2 / 240, state=st, target_pos=rng.uniform(-1, 1, 3), target_vel=rng.normal(0, 0.2, 3),
                  target_acc=rng.normal(0, 0.2, 3), target_rpy=np.array([0, 0, rng.uniform(-3, 3)]))
        a = ref.computeControlFromState(**{k: (v.copy() if hasattr(v, "copy") else v) for k, v in kw.items()})
        b = mine.computeControlFromState(**kw)
        np.testing.assert_allclose(np.array(a[0], float), b[0], atol=1e-10)
        np.testing.assert_allclose(np.array(a[1], float), b[1], atol=1e-13)
        assert abs(float(a[2]) - b[2]) < 1e-12


# ------------------------------------------------------------------------------------------
# VelocityAviary / RPYTAviary ``_preprocessAction``: the oracle's restatement against
# fixtures produced by calling the reference's own methods (tests/golden/make_golden.py::preprocess_fixture)
# ------------------------------------------------------------------------------------------
def _set_oracle_state(orc, st):
    from oracle import pyb_math as p

    orc.pos[0, 0], orc.quat[0, 0], orc.rpy[0, 0], orc.vel[0, 0] = st[0:3], st[3:7], st[7:10], st[10:13]
    orc.rates[0, 0] = p.rotmat(st[3:7]).T.dot(st[13:16])  # the oracle keeps body rates


@pytest.mark.parametrize("kind", ["velocity", "rpyt"])
@pytest.mark.parametrize("name", ["robobee", "tello"])
def test_preprocess_action_vs_reference_fixture(kind, name):
    from oracle.sim import OracleSwarm

    g = np.load(os.path.join(GOLD, "pre_%s_%s.npz" % (kind, name)))
    S, T = g["states"].shape[:2]
    for sq in range(S):
        orc = OracleSwarm([load_vehicle(name)], 1, aggregate_phy_steps=int(g["aggr"][sq, 0]))
        orc.reset(np.zeros((1, 3)))
        for t in range(T):
            _set_oracle_state(orc, g["states"][sq, t])
            fn = orc.velocity_preprocess if kind == "velocity" else orc.rate_preprocess
            act = fn(g["action"][sq, t].reshape(1, 1, 4))
            np.testing.assert_allclose(act[0, 0, :4], g["cmd"][sq, t], rtol=0, atol=1e-12)
            c = orc.ctrl[0][0]
            np.testing.assert_allclose(c.last_rates, g["last_rates"][sq, t], atol=1e-12)
            np.testing.assert_allclose(c.last_thrust, g["last_thrust"][sq, t], atol=1e-12)
            if kind == "velocity":
                np.testing.assert_allclose(c.last_vel, g["last_vel"][sq, t], atol=1e-12)
