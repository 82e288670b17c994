"""CPU: pin the oracle's dynamics half against fixtures produced by EXECUTING the reference's own
``BaseAviary._dynamics / _drag / _downwash / _groundEffect`` bodies (dead code inside the reference, but executable
unbound on a stand-in ``self``; tests/golden/make_golden.py::dynamics_fixture).  The stand-in is a quad whose arm
matches its URDF rotor sites, so what is pinned is every formula of the path for the 4-rotor airframes; the
per-drone / rotor-geometry generalisations that let the same formulas fly the hexa (repairs R1-R7 in
oracle/dynamics.py) remain restated."""
import os

import numpy as np
import pytest

from dronesim_b200.vehicles import load_vehicle
from oracle import dynamics as od
from oracle import pyb_math as p

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["robobee", "tello"])
def test_dynamics_substep_vs_reference_dynamics(name):
    """``_dynamics`` (BaseAviary.py:1767-1828): force model, gyroscopic term, semi-implicit Euler, Euler-angle
    integration, quaternion from the integrated angles."""
    g = np.load(os.path.join(GOLD, "dyn_%s.npz" % name))
    pp = od.PhysParams(load_vehicle(name), composite=False)
    for c in range(g["pos"].shape[0]):
        F, tau, R = od.body_wrench(pp, g["cmd"][c], 0.0, g["pos"][c], g["quat"][c], g["rpy"][c], g["vel"][c], [],
                                   False, False, False)
        pos, quat, rpy, vel, rates = od.substep_rpy(pp, 1.0 / 240, g["pos"][c], g["quat"][c], g["rpy"][c], g["vel"][c],
                                                    g["rates"][c], F, tau, R)
        np.testing.assert_allclose(pos, g["dyn_pos"][c], rtol=0, atol=1e-12)
        np.testing.assert_allclose(vel, g["dyn_vel"][c], rtol=0, atol=1e-12)
        np.testing.assert_allclose(rates, g["dyn_rates"][c], rtol=0, atol=1e-10)
        np.testing.assert_allclose(quat, g["dyn_quat"][c], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", ["robobee", "tello"])
def test_add_on_forces_vs_reference_methods(name):
    """``_drag`` (:1705-1732), ``_downwash`` (:1736-1763), ``_groundEffect`` (:1648-1699): the body-frame force each
    add-on contributes = wrench with the add-on minus wrench without."""
    g = np.load(os.path.join(GOLD, "dyn_%s.npz" % name))
    vt = load_vehicle(name)
    pp = od.PhysParams(vt, composite=False)
    gated_dw = gated_gnd = 0
    for c in range(g["pos"].shape[0]):
        args = (pp, g["cmd"][c])
        rpm_sum = float(np.sum(g["rpm"][c]))
        F0, _, _ = od.body_wrench(*args, rpm_sum, g["pos"][c], g["quat"][c], g["rpy"][c], g["vel"][c], [], False, False, False)
        Fd, _, _ = od.body_wrench(*args, rpm_sum, g["pos"][c], g["quat"][c], g["rpy"][c], g["vel"][c], [], False, True, False)
        np.testing.assert_allclose(Fd - F0, g["drag_force"][c], rtol=1e-9, atol=1e-15)
        Fw, _, _ = od.body_wrench(*args, rpm_sum, g["pos"][c], g["quat"][c], g["rpy"][c], g["vel"][c], list(g["others"][c]),
                                  False, False, True)
        np.testing.assert_allclose(Fw - F0, g["dw_force"][c], rtol=1e-9, atol=1e-13)
        gated_dw += int(not g["dw_force"][c].any())
        Fg, _, _ = od.body_wrench(*args, rpm_sum, g["pos"][c], g["quat"][c], g["gnd_rpy"][c], g["vel"][c], [], True, False, False)
        # the reference applies [0, 0, g_i] in each rotor's LINK frame = along the rotor axis (z for the quads)
        np.testing.assert_allclose(Fg - F0, g["gnd_forces"][c].sum(axis=0), rtol=1e-9, atol=1e-13)
        gated_gnd += int(not g["gnd_forces"][c].any())
        # the heights the oracle derives from the URDF rotor sites are the link heights the fixture handed the reference
        R = p.rotmat(g["quat"][c])
        h = [g["pos"][c][2] + R[2, :].dot(pp.rotor_pos[i]) for i in range(4)]
        np.testing.assert_allclose(h, g["gnd_heights"][c], atol=1e-14)
    assert gated_gnd >= 1  # the |roll| >= pi/2 gate is exercised


@pytest.mark.parametrize("name", ["robobee", "tello", "hexa_6DOF", "hexa_6DOF_simple"])
def test_rotor_forces_vs_live_reference_functions(name):
    """``_quad_copter_physics`` (BaseAviary.py:1477-1543) / ``_morphing_hexa_physics`` (:1389-1457) are LIVE reference code;
    the fixture records what they hand to PyBullet (noise source zeroed).  Pinned here: PWM -> RPM map, KF rpm^2,
    KM rpm^2 with the spin signs, and the link each force is applied to (what the loader takes rotor sites from)."""
    g = np.load(os.path.join(GOLD, "rotor_%s.npz" % name))
    vt = load_vehicle(name)
    pp = od.PhysParams(vt, composite=True)
    n_u = pp.n_u
    hexa = n_u == 6
    for c in range(g["cmd"].shape[0]):
        rpm = od.rpm_of_cmd(pp, g["cmd"][c])
        T, Q = pp.kf * rpm**2, pp.km * rpm**2
        # forces: [0, 0, T_j] in the LINK frame of link 2j+1 (hexa, :1442) / link j (quad, :1528)
        np.testing.assert_array_equal(g["force_link"][c], [2 * j + 1 for j in range(n_u)] if hexa else list(range(n_u)))
        np.testing.assert_allclose(g["force"][c][:, 2], T, rtol=1e-12)
        assert not g["force"][c][:, :2].any()
        if hexa:  # torque [0, 0, -/+ Q_j] on the same link, rotors 0, 2, 4 flipped (:1439-1440)
            np.testing.assert_array_equal(g["torque_link"][c], g["force_link"][c])
            np.testing.assert_allclose(g["torque"][c][:, 2], pp.spin * Q, rtol=1e-12)
        else:  # one base-link torque -t0 + t1 - t2 + t3 about base z (:1527, 1537-1543)
            np.testing.assert_array_equal(g["torque_link"][c], [-1])
            np.testing.assert_allclose(g["torque"][c][0, 2], float(np.sum(pp.spin * Q)), rtol=1e-10, atol=1e-16)
            np.testing.assert_allclose(pp.torque_axis, np.tile([0.0, 0.0, 1.0], (n_u, 1)))
        # the body wrench the oracle (and the CUDA core) integrates is exactly these link-frame vectors moved to the
        # centre of mass with the loader's link geometry
        F, tau, _ = od.body_wrench(pp, g["cmd"][c], 0.0, np.zeros(3), np.array([0, 0, 0, 1.0]), np.zeros(3), np.zeros(3), [],
                                   False, False, False)
        Fx = sum(g["force"][c][j, 2] * pp.rotor_axis[j] for j in range(n_u))
        tq = sum(np.cross(pp.rotor_pos[j] - pp.r_com, g["force"][c][j, 2] * pp.rotor_axis[j]) for j in range(n_u))
        tq = tq + (sum(g["torque"][c][j, 2] * pp.torque_axis[j] for j in range(n_u)) if hexa else g["torque"][c][0])
        np.testing.assert_allclose(F, Fx, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(tau, tq, rtol=1e-10, atol=1e-14)


def test_advanced_propeller_model_vs_reference_branch():
    """the oblique-flow propeller model ("advanced" in TYPE, BaseAviary.py:1493-1512, 1570-1644;
    utils/utils.py:149-202, 343-416) against the reference's own branch executed (rotor_tello_advanced.npz)."""
    from dronesim_b200.vehicles import as_advanced, load_propeller

    g = np.load(os.path.join(GOLD, "rotor_tello_advanced.npz"))
    vt = as_advanced(load_vehicle("tello"))
    assert "advanced" in vt.TYPE and vt.INDI_ACTUATOR_NR == 4
    pp = od.PhysParams(vt, composite=True)
    prop = load_propeller()
    assert len(prop["coeff"]) == 14 and abs(prop["radius"] - 0.1016) < 1e-12
    for c in range(g["cmd"].shape[0]):
        rpm = od.rpm_of_cmd(pp, g["cmd"][c])
        F_b, M_b = od.advanced_rotor_FMs(prop, g["quat"][c], g["vel"][c], rpm)
        np.testing.assert_array_equal(g["force_link"][c], [0, 1, 2, 3])
        np.testing.assert_array_equal(g["torque_link"][c], [0, 1, 2, 3])
        np.testing.assert_allclose(np.array(F_b), g["force"][c], rtol=1e-12, atol=1e-15)
        direction = np.array([-1.0, 1.0, -1.0, 1.0])
        np.testing.assert_allclose(np.array([m[2] for m in M_b]) * direction, g["torque"][c][:, 2], rtol=1e-12, atol=1e-16)
        assert not g["torque"][c][:, :2].any()
        F, tau, _ = od.body_wrench(pp, g["cmd"][c], 0.0, np.zeros(3), g["quat"][c], np.zeros(3), g["vel"][c], [], False, False, False)
        np.testing.assert_allclose(F, g["force"][c].sum(axis=0), rtol=1e-12)
        tq = sum(np.cross(pp.rotor_pos[i] - pp.r_com, g["force"][c][i]) + g["torque"][c][i] for i in range(4))
        np.testing.assert_allclose(tau, tq, rtol=1e-10, atol=1e-14)
