"""CPU: the restated PyBullet math trio (third-party Bullet3, absent: parity unpinned) is checked for
self-consistency and against scipy's rotation conventions; the dynamics oracle for basic physics."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from dronesim_b200.vehicles import load_vehicle
from oracle import dynamics as od
from oracle import pyb_math as p
from oracle.sim import OracleSwarm, initial_waypoints, next_waypoint


def test_euler_quat_roundtrip_and_scipy():
    rng = np.random.default_rng(0)
    for _ in range(500):
        rpy = rng.uniform(-1.4, 1.4, 3)
        q = p.getQuaternionFromEuler(rpy)
        np.testing.assert_allclose(p.getEulerFromQuaternion(q), rpy, atol=1e-12)
        r = Rotation.from_euler("xyz", rpy)  # extrinsic xyz == intrinsic ZYX
        qs = r.as_quat()
        assert min(np.abs(np.array(q) - qs).max(), np.abs(np.array(q) + qs).max()) < 1e-14
        np.testing.assert_allclose(p.rotmat(q), r.as_matrix(), atol=1e-14)


def test_gimbal_branch():
    for sgn in (1.0, -1.0):
        q = p.getQuaternionFromEuler([0.0, sgn * np.pi / 2, 0.3])
        r, pt, y = p.getEulerFromQuaternion(q)
        assert r == 0.0 and abs(pt - sgn * np.pi / 2) < 1e-12
        np.testing.assert_allclose(p.rotmat(p.getQuaternionFromEuler([r, pt, y])), p.rotmat(q), atol=1e-6)


def test_rotmat_orthonormal_for_unnormalised_quat():
    rng = np.random.default_rng(1)
    for _ in range(100):
        q = rng.normal(size=4) * rng.uniform(0.5, 2.0)
        R = p.rotmat(q)
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-13)
        assert abs(np.linalg.det(R) - 1) < 1e-13


@pytest.mark.parametrize("name", ["robobee", "tello", "hexa_6DOF", "hexa_6DOF_simple"])
@pytest.mark.parametrize("integ", ["quat", "rpy"])
def test_free_fall_and_hover_balance(name, integ):
    vt = load_vehicle(name)
    pp = od.PhysParams(vt, composite=(integ == "quat"))
    # zero command: pure free fall with G = 9.8 (BaseAviary.py:182), semi-implicit Euler
    sw = OracleSwarm([vt], 1, integrator=integ, composite=(integ == "quat"), aggregate_phy_steps=10)
    sw.reset(np.array([[0.0, 0.0, 5.0]]))
    if (np.array(vt.PWM2RPM_CONST) == 0).all():
        sw.physics_step(np.zeros((1, 1, 6)))
        dt = 1 / 240
        np.testing.assert_allclose(sw.vel[0, 0], [0, 0, -9.8 * 10 * dt], atol=1e-12)
        np.testing.assert_allclose(sw.pos[0, 0, 2], 5.0 - 9.8 * dt * dt * 55, atol=1e-12)
    # total thrust at a uniform command equals sum kf rpm^2 along the rotor axes
    cmd = np.full(vt.INDI_ACTUATOR_NR, 0.5)
    F, tau, R = od.body_wrench(pp, cmd, 0.0, np.zeros(3), np.array([0, 0, 0, 1.0]), np.zeros(3), np.zeros(3), [], False,
                               False, False)
    rpm = od.rpm_of_cmd(pp, cmd)
    np.testing.assert_allclose(F, (pp.kf * rpm**2) @ pp.rotor_axis, atol=1e-12)
    assert abs(tau[2]) < 1e-9 * max(1.0, np.abs(F).max())  # alternating spins cancel at a uniform command


def test_downwash_only_from_above_and_decays():
    vt = load_vehicle("robobee")
    pp = od.PhysParams(vt, True)
    q0 = np.array([0, 0, 0, 1.0])
    args = (pp, np.zeros(4), 0.0)
    F_above, _, _ = od.body_wrench(*args, np.array([0, 0, 1.0]), q0, np.zeros(3), np.zeros(3), [np.array([0.0, 0, 1.5])], False, False, True)
    F_below, _, _ = od.body_wrench(*args, np.array([0, 0, 1.0]), q0, np.zeros(3), np.zeros(3), [np.array([0.0, 0, 0.5])], False, False, True)
    F_far, _, _ = od.body_wrench(*args, np.array([0, 0, 1.0]), q0, np.zeros(3), np.zeros(3), [np.array([10.5, 0, 1.5])], False, False, True)
    assert F_above[2] < 0 and F_below[2] == 0 and F_far[2] == 0
    alpha = pp.dw[0] * (pp.prop_radius / (4 * 0.5)) ** 2  # BaseAviary.py:1753
    assert abs(F_above[2] + alpha) < 1e-12


def test_ground_effect_clip_and_gate():
    vt = load_vehicle("hexa_6DOF")
    pp = od.PhysParams(vt, True)
    q0 = np.array([0, 0, 0, 1.0])
    cmd = np.full(6, 0.4)
    base, _, _ = od.body_wrench(pp, cmd, 0.0, np.array([0, 0, 0.05]), q0, np.zeros(3), np.zeros(3), [], False, False, False)
    lo, _, _ = od.body_wrench(pp, cmd, 0.0, np.array([0, 0, 0.0001]), q0, np.zeros(3), np.zeros(3), [], True, False, False)
    lo2, _, _ = od.body_wrench(pp, cmd, 0.0, np.array([0, 0, -0.5]), q0, np.zeros(3), np.zeros(3), [], True, False, False)
    assert lo[2] > base[2]
    np.testing.assert_allclose(lo, lo2, rtol=1e-2)  # both clipped to GND_EFF_H_CLIP (BaseAviary.py:235), up to the rotor z offsets
    flipped, _, _ = od.body_wrench(pp, cmd, 0.0, np.array([0, 0, 0.05]), q0, np.array([2.0, 0, 0]), np.zeros(3), [], True, False, False)
    np.testing.assert_allclose(flipped, base, atol=1e-15)  # |roll| >= pi/2: no ground effect (:1687-1690)


def test_integer_bookkeeping():
    wp = initial_waypoints(6, 720)  # fly_INDI.py:161-163
    np.testing.assert_array_equal(wp, [0, 120, 240, 360, 480, 600])
    w = np.array([0, 718, 719])
    np.testing.assert_array_equal(next_waypoint(w, 720), [1, 719, 0])  # fly_INDI.py:242-245
    vt = load_vehicle("robobee")
    sw = OracleSwarm([vt, vt, vt], 2, aggregate_phy_steps=5, neighbourhood_radius=1.0)
    pos = np.array([[[0, 0, 1.0], [0.5, 0, 1.0], [1.0, 0, 1.0]], [[0, 0, 1.0], [1.0, 0, 1.0], [5, 5, 5.0]]])
    sw.reset(pos)
    np.testing.assert_array_equal(sw.adjacency_bits(), [[0b011, 0b111, 0b110], [0b001, 0b010, 0b100]])  # strict <
    sw.physics_step(np.zeros((2, 3, 6)))
    assert sw.step_counter == 5  # BaseAviary.py:554


def test_philox_known_answers_and_normals():
    """The noise source (extension: the reference uses the unseeded global numpy generator) is Philox-4x32-10; the
    three known-answer vectors are the ones published with Random123 (kat_vectors, philox4x32 10 rounds)."""
    from oracle import noise as on

    assert on.philox4x32((0, 0, 0, 0), 0, 0) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert on.philox4x32((0xFFFFFFFF,) * 4, 0xFFFFFFFF, 0xFFFFFFFF) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert on.philox4x32((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), 0xA4093822, 0x299F31D0) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)
    z = np.array([on.normals12(v, s_, 1234) for v in range(200) for s_ in range(10)]).reshape(-1)
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03 and abs((z**3).mean()) < 0.1
    assert not np.array_equal(on.normals12(1, 2, 3), on.normals12(2, 1, 3))
