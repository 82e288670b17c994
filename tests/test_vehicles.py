"""CPU: URDF vehicle-table loader (BaseAviary._parseURDFParameters, INDIControl._parseURDFControlParameters)."""
import os

import numpy as np
import pytest

from dronesim_b200 import vehicles as V

ASSETS = "/root/reference/dronesim/assets"
NAMES = ["tello", "robobee", "hexa_6DOF", "hexa_6DOF_simple"]


@pytest.mark.parametrize("name", NAMES)
def test_frozen_table_has_the_reference_fields(name):
    vt = V.load_vehicle(name)
    n_u, n_v = vt.INDI_ACTUATOR_NR, vt.INDI_OUTPUT_NR
    assert (n_u, n_v) == {"tello": (4, 4), "robobee": (4, 4), "hexa_6DOF": (6, 6), "hexa_6DOF_simple": (6, 4)}[name]
    assert vt.G1.shape == (n_v, n_u)
    assert len(vt.PWM2RPM_SCALE) == len(vt.PWM2RPM_CONST) == len(vt.MIN_PWM) == len(vt.MAX_PWM) == n_u
    assert vt.rotor_pos.shape == (n_u, 3) and vt.rotor_axis.shape == (n_u, 3)
    np.testing.assert_allclose(np.linalg.norm(vt.rotor_axis, axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(vt.J_INV @ vt.J, np.eye(3), atol=1e-12)
    assert vt.M_TOTAL >= vt.M > 0
    assert vt.law == (V.LAW_6DOF if n_v == 6 else V.LAW_QUAD)
    # allocation constants are consistent with their definitions
    P = vt.pinv_alloc()
    assert P.shape == (n_u, n_v)
    np.testing.assert_allclose((vt.G1 / 0.05) @ P @ (vt.G1 / 0.05), vt.G1 / 0.05, atol=1e-8 * np.abs(vt.G1 / 0.05).max())


def test_hexa_composite_inertia_probe():
    """SURVEY section 7 step 1 [probed]: the hexa tree sums to 0.86 kg, J ~ diag(5.36e-3, 5.38e-3, 9.26e-3), CoM z ~ -0.011."""
    vt = V.load_vehicle("hexa_6DOF")
    assert abs(vt.M - 0.2) < 1e-12 and abs(vt.M_TOTAL - 0.86) < 1e-9
    np.testing.assert_allclose(np.diag(vt.J_TOTAL), [5.36e-3, 5.38e-3, 9.26e-3], rtol=5e-3)
    assert abs(vt.COM[2] + 0.011) < 1e-3
    # tilted rotors: thrust axes are proportional to the fx/fy rows of G1 (SURVEY section 4 cross-check)
    fx, fy = vt.G1[3], vt.G1[4]
    lat = vt.rotor_axis[:, :2]
    c = np.corrcoef(np.concatenate([lat[:, 0], lat[:, 1]]), np.concatenate([fx, fy]))[0, 1]
    assert abs(c) > 0.99


@pytest.mark.skipif(not os.path.isdir(ASSETS), reason="reference assets not present (GPU box)")
@pytest.mark.parametrize("name", NAMES)
def test_urdf_parse_equals_frozen_table(name):
    live = V.parse_urdf(os.path.join(ASSETS, name + ".urdf")).to_json()
    V._cache.clear()
    frozen = V.load_vehicle(name, assets_dir="/nonexistent").to_json()
    assert sorted(live) == sorted(frozen)
    for k in live:
        if isinstance(live[k], (list, float, int)) and not isinstance(live[k], str):
            np.testing.assert_allclose(np.array(live[k], float), np.array(frozen[k], float), rtol=0, atol=0, err_msg=k)
        else:
            assert live[k] == frozen[k], k


@pytest.mark.skipif(not os.path.isdir(ASSETS), reason="reference assets not present (GPU box)")
@pytest.mark.parametrize("name", ["robobee", "hexa_6DOF"])
def test_control_params_equal_the_reference_parser(name):
    """Execute the reference's own _parseURDFControlParameters (behind the shims) and compare."""
    from oracle import ref_shims

    vt = V.parse_urdf(os.path.join(ASSETS, name + ".urdf"))
    ref = ref_shims.hexa_controller(name) if vt.INDI_OUTPUT_NR == 6 else ref_shims.quad_controller(name)
    np.testing.assert_array_equal(np.array(ref.G1, float), vt.G1)
    assert ref.guidance_indi_pos_gain == vt.guidance_indi_pos_gain
    assert ref.guidance_indi_speed_gain == vt.guidance_indi_speed_gain
    np.testing.assert_array_equal([ref.indi_gains.att.p, ref.indi_gains.att.q, ref.indi_gains.att.r], vt.att_gain)
    np.testing.assert_array_equal([ref.indi_gains.rate.p, ref.indi_gains.rate.q, ref.indi_gains.rate.r], vt.rate_gain)
    np.testing.assert_array_equal(np.array(ref.MIN_PWM, float), vt.MIN_PWM)
    np.testing.assert_array_equal(np.array(ref.MAX_PWM, float), vt.MAX_PWM)


def test_unknown_model_reports_like_the_reference(capsys):
    with pytest.raises(KeyError):
        V.load_vehicle("no_such_drone", assets_dir="/nonexistent")
    assert "[ERROR]" in capsys.readouterr().out
