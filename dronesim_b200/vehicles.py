"""Vehicle parameter loading: URDF -> per-type constant tables for the CUDA core.

Mirrors, field for field, what the reference reads from its URDF files:

* ``BaseAviary._parseURDFParameters``  (dronesim/envs/BaseAviary.py:2041-2140) and the
  ``Drone`` dataclass (BaseAviary.py:69-95)  -> the ``TYPE, M, L, ... MAX_PWM`` fields below;
* ``INDIControl._parseURDFControlParameters`` (dronesim/control/INDIControl.py:55-106) and
  ``BaseControl.__init__`` (dronesim/control/BaseControl.py:37-42) -> ``m, G1, guidance gains,
  att/rate gains, PWM map``.

The reference then hands the URDF to PyBullet, which gets the rotor link poses and the
whole-tree mass distribution implicitly from the kinematic chain.  Since the CUDA core replaces
PyBullet on this path, ``VehicleType`` additionally extracts from the same URDF:

* rotor positions / thrust axes / spin signs, using PyBullet's depth-first link numbering
  (force application sites of ``_quad_copter_physics`` BaseAviary.py:1528-1543 = links
  ``0..n_u-1``; of ``_morphing_hexa_physics`` BaseAviary.py:1439-1457 = links ``1,3,..,11``);
* composite mass, centre of mass and inertia tensor of the whole tree (all joints at zero).

Sources: a directory holding the reference's ``*.urdf`` files (``assets_dir=``, the
``DRONESIM_ASSETS`` environment variable, or an importable ``dronesim`` package) - or, when
none is available (e.g. on a GPU box without the reference), the frozen copy of the FOUR
shipped vehicles in ``dronesim_b200/assets/vehicle_tables.json`` generated from those URDFs by
``tools/freeze_vehicle_tables.py``.
"""
from __future__ import annotations

import json
import math
import os
import xml.etree.ElementTree as etxml
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_FROZEN = os.path.join(_HERE, "assets", "vehicle_tables.json")

MAX_ROTORS = 6
LAW_QUAD = 0  # INDIControl.py        (n_v = 4, pinv allocation)
LAW_6DOF = 1  # INDIControl_6DOF.py   (n_v = 6, WLS allocation)


# --------------------------------------------------------------------------------------
# small rigid-transform helpers (URDF convention: rpy = fixed-axis XYZ)
# --------------------------------------------------------------------------------------
def _rpy_matrix(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ]
    )


def _floats(text) -> List[float]:
    return [float(s) for s in str(text).split(" ") if s != ""]


def _origin(elem):
    """(xyz, R) of an optional <origin> child."""
    xyz, rpy = np.zeros(3), np.zeros(3)
    if elem is not None:
        o = elem.find("origin")
        if o is not None:
            if "xyz" in o.attrib:
                xyz = np.array(_floats(o.attrib["xyz"]))
            if "rpy" in o.attrib:
                rpy = np.array(_floats(o.attrib["rpy"]))
    return xyz, _rpy_matrix(rpy)


@dataclass
class VehicleType:
    """One airframe type.  Upper-case names are the reference ``Drone`` dataclass fields."""

    name: str
    # ---- BaseAviary._parseURDFParameters (BaseAviary.py:2041-2140), verbatim semantics ----
    TYPE: str
    M: float
    L: float
    THRUST2WEIGHT_RATIO: float
    J: np.ndarray
    J_INV: np.ndarray
    KF: float
    KM: float
    COLLISION_H: float
    COLLISION_R: float
    COLLISION_Z_OFFSET: float
    MAX_SPEED_KMH: float
    GND_EFF_COEFF: float
    PROP_RADIUS: float
    DRAG_COEFF: np.ndarray
    DW_COEFF_1: float
    DW_COEFF_2: float
    DW_COEFF_3: float
    PWM2RPM_SCALE: List[float]
    PWM2RPM_CONST: List[float]
    INDI_ACTUATOR_NR: int
    INDI_OUTPUT_NR: int
    G1: np.ndarray
    MIN_PWM: List[float]
    MAX_PWM: List[float]
    # ---- INDIControl._parseURDFControlParameters (INDIControl.py:55-106) ----
    guidance_indi_pos_gain: float = 0.0
    guidance_indi_speed_gain: float = 0.0
    att_gain: np.ndarray = field(default_factory=lambda: np.zeros(3))  # indi_gains.att.p/q/r
    rate_gain: np.ndarray = field(default_factory=lambda: np.zeros(3))  # indi_gains.rate.p/q/r
    # ---- what PyBullet derives from the kinematic tree ----
    rotor_pos: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))  # base frame
    rotor_axis: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))  # thrust direction
    torque_axis: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))  # reaction torque dir
    rotor_spin: np.ndarray = field(default_factory=lambda: np.zeros(0))  # sign of KM*rpm^2
    M_TOTAL: float = 0.0
    COM: np.ndarray = field(default_factory=lambda: np.zeros(3))
    J_TOTAL: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))

    # ------------------------------------------------------------------
    @property
    def n_u(self) -> int:
        return self.INDI_ACTUATOR_NR

    @property
    def n_v(self) -> int:
        return self.INDI_OUTPUT_NR

    @property
    def law(self) -> int:
        """Which reference controller class flies this type in the examples: 6 virtual
        controls -> INDIControl_6DOF (examples/fly_hexa_6DOF.py), else INDIControl."""
        return LAW_6DOF if self.INDI_OUTPUT_NR == 6 else LAW_QUAD

    @property
    def GND_EFF_H_CLIP(self) -> float:
        """BaseAviary.py:226-235 (commented out there; the formula is the spec).  With
        MAX_RPM = sqrt(T2W*G*M/(4 KF)) and MAX_THRUST = 4 KF MAX_RPM^2 the expression reduces
        to 0.25 * PROP_RADIUS * sqrt(15 * GND_EFF_COEFF / 4)."""
        g = 9.8 * self.M
        max_rpm = math.sqrt((self.THRUST2WEIGHT_RATIO * g) / (4 * self.KF))
        max_thrust = 4 * self.KF * max_rpm**2
        return 0.25 * self.PROP_RADIUS * math.sqrt((15 * max_rpm**2 * self.KF * self.GND_EFF_COEFF) / max_thrust)

    def pinv_alloc(self) -> np.ndarray:
        """``np.linalg.pinv(G1 / 0.05)``  (INDIControl.py:459) - a per-type constant."""
        return np.linalg.pinv(self.G1 / 0.05)

    WLS_WV = (1000.0, 1000.0, 0.1, 10.0, 10.0, 100.0)  # INDIControl_6DOF.py:618
    WLS_GAMMA = 100000.0  # wls_alloc.py:125 (``gamma_sq``, used un-squared at :193-202)

    def wls_unconstrained(self) -> np.ndarray:
        """Matrix of the first ``wls_alloc`` iteration: with Wu = 1, up = 0 the unconstrained
        regularised least-squares step is ``du = M nu`` with M = pinv([gamma Wv B; I]) [:, :n_v]
        gamma Wv   (wls_alloc.py:190-259; the initial guess cancels)."""
        n_v, n_u = self.G1.shape
        wv = np.array(self.WLS_WV[:n_v]) if n_v == 6 else np.array([1000.0, 1000.0, 0.1, 10.0])
        W = self.WLS_GAMMA * np.diag(wv)
        A = np.vstack([W @ (self.G1 / 0.05), np.eye(n_u)])
        rhs = np.vstack([W, np.zeros((n_u, n_v))])
        return np.linalg.lstsq(A, rhs, rcond=None)[0]

    # ------------------------------------------------------------------
    def to_json(self) -> dict:
        out = {}
        for k, v in self.__dict__.items():
            out[k] = v.tolist() if isinstance(v, np.ndarray) else v
        return out

    @staticmethod
    def from_json(d: dict) -> "VehicleType":
        kw = {}
        for k, v in d.items():
            kw[k] = np.array(v, dtype=np.float64) if isinstance(v, list) and k not in (
                "PWM2RPM_SCALE", "PWM2RPM_CONST", "MIN_PWM", "MAX_PWM") else v
        return VehicleType(**kw)


# --------------------------------------------------------------------------------------
# URDF parsing
# --------------------------------------------------------------------------------------
def _pybullet_link_order(root):
    """Link frames in PyBullet's multibody numbering: base = -1, then depth-first pre-order
    over child joints in file order (Bullet3 URDF2Bullet ``ComputeParentIndices``)."""
    links = {l.attrib["name"]: l for l in root.findall("link")}
    joints = root.findall("joint")
    children: Dict[str, list] = {n: [] for n in links}
    has_parent = set()
    for j in joints:
        parent = j.find("parent").attrib["link"]
        child = j.find("child").attrib["link"]
        children[parent].append((j, child))
        has_parent.add(child)
    base = [n for n in links if n not in has_parent][0]
    order = []  # (name, link_elem, xyz_in_base, R_in_base)

    def visit(name, xyz, R):
        order.append((name, links[name], xyz, R))
        for j, child in children[name]:
            jx, jR = _origin(j)  # joint angle = 0 (revolute arms locked at their zero pose)
            visit(child, xyz + R @ jx, R @ jR)

    visit(base, np.zeros(3), np.eye(3))
    return order  # order[0] is the base (PyBullet index -1); order[k+1] is link index k


def parse_urdf(path: str) -> VehicleType:
    root = etxml.parse(path).getroot()
    name = os.path.splitext(os.path.basename(path))[0]

    # ---- literal restatement of BaseAviary._parseURDFParameters (BaseAviary.py:2048-2112) ----
    TYPE = str(root.find("configuration").attrib["type"])
    M = float(root.find("link/inertial/mass").attrib["value"])
    prop = root.find("properties")
    L = float(prop.attrib["arm"])
    T2W = float(prop.attrib["thrust2weight"])
    KF = float(prop.attrib["kf"])
    KM = float(prop.attrib["km"])
    inertia = root.find("link/inertial/inertia")
    J = np.diag([float(inertia.attrib["ixx"]), float(inertia.attrib["iyy"]), float(inertia.attrib["izz"])])
    J_INV = np.linalg.inv(J)
    coll = root.find("link/collision/geometry/cylinder")
    COLLISION_H = float(coll.attrib["length"])
    COLLISION_R = float(coll.attrib["radius"])
    COLLISION_Z_OFFSET = _floats(root.find("link/collision/origin").attrib["xyz"])[2]
    DRAG_XY = float(prop.attrib["drag_coeff_xy"])
    DRAG_Z = float(prop.attrib["drag_coeff_z"])
    indi = root.find("control/indi")
    n_u = int(indi.attrib["actuator_nr"])
    n_v = int(indi.attrib["output_nr"])
    G1 = np.zeros((n_v, n_u))
    control = root.find("control")
    for i in range(n_v):  # rows are the children control[1..n_v] (BaseAviary.py:2097-2100)
        vals = [str(k) for k in control[i + 1].attrib.values()]
        G1[i] = _floats(vals[0])
    vals = [str(k) for k in root.find("control/pwm/pwm2rpm").attrib.values()]  # attribute ORDER
    scale, const = _floats(vals[0]), _floats(vals[1])
    vals = [str(k) for k in root.find("control/pwm/limit").attrib.values()]
    min_pwm, max_pwm = _floats(vals[0]), _floats(vals[1])

    # ---- INDIControl._parseURDFControlParameters (INDIControl.py:79-92) ----
    gg = root.find("control/indi_guidance_gains/pos")
    att = root.find("control/indi_att_gains/att")
    rate = root.find("control/indi_att_gains/rate")

    # ---- kinematic tree: what PyBullet would see ----
    order = _pybullet_link_order(root)
    bx, bR = _origin(order[0][1].find("inertial"))  # base inertial frame = PyBullet "base" frame
    to_base = lambda x: bR.T @ (x - bx)  # noqa: E731
    m_tot, mc = 0.0, np.zeros(3)
    parts = []
    for (_, link, xyz, R) in order:
        inert = link.find("inertial")
        if inert is None:
            continue
        ix, iR = _origin(inert)
        m = float(inert.find("mass").attrib["value"])
        ie = inert.find("inertia").attrib
        I = np.array(
            [
                [float(ie.get("ixx", 0)), float(ie.get("ixy", 0)), float(ie.get("ixz", 0))],
                [float(ie.get("ixy", 0)), float(ie.get("iyy", 0)), float(ie.get("iyz", 0))],
                [float(ie.get("ixz", 0)), float(ie.get("iyz", 0)), float(ie.get("izz", 0))],
            ]
        )
        c = to_base(xyz + R @ ix)
        Rb = bR.T @ R @ iR
        parts.append((m, c, Rb @ I @ Rb.T))
        m_tot += m
        mc += m * c
    com = mc / m_tot
    J_tot = np.zeros((3, 3))
    for m, c, I in parts:
        d = c - com
        J_tot += I + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))

    if "morphing_hexa" in TYPE:  # BaseAviary.py:937, force sites :1442
        rotor_links = [2 * j + 1 for j in range(n_u)]
    else:  # "quad" and the generic branch: links 0..n_u-1  (BaseAviary.py:1528, 954)
        rotor_links = list(range(n_u))
    rpos, raxis, taxis = [], [], []
    for li in rotor_links:
        _, link, xyz, R = order[li + 1]
        ix, iR = _origin(link.find("inertial"))  # LINK_FRAME forces act at the link's inertial frame
        Rl = bR.T @ R @ iR
        rpos.append(to_base(xyz + R @ ix))
        raxis.append(Rl[:, 2])
        # reaction torque: hexa applies it on the prop link about its own z (BaseAviary.py:1451-1457);
        # the quad applies the summed z_torque on the base link about base z (BaseAviary.py:1537-1543)
        taxis.append(Rl[:, 2] if "morphing_hexa" in TYPE else np.array([0.0, 0.0, 1.0]))
    # spin signs: -t0 + t1 - t2 + t3 (BaseAviary.py:1527) / rotors 0,2,4 flipped (BaseAviary.py:1439-1440)
    spin = np.array([-1.0 if (j % 2 == 0) else 1.0 for j in range(n_u)])

    return VehicleType(
        name=name, TYPE=TYPE, M=M, L=L, THRUST2WEIGHT_RATIO=T2W, J=J, J_INV=J_INV, KF=KF, KM=KM,
        COLLISION_H=COLLISION_H, COLLISION_R=COLLISION_R, COLLISION_Z_OFFSET=COLLISION_Z_OFFSET,
        MAX_SPEED_KMH=float(prop.attrib["max_speed_kmh"]), GND_EFF_COEFF=float(prop.attrib["gnd_eff_coeff"]),
        PROP_RADIUS=float(prop.attrib["prop_radius"]), DRAG_COEFF=np.array([DRAG_XY, DRAG_XY, DRAG_Z]),
        DW_COEFF_1=float(prop.attrib["dw_coeff_1"]), DW_COEFF_2=float(prop.attrib["dw_coeff_2"]),
        DW_COEFF_3=float(prop.attrib["dw_coeff_3"]), PWM2RPM_SCALE=scale, PWM2RPM_CONST=const,
        INDI_ACTUATOR_NR=n_u, INDI_OUTPUT_NR=n_v, G1=G1, MIN_PWM=min_pwm, MAX_PWM=max_pwm,
        guidance_indi_pos_gain=float(gg.attrib["kp"]), guidance_indi_speed_gain=float(gg.attrib["kd"]),
        att_gain=np.array([float(att.attrib[k]) for k in "pqr"]),
        rate_gain=np.array([float(rate.attrib[k]) for k in "pqr"]),
        rotor_pos=np.array(rpos), rotor_axis=np.array(raxis), torque_axis=np.array(taxis), rotor_spin=spin,
        M_TOTAL=m_tot, COM=com, J_TOTAL=J_tot,
    )


# --------------------------------------------------------------------------------------
# lookup
# --------------------------------------------------------------------------------------
def _assets_dirs(assets_dir: Optional[str]) -> List[str]:
    dirs = []
    if assets_dir:
        dirs.append(assets_dir)
    if os.environ.get("DRONESIM_ASSETS"):
        dirs.append(os.environ["DRONESIM_ASSETS"])
    try:  # an installed reference package
        import importlib.util

        spec = importlib.util.find_spec("dronesim")
        if spec is not None and spec.submodule_search_locations:
            dirs.append(os.path.join(list(spec.submodule_search_locations)[0], "assets"))
    except Exception:
        pass
    return dirs


_cache: Dict[tuple, VehicleType] = {}

_FROZEN_PROPS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "propeller_tables.json")
ADVANCED_PROPELLER = "mamr-8x4.5"  # BaseAviary.py:1617


def load_propeller(name: str = ADVANCED_PROPELLER) -> dict:
    """The oblique-flow fit the "advanced" quad model evaluates (``Data_section5_ObliqueFlow[name]``,
    dronesim/database/propeller_database.py:232-552; used by BaseAviary._get_prop_FMs :1617-1627 through
    utils.calculate_propeller_forces_moments, method 2): 14 coefficients + the radius in metres
    (``diameter_in / 2 * 0.0254``, utils/utils.py:172-174).  Read from the frozen table (tools/freeze_vehicle_tables.py)."""
    with open(_FROZEN_PROPS) as f:
        tab = json.load(f)
    if name not in tab:
        print("[ERROR] in load_propeller(), no table for propeller '%s'" % name)
        raise KeyError(name)
    return {"coeff": [float(x) for x in tab[name]["section5_oblique_flow"]],
            "radius": float(tab[name]["diameter_in"]) / 2 * 0.0254}


def as_advanced(vt: VehicleType) -> VehicleType:
    """A copy of a quad type whose TYPE selects the oblique-flow propeller model ("advanced" in TYPE,
    BaseAviary.py:1493).  No shipped URDF declares it; this is how a user (or a test) opts in."""
    import dataclasses

    return dataclasses.replace(vt, name=vt.name + "_advanced", TYPE=vt.TYPE + "_advanced")


def load_vehicle(name: str, assets_dir: Optional[str] = None) -> VehicleType:
    """``name`` is what the reference passes in ``drone_model=[...]`` (URDF stem)."""
    key = (name, assets_dir)
    if key in _cache:
        return _cache[key]
    vt = None
    for d in _assets_dirs(assets_dir):
        p = os.path.join(d, name + ".urdf")
        if os.path.isfile(p):
            vt = parse_urdf(p)
            break
    if vt is None:
        with open(_FROZEN) as f:
            frozen = json.load(f)
        if name not in frozen:
            print("[ERROR] in load_vehicle(), no URDF and no frozen table for drone model '%s'" % name)
            raise KeyError(name)
        vt = VehicleType.from_json(frozen[name])
    _cache[key] = vt
    return vt
