"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE'S OWN CODE.

Run only where the reference checkout exists (this container):

    python tests/golden/make_golden.py

It imports, unmodified, from /root/reference (behind oracle/ref_shims.py, which supplies the
three PyBullet math functions and an empty gym):

* dronesim/control/wls_alloc.py         -> wls_cases.npz   (incl. the __main__ known answer :381-408)
* dronesim/control/INDIControl.py       -> ctrl_<vehicle>.npz for robobee, tello, hexa_6DOF_simple
* dronesim/control/INDIControl_6DOF.py  -> ctrl_hexa_6DOF.npz
* dronesim/utils/math.py                -> quat_helpers.npz
* dronesim/utils/trajGen.py             -> traj_3gates.npz (the table of examples/fly_INDI_TrajectoryTrack.py:127-160)
* dronesim/envs/VelocityAviary.py       -> pre_velocity_<vehicle>.npz (``_preprocessAction`` :221-264, called unbound)
* dronesim/envs/RPYTAviary.py           -> pre_rpyt_<vehicle>.npz     (``_preprocessAction`` :180-193, called unbound)
* dronesim/envs/BaseAviary.py           -> dyn_<vehicle>.npz (``_dynamics`` :1767-1828, ``_drag`` :1705-1732, ``_downwash``
                                           :1736-1763, ``_groundEffect`` :1648-1699 called unbound on a stand-in ``self``
                                           with a recording stand-in for the module's ``p``)
                                        -> rotor_<vehicle>.npz (the LIVE ``_quad_copter_physics`` :1477-1543 /
                                           ``_morphing_hexa_physics`` :1389-1457, noise source zeroed)
                                        -> rotor_tello_advanced.npz (the "advanced" branch :1493-1512 ->
                                           ``_get_prop_FMs`` :1570-1644 -> utils.calculate_propeller_forces_moments)

The fixtures are small .npz files; they are what travels to the GPU box (the reference does not).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pyb_math, ref_shims  # noqa: E402

ref_shims.install()


def rand_state_sequence(rng, n_u, T, ang_lim=1.0):
    """A smooth random walk of T aviary state vectors (BaseAviary.py:780-790 layout)."""
    pos = rng.uniform(-2, 2, 3)
    rpy = rng.uniform(-ang_lim, ang_lim, 3)
    vel = rng.normal(0, 0.5, 3)
    angv = rng.normal(0, 0.5, 3)
    states = np.zeros((T, 16 + n_u))
    for t in range(T):
        q = pyb_math.getQuaternionFromEuler(rpy)
        states[t, 0:3] = pos
        states[t, 3:7] = q
        states[t, 7:10] = pyb_math.getEulerFromQuaternion(q)
        states[t, 10:13] = vel
        states[t, 13:16] = angv
        pos = pos + 0.02 * vel
        vel = vel + rng.normal(0, 0.05, 3)
        angv = angv + rng.normal(0, 0.2, 3)
        rpy = np.clip(rpy + 0.02 * angv, -1.3, 1.3)
    return states


def controller_fixture(make_ctrl, n_u, seed, T=40, n_seq=6):
    rng = np.random.default_rng(seed)
    out = dict(states=[], dt=[], tpos=[], tvel=[], tacc=[], trpy=[], cmd=[], pos_e=[], yaw_err=[],
               last_vel=[], last_rates=[], last_thrust=[])
    for s in range(n_seq):
        ctrl = make_ctrl()
        dt = [5 / 240, 2 / 240, 8 / 240][s % 3]
        states = rand_state_sequence(rng, n_u, T, ang_lim=[0.3, 1.0][s % 2])
        seq = {k: [] for k in out}
        for t in range(T):
            tpos = states[t, 0:3] + rng.normal(0, 0.3, 3)
            tvel = rng.normal(0, 0.3, 3)
            tacc = rng.normal(0, 0.3, 3)
            trpy = np.array([0.0, 0.0, rng.uniform(-3.5, 3.5)])
            cmd, pos_e, yaw_err = ctrl.computeControlFromState(
                control_timestep=dt, state=states[t].copy(), target_pos=tpos, target_vel=tvel,
                target_acc=tacc, target_rpy=trpy)
            for k, v in (("states", states[t]), ("dt", dt), ("tpos", tpos), ("tvel", tvel), ("tacc", tacc),
                         ("trpy", trpy), ("cmd", np.array(cmd, float)), ("pos_e", np.array(pos_e, float)),
                         ("yaw_err", float(yaw_err)), ("last_vel", np.array(ctrl.last_vel, float)),
                         ("last_rates", np.array(ctrl.last_rates, float)),
                         ("last_thrust", float(ctrl.last_thrust))):
                seq[k].append(np.array(v, dtype=np.float64))
        for k in out:
            out[k].append(np.array(seq[k]))
    return {k: np.array(v) for k, v in out.items()}


def hover_fixture():
    """The SURVEY 8(c) probe: robobee at rest at z=0.5, target yaw 0.4, dt=5/240."""
    c = ref_shims.quad_controller("robobee")
    st = np.zeros(20)
    st[2] = 0.5
    st[6] = 1.0
    cmds = []
    for _ in range(3):
        cmd, _, _ = c.computeControlFromState(control_timestep=5 / 240, state=st,
                                              target_pos=np.array([0, 0, 0.5]), target_rpy=np.array([0, 0, 0.4]))
        cmds.append(np.array(cmd))
    return np.array(cmds)


def wls_fixture():
    wls = ref_shims.wls_alloc()
    # --- the reference's own known-answer case (wls_alloc.py:381-404)
    umin = np.zeros(6)
    umax = np.ones(6) * 9600.0
    uc = np.array([4614, 4210, 4210, 4614, 4210, 4210], float)
    dumin, dumax = umin - uc, umax - uc
    v = np.array([240, -240.5658, 600.0, 1.8532])
    Wv = np.array([100, 100, 1, 10], float)
    A = np.array([[0.0, -0.015, 0.015, 0.0, -0.015, 0.015], [0.015, -0.010, -0.010, 0.015, -0.010, -0.010],
                  [0.103, 0.103, 0.103, -0.103, -0.103, -0.103],
                  [-0.0009, -0.0009, -0.0009, -0.0009, -0.0009, -0.0009]])
    du, it = wls(v, dumin, dumax, A, None, None, Wv, None, dumin.copy())
    kat = dict(kat_v=v, kat_umin=dumin, kat_umax=dumax, kat_B=A, kat_Wv=Wv, kat_up=dumin, kat_du=du,
               kat_iter=it,
               kat_matlab=np.array([-4614.0, 426.064612091305, 5390.0, -4614.0, -4210.0, 5390.0]))
    # --- random hexa-shaped cases through the same call the 6-DOF controller makes (:607-628).  Inputs are rounded to
    # float32 first (what the CUDA core is handed), so both sides solve bit-identical problems; the final working set W
    # is a local of the reference function (:171), read off its frame when it returns (the function itself is unmodified)
    from dronesim_b200.vehicles import parse_urdf
    vt = parse_urdf(os.path.join(ref_shims.REFERENCE_ROOT, "dronesim", "assets", "hexa_6DOF.urdf"))
    B = vt.G1 / 0.05
    Wv6 = np.array([1000, 1000, 0.1, 10, 10, 100], float)
    rng = np.random.default_rng(7)
    # rnd_stale marks the runs that pass through the reference's stale-alpha path: a feasible iterate whose multipliers
    # are not all >= -FLT_EPSILON does not `break`, and falls into the step-length search WITHOUT resetting `alpha`
    # (only the infeasible branch does, :299-302), so a leftover step length of an earlier iteration is applied.  From
    # there on the multipliers of bound variables are O(1e18) with O(1e4) rounding residue on the others, and the sign
    # of that residue - i.e. LAPACK gelsd's rounding inside np.linalg.lstsq (:252) - decides how many more iterations
    # follow: replacing lstsq by an equally exact QR solve changes the iteration count of 10-14 % of these runs and of
    # none of the others.  Integer-exactness is demanded on the regular runs; the stale ones are kept for coverage.
    import inspect
    src_lines, first = inspect.getsourcelines(wls)
    search_line = first + next(i for i, l in enumerate(src_lines) if "Find the lowest distance from the limit" in l) + 1
    n_cases = 12288
    vs, cmds = np.zeros((n_cases, 6), np.float32), np.zeros((n_cases, 6), np.float32)
    dus, its = np.zeros((n_cases, 6)), np.zeros(n_cases, np.int32)
    ok, Ws, stale = np.zeros(n_cases, bool), np.zeros((n_cases, 6), np.int8), np.zeros(n_cases, bool)
    for k in range(n_cases):
        sigma = [1.0, 10.0, 50.0, 150.0, 400.0, 3.0, 25.0, 80.0][k % 8]
        centre = [0.5, 0.95, 0.05, 0.5, 0.0, 1.0, 0.3, 0.7][(k // 8) % 8]
        v6 = rng.normal(0, sigma, 6).astype(np.float32)
        cmd = np.clip(centre + rng.normal(0, 0.03, 6), 0, 1).astype(np.float32)
        (du, it), loc = call_with_locals(wls, v6.astype(float), 0.0 - cmd.astype(float), 1.0 - cmd.astype(float), B, None,
                                         None, Wv6, np.ones(6), None,
                                         watch=(search_line, lambda f: f.f_locals.get("n_infeasible", 1) == 0))
        vs[k], cmds[k], its[k], ok[k] = v6, cmd, it, du is not None
        dus[k] = np.zeros(6) if du is None else np.array(du)
        Ws[k] = np.asarray(loc["W"]).astype(np.int8)
        stale[k] = loc["__watch_hit__"]
    return dict(kat, rnd_v=vs, rnd_cmd=cmds, rnd_du=dus, rnd_iter=its, rnd_ok=ok, rnd_W=Ws, rnd_stale=stale, rnd_B=B,
                rnd_Wv=Wv6)


def call_with_locals(fn, *args, watch=None):
    """Run the UNMODIFIED ``fn(*args)`` and also return the locals of its frame at the moment it returns.
    ``watch = (line_number, predicate(frame))``: ``__watch_hit__`` tells whether that source line was ever reached
    with the predicate true."""
    import sys

    box = {"__watch_hit__": False}
    code = fn.__code__

    def tracer(frame, event, arg):
        if frame.f_code is code:
            def local(frame, event, arg):
                if event == "line" and watch is not None and frame.f_lineno == watch[0] and watch[1](frame):
                    box["__watch_hit__"] = True
                if event == "return":
                    box.update({k: (v.copy() if hasattr(v, "copy") else v) for k, v in frame.f_locals.items()})
                return local
            return local
        return None

    old = sys.gettrace()
    sys.settrace(tracer)
    try:
        out = fn(*args)
    finally:
        sys.settrace(old)
    return out, box


def quat_fixture():
    from dronesim.utils import math as rm
    rng = np.random.default_rng(3)
    q1 = rng.normal(size=(64, 4))
    q1 /= np.linalg.norm(q1, axis=1, keepdims=True)
    q2 = rng.normal(size=(64, 4))
    q2 /= np.linalg.norm(q2, axis=1, keepdims=True)
    inv = np.array([rm.quat_inv_comp(a, b) for a, b in zip(q1, q2)])
    comp = np.array([rm.quat_comp(a, b) for a, b in zip(q1, q2)])
    wrap = np.array([rm.quat_wrap_shortest(e.copy()) for e in inv])
    ang = rng.uniform(-12, 12, 64)
    nang = np.array([rm.norm_ang(a) for a in ang])
    return dict(q1=q1, q2=q2, inv_comp=inv, comp=comp, wrap=wrap, ang=ang, norm_ang=nang)


def traj_fixture():
    """examples/fly_INDI_TrajectoryTrack.py:127-160,178-186 with control_freq_hz = 96."""
    from dronesim.utils.trajGen import trajGenerator
    gates = np.vstack((np.array([[-3.0, 0, 2]]), np.array([0.5, 1, 5]), np.array([3, 0, 2])))
    traj = trajGenerator(gates, max_vel=0.7, gamma=1e6)
    ts = np.arange(0, traj.TS[-1], 1 / 96)
    rows = []
    for ti in ts:
        s = traj.get_des_state(ti)
        rows.append(np.concatenate([s.pos, s.vel, s.acc, [s.yaw]]))
    return dict(TS=np.array(traj.TS), table=np.array(rows), gates=gates)


def preprocess_fixture(kind, name, seed, T=30, n_seq=4):
    """``VelocityAviary._preprocessAction`` (VelocityAviary.py:221-264) / ``RPYTAviary._preprocessAction``
    (RPYTAviary.py:180-193) executed UNBOUND on a stand-in ``self`` that carries exactly what the method reads:
    ``_getDroneStateVector``, ``ctrl`` (a reference INDIControl), ``AGGR_PHY_STEPS``, ``TIMESTEP``, ``SPEED_LIMIT``."""
    import types

    from dronesim.envs.RPYTAviary import RPYTAviary
    from dronesim.envs.VelocityAviary import VelocityAviary
    from dronesim_b200.vehicles import load_vehicle

    rng = np.random.default_rng(seed)
    vt = load_vehicle(name)
    method = VelocityAviary._preprocessAction if kind == "velocity" else RPYTAviary._preprocessAction
    out = dict(states=[], action=[], cmd=[], aggr=[], last_vel=[], last_rates=[], last_thrust=[])
    for sq in range(n_seq):
        ctrl = ref_shims.quad_controller(name)
        aggr = [5, 2, 8, 1][sq % 4]
        states = rand_state_sequence(rng, 4, T, ang_lim=[0.3, 1.0][sq % 2])
        seq = {k: [] for k in out}
        for t in range(T):
            fake = types.SimpleNamespace(
                _getDroneStateVector=lambda i, _s=states[t]: _s.copy(), ctrl=[ctrl], AGGR_PHY_STEPS=aggr,
                TIMESTEP=1.0 / 240, SPEED_LIMIT=[vt.MAX_SPEED_KMH * (1000 / 3600)])
            if kind == "velocity":
                a = np.concatenate([rng.normal(0, 1, 3), [rng.uniform(-1, 1)]])
                if t % 7 == 3:
                    a[0:3] = 0.0  # the zero-direction branch (:239-242)
            else:
                a = np.concatenate([rng.normal(0, 0.5, 3), [rng.uniform(0.0, 1.0)]])  # p, q, r set-point, thrust
            cmd = np.array(method(fake, {"0": a})["0"], dtype=np.float64)
            for k, v in (("states", states[t]), ("action", a), ("cmd", cmd), ("aggr", aggr),
                         ("last_vel", np.array(ctrl.last_vel, float)), ("last_rates", np.array(ctrl.last_rates, float)),
                         ("last_thrust", float(ctrl.last_thrust))):
                seq[k].append(np.array(v, dtype=np.float64))
        for k in out:
            out[k].append(np.array(seq[k]))
    return {k: np.array(v) for k, v in out.items()}


class _RecordingBullet:
    """Stand-in for the ``pybullet`` module object the reference's BaseAviary methods call: the three math functions
    plus recorders for the calls that hand forces / states to the physics engine."""

    LINK_FRAME = 1

    def __init__(self):
        self.forces, self.torques, self.reset_pose, self.reset_vel, self.link_pos = [], [], None, None, None
        self.getMatrixFromQuaternion = pyb_math.getMatrixFromQuaternion
        self.getQuaternionFromEuler = pyb_math.getQuaternionFromEuler
        self.getEulerFromQuaternion = pyb_math.getEulerFromQuaternion

    def applyExternalForce(self, body, link, forceObj, posObj, flags, physicsClientId):
        self.forces.append((int(link), np.array(forceObj, float), int(flags)))

    def applyExternalTorque(self, body, link, torqueObj, flags, physicsClientId):
        self.torques.append((int(link), np.array(torqueObj, float), int(flags)))

    def resetBasePositionAndOrientation(self, body, pos, quat, physicsClientId):
        self.reset_pose = (np.array(pos, float), np.array(quat, float))

    def resetBaseVelocity(self, body, vel, ang, physicsClientId):
        self.reset_vel = (np.array(vel, float), np.array(ang, float))

    def getLinkStates(self, body, linkIndices, computeLinkVelocity, computeForwardKinematics, physicsClientId):
        # the reference reads link_states[i, 0][2] only (BaseAviary.py:1672-1679): entry 0 = link world position
        return [[list(self.link_pos[i])] + [[0.0, 0.0, 0.0]] * 7 for i in linkIndices]


def dynamics_fixture(name, seed, n_cases=48):
    """``BaseAviary._dynamics / _drag / _downwash / _groundEffect`` (BaseAviary.py:1767-1828, 1705-1732, 1736-1763,
    1648-1699) are dead code INSIDE the reference (their ``self.KF, self.M, ...`` are never assigned, ``x_torque`` is
    unbound for a list ``DRONE_MODEL``) but their bodies execute unchanged when called UNBOUND on a stand-in ``self``
    that supplies those attributes, with the module's ``p`` replaced by a recorder.  The stand-in describes a quad
    whose arm matches its URDF rotor sites (``L / sqrt(2)`` = the rotor x offset, CF2X layout), so the fixture pins
    the formulas; the per-drone / rotor-geometry generalisations (repairs R1-R7 of oracle/dynamics.py) remain restated."""
    import types

    import dronesim.envs.BaseAviary as BA
    from dronesim_b200.vehicles import load_vehicle

    vt = load_vehicle(name)
    rng = np.random.default_rng(seed)
    rec = _RecordingBullet()
    BA.p = rec  # the module-level name the methods resolve at call time
    arm = abs(float(vt.rotor_pos[0][0])) * np.sqrt(2.0)
    out = {k: [] for k in ("pos", "rpy", "quat", "vel", "rates", "cmd", "rpm", "dyn_pos", "dyn_quat", "dyn_vel", "dyn_rates",
                           "drag_force", "others", "dw_force", "gnd_rpy", "gnd_forces", "gnd_heights")}
    for c in range(n_cases):
        pos = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0.02, 2.0)])
        rpy = rng.uniform(-0.6, 0.6, 3)
        quat = np.array(pyb_math.getQuaternionFromEuler(rpy))
        vel = rng.normal(0, 1.0, 3)
        rates = rng.normal(0, 1.0, 3)
        cmd = rng.uniform(0.2, 0.9, 4)
        rpm = np.array(vt.PWM2RPM_SCALE) * cmd + np.array(vt.PWM2RPM_CONST)  # BaseAviary.py:1487-1490
        others = pos + np.concatenate([rng.uniform(-0.6, 0.6, (3, 2)), rng.uniform(-0.8, 1.5, (3, 1))], axis=1)
        if c % 8 == 5:
            others[0, 0:2] = pos[0:2] + [11.0, 0.0]  # beyond the 10 m gate (:1752)
        self = types.SimpleNamespace(
            pos=np.vstack([pos, others]), quat=np.tile(quat, (4, 1)), rpy=np.tile(rpy, (4, 1)),
            vel=np.tile(vel, (4, 1)), rpy_rates=np.tile(rates, (4, 1)), NUM_DRONES=4, DRONE_IDS=[1, 2, 3, 4], CLIENT=0,
            KF=vt.KF, KM=vt.KM, M=vt.M, GRAVITY=9.8 * vt.M, L=arm, J=np.array(vt.J), J_INV=np.array(vt.J_INV),
            TIMESTEP=1.0 / 240, DRONE_MODEL=BA.DroneModel.CF2X, DRAG_COEFF=np.array(vt.DRAG_COEFF),
            DW_COEFF_1=vt.DW_COEFF_1, DW_COEFF_2=vt.DW_COEFF_2, DW_COEFF_3=vt.DW_COEFF_3, PROP_RADIUS=vt.PROP_RADIUS,
            GND_EFF_COEFF=vt.GND_EFF_COEFF, GND_EFF_H_CLIP=vt.GND_EFF_H_CLIP)
        # _dynamics
        BA.BaseAviary._dynamics(self, rpm, 0)
        dyn_pos, dyn_quat = rec.reset_pose
        dyn_vel = rec.reset_vel[0]
        dyn_rates = self.rpy_rates[0].copy()
        # _drag
        rec.forces = []
        BA.BaseAviary._drag(self, rpm, 0)
        assert len(rec.forces) == 1 and rec.forces[0][0] == 4 and rec.forces[0][2] == rec.LINK_FRAME
        drag_force = rec.forces[0][1]
        # _downwash
        rec.forces = []
        BA.BaseAviary._downwash(self, 0)
        dw_force = np.sum([f for _, f, _ in rec.forces], axis=0) if rec.forces else np.zeros(3)
        # _groundEffect: rotor link world positions from the URDF rotor sites (what getLinkStates reports)
        R = pyb_math.rotmat(quat)
        rec.link_pos = [pos + R.dot(np.array(vt.rotor_pos[i])) for i in range(4)] + [pos]
        grpy = rpy.copy()
        if c % 8 == 6:
            grpy[0] = 1.7  # beyond the pi/2 gate (:1687-1690)
        self.rpy = np.tile(grpy, (4, 1))
        rec.forces = []
        BA.BaseAviary._groundEffect(self, rpm, 0)
        gnd = np.zeros((4, 3))
        for link, f, _ in rec.forces:
            gnd[link] = f
        for k, v in (("pos", pos), ("rpy", rpy), ("quat", quat), ("vel", vel), ("rates", rates), ("cmd", cmd), ("rpm", rpm),
                     ("dyn_pos", dyn_pos), ("dyn_quat", dyn_quat), ("dyn_vel", dyn_vel), ("dyn_rates", dyn_rates),
                     ("drag_force", drag_force), ("others", others), ("dw_force", dw_force), ("gnd_rpy", grpy),
                     ("gnd_forces", gnd), ("gnd_heights", np.array([lp[2] for lp in rec.link_pos[:4]]))):
            out[k].append(np.array(v, float))
    return {k: np.array(v) for k, v in out.items()}


def rotor_fixture(name, seed, n_cases=24):
    """The LIVE rotor force models ``_quad_copter_physics`` (BaseAviary.py:1477-1543) / ``_morphing_hexa_physics``
    (:1389-1457) called unbound: PWM -> RPM map, KF rpm^2 / KM rpm^2, spin signs, the link every force / torque is
    applied to.  The unseeded ``np.random.normal`` noise source (:1429-1432, 1518-1525) is replaced by zeros for the
    duration of the call - the noise-off mode all parity runs use."""
    import types

    import dronesim.envs.BaseAviary as BA
    from dronesim_b200.vehicles import load_vehicle

    vt = load_vehicle(name)
    n_u = vt.INDI_ACTUATOR_NR
    rng = np.random.default_rng(seed)
    rec = _RecordingBullet()
    BA.p = rec
    drone = types.SimpleNamespace(PWM2RPM_SCALE=np.array(vt.PWM2RPM_SCALE), PWM2RPM_CONST=np.array(vt.PWM2RPM_CONST),
                                  KF=vt.KF, KM=vt.KM, INDI_ACTUATOR_NR=n_u, TYPE=vt.TYPE)
    self = types.SimpleNamespace(drones=[drone], DRONE_IDS=[1], CLIENT=0, quat=np.array([[0, 0, 0, 1.0]]), vel=np.zeros((1, 3)))
    fn = BA.BaseAviary._morphing_hexa_physics if n_u == 6 else BA.BaseAviary._quad_copter_physics
    out = dict(cmd=[], force_link=[], force=[], torque_link=[], torque=[])
    real_normal = np.random.normal
    np.random.normal = lambda loc, scale, size=None: np.zeros(size)
    try:
        for _ in range(n_cases):
            cmd = rng.uniform(0.0, 1.0, n_u)
            rec.forces, rec.torques = [], []
            fn(self, cmd, 0)
            out["cmd"].append(cmd)
            out["force_link"].append([l for l, _, _ in rec.forces])
            out["force"].append([f for _, f, _ in rec.forces])
            out["torque_link"].append([l for l, _, _ in rec.torques])
            out["torque"].append([t for _, t, _ in rec.torques])
            assert all(fl == rec.LINK_FRAME for _, _, fl in rec.forces + rec.torques)
    finally:
        np.random.normal = real_normal
    return {k: np.array(v, float) for k, v in out.items()}


def advanced_rotor_fixture(name, seed, n_cases=40):
    """``_quad_copter_physics``'s "advanced" branch (BaseAviary.py:1493-1512) -> ``_get_prop_FMs`` (:1570-1644) ->
    ``utils.calculate_propeller_forces_moments`` method 2, executed unbound for a quad whose TYPE carries "advanced"."""
    import types
    import warnings

    import dronesim.envs.BaseAviary as BA
    from dronesim_b200.vehicles import load_vehicle

    vt = load_vehicle(name)
    rng = np.random.default_rng(seed)
    rec = _RecordingBullet()
    BA.p = rec
    drone = types.SimpleNamespace(PWM2RPM_SCALE=np.array(vt.PWM2RPM_SCALE), PWM2RPM_CONST=np.array(vt.PWM2RPM_CONST),
                                  KF=vt.KF, KM=vt.KM, INDI_ACTUATOR_NR=4, TYPE=vt.TYPE + "_advanced")
    out = dict(cmd=[], quat=[], vel=[], force_link=[], force=[], torque_link=[], torque=[])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # the model warns beyond mu / lambda = 0.3 (utils.py:386-396)
        for c in range(n_cases):
            cmd = rng.uniform(0.2, 1.0, 4)
            quat = np.array(pyb_math.getQuaternionFromEuler(rng.uniform(-0.5, 0.5, 3)))
            vel = rng.normal(0, [0.02, 1.5, 4.0][c % 3], 3)  # below the 0.1 m/s substitution, moderate, fast
            self = types.SimpleNamespace(drones=[drone], DRONE_IDS=[1], CLIENT=0, quat=np.array([quat]), vel=np.array([vel]))
            self._get_prop_FMs = lambda rpm, n, _s=self: BA.BaseAviary._get_prop_FMs(_s, rpm, n)
            rec.forces, rec.torques = [], []
            BA.BaseAviary._quad_copter_physics(self, cmd, 0)
            for k, v in (("cmd", cmd), ("quat", quat), ("vel", vel), ("force_link", [l for l, _, _ in rec.forces]),
                         ("force", [f for _, f, _ in rec.forces]), ("torque_link", [l for l, _, _ in rec.torques]),
                         ("torque", [t for _, t, _ in rec.torques])):
                out[k].append(np.array(v, float))
    return {k: np.array(v) for k, v in out.items()}


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "rotor_tello_advanced.npz"), **advanced_rotor_fixture("tello", seed=600))
    for i, name in enumerate(["robobee", "tello", "hexa_6DOF", "hexa_6DOF_simple"]):
        np.savez_compressed(os.path.join(HERE, "rotor_%s.npz" % name), **rotor_fixture(name, seed=500 + i))
    for i, name in enumerate(["robobee", "tello"]):
        np.savez_compressed(os.path.join(HERE, "dyn_%s.npz" % name), **dynamics_fixture(name, seed=400 + i))
    for kind in ("velocity", "rpyt"):
        for i, name in enumerate(["robobee", "tello"]):
            fx = preprocess_fixture(kind, name, seed=300 + i)
            np.savez_compressed(os.path.join(HERE, "pre_%s_%s.npz" % (kind, name)), **fx)
    np.savez_compressed(os.path.join(HERE, "wls_cases.npz"), **wls_fixture())
    np.savez_compressed(os.path.join(HERE, "quat_helpers.npz"), **quat_fixture())
    np.savez_compressed(os.path.join(HERE, "traj_3gates.npz"), **traj_fixture())
    np.savez_compressed(os.path.join(HERE, "hover_robobee.npz"), cmds=hover_fixture())
    for i, name in enumerate(["robobee", "tello", "hexa_6DOF_simple"]):
        n_u = 6 if "hexa" in name else 4
        fx = controller_fixture(lambda: ref_shims.quad_controller(name), n_u, seed=100 + i)
        np.savez_compressed(os.path.join(HERE, "ctrl_%s.npz" % name), **fx)
    fx = controller_fixture(lambda: ref_shims.hexa_controller("hexa_6DOF"), 6, seed=200)
    np.savez_compressed(os.path.join(HERE, "ctrl_hexa_6DOF.npz"), **fx)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
