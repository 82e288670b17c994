"""``BaseAviary`` facade: the reference's gym-style environment API over the batched CUDA core.

Mirrors the constructor, attributes and ``reset`` / ``step`` contract of
``dronesim/envs/BaseAviary.py:129-198, 406-424, 428-555`` for ``num_envs`` independent copies of
the aviary at once (``num_envs=1`` reproduces the reference's shapes and dict-of-drones views
exactly).  Everything below ``step()`` - action clipping, the K physics substeps, the state
vector, the adjacency matrix - runs in ``libdronesim_b200.so``; PyBullet is not involved.

``physics`` selects the model exactly as the reference's ``Physics`` enum names it
(BaseAviary.py:41-49):

* ``Physics.DYN``  -> the literal ``_dynamics`` update (BaseAviary.py:1767-1828): Euler-angle
  integration, mass / inertia as the reference's parser reads them (first URDF link);
* ``Physics.PYB*`` -> rigid-body Newton-Euler about the whole-tree centre of mass with a
  quaternion update (what PyBullet's stepSimulation integrates, minus contacts), plus the
  add-ons the suffix names: ``_gnd`` ground effect (:1648-1699), ``_drag`` (:1705-1732),
  ``_dw`` downwash (:1736-1763).

GUI, video recording, obstacles and the vision attributes are out of scope (host-side
visualisation); the corresponding kwargs are accepted and ignored, like the reference does in
``p.DIRECT`` mode.
"""
from __future__ import annotations

import time
from enum import Enum

import numpy as np

from ..vehicles import VehicleType, load_vehicle


class Physics(Enum):
    """Physics implementations (same names / values as BaseAviary.py:41-49)."""

    PYB = "pyb"
    DYN = "dyn"
    PYB_GND = "pyb_gnd"
    PYB_DRAG = "pyb_drag"
    PYB_DW = "pyb_dw"
    PYB_GND_DRAG_DW = "pyb_gnd_drag_dw"


_FLAGS = {
    Physics.PYB: (False, False, False), Physics.DYN: (False, False, False), Physics.PYB_GND: (True, False, False),
    Physics.PYB_DRAG: (False, True, False), Physics.PYB_DW: (False, False, True), Physics.PYB_GND_DRAG_DW: (True, True, True),
}


class BaseAviary:
    """Base class of the batched aviary facades."""

    metadata = {"render.modes": ["human"]}

    def __init__(self, drone_model: list = ["tello"], num_drones: int = 1, neighbourhood_radius: float = np.inf,
                 initial_xyzs=None, initial_vels=None, initial_rpys=None, physics: Physics = Physics.PYB, freq: int = 240,
                 aggregate_phy_steps: int = 1, gui=False, record=False, obstacles=False, user_debug_gui=False,
                 vision_attributes=False, dynamics_attributes=False, *, num_envs: int = 1, device: int = 0,
                 goal=None, goal_radius: float = 0.3, z_min=None, max_steps: int = 0, ground_plane: bool = False,
                 auto_reset: bool = False):
        from ..core import SwarmCore

        #### Constants (BaseAviary.py:182-187)
        self.G = 9.8
        self.RAD2DEG = 180 / np.pi
        self.DEG2RAD = np.pi / 180
        self.SIM_FREQ = freq
        self.TIMESTEP = 1.0 / self.SIM_FREQ
        self.AGGR_PHY_STEPS = aggregate_phy_steps
        #### Parameters / options (:189-198)
        self.NUM_DRONES = num_drones
        self.NUM_ENVS = int(num_envs)
        self.NEIGHBOURHOOD_RADIUS = neighbourhood_radius
        self.DRONE_MODEL = drone_model
        self.GUI, self.RECORD, self.OBSTACLES, self.USER_DEBUG = False, False, False, False
        self.PHYSICS = physics if isinstance(physics, Physics) else Physics(physics)
        self.URDF = [drone + ".urdf" for drone in drone_model]
        if len(drone_model) != num_drones:
            print("[ERROR] in BaseAviary.__init__(), drone_model must list one URDF name per drone")
            raise ValueError("len(drone_model) != num_drones")
        #### self.drones (:219): VehicleType carries the Drone dataclass fields under the same names
        self.drones: list[VehicleType] = [load_vehicle(d) for d in drone_model]
        #### Initial poses (:360-389).  The reference's default uses the undefined self.L / self.COLLISION_H
        #### (SURVEY quirk Q8); the per-drone values are used here.
        if initial_xyzs is None:
            self.INIT_XYZS = np.array([[i * 4 * d.L, i * 4 * d.L, d.COLLISION_H / 2 - d.COLLISION_Z_OFFSET + 0.1]
                                       for i, d in enumerate(self.drones)])
        elif np.array(initial_xyzs).shape in ((num_drones, 3), (self.NUM_ENVS, num_drones, 3)):
            self.INIT_XYZS = np.array(initial_xyzs, dtype=float)
        else:
            print("[ERROR] invalid initial_xyzs in BaseAviary.__init__(), try initial_xyzs.reshape(NUM_DRONES,3)")
            raise ValueError("initial_xyzs")
        self.INIT_VELS = initial_vels
        if initial_rpys is None:
            self.INIT_RPYS = np.zeros((num_drones, 3))
        elif np.array(initial_rpys).shape in ((num_drones, 3), (self.NUM_ENVS, num_drones, 3)):
            self.INIT_RPYS = np.array(initial_rpys, dtype=float)
        else:
            print("[ERROR] invalid initial_rpys in BaseAviary.__init__(), try initial_rpys.reshape(NUM_DRONES,3)")
            raise ValueError("initial_rpys")
        gnd, drag, dw = _FLAGS[self.PHYSICS]
        integ = "rpy" if self.PHYSICS == Physics.DYN else "quat"
        self._core = SwarmCore(self.drones, self.NUM_ENVS, integrator=integ, ground=gnd, drag=drag, downwash=dw,
                               freq=float(freq), aggregate_phy_steps=int(aggregate_phy_steps),
                               neighbourhood_radius=float(neighbourhood_radius), gravity=self.G, device=device, goal=goal,
                               goal_radius=goal_radius, z_min=z_min, max_steps=max_steps,
                               ground_plane_z=(0.0 if ground_plane else None))
        #### Batched extension: envs whose ``done`` fired are reset on the device right after the step that reports it
        #### (BaseAviary.reset per environment, BaseAviary.py:406-424), with no host round trip
        self.AUTO_RESET = bool(auto_reset)
        self.CLIENT = -1  # no PyBullet client
        self.DRONE_IDS = np.arange(1, num_drones + 1)
        self._n_u = [d.INDI_ACTUATOR_NR for d in self.drones]
        self.action_space = self._actionSpace()
        self.observation_space = self._observationSpace()
        self._housekeeping()
        self._updateAndStoreKinematicInformation()

    # ------------------------------------------------------------------ gym API
    def reset(self):
        """BaseAviary.reset (BaseAviary.py:406-424)."""
        self._housekeeping()
        self._updateAndStoreKinematicInformation()
        return self._computeObs()

    def step(self, action):
        """BaseAviary.step (BaseAviary.py:428-555): clip, AGGR_PHY_STEPS substeps with the same action,
        refresh the state cache, return ``obs, reward, done, info``."""
        self._advance(action)
        self._updateAndStoreKinematicInformation()
        obs = self._computeObs()
        reward = self._computeReward()
        done = self._computeDone()
        info = self._computeInfo()
        self.step_counter = self._core.step_counter  # += AGGR_PHY_STEPS (:554)
        if self.AUTO_RESET:
            self.reset_envs(self._done_dev)  # the returned obs / done describe the finished episode's last step
        return obs, reward, done, info

    def reset_envs(self, mask):
        """``reset()`` for the environments with ``mask[e] != 0`` only ([NUM_ENVS] bool / uint8, device tensor or array),
        on the device and without a host synchronisation; the others keep flying."""
        self._core.reset_envs(mask, self._init_dev[0], rpy0=self._init_dev[1], vel0=self._init_dev[2])

    def _advance(self, action):
        """``_preprocessAction`` + the AGGR_PHY_STEPS substeps (BaseAviary.py:507-545), one kernel launch.
        CtrlAviary: the action is the PWM command, clipped on the device (CtrlAviary.py:236-263)."""
        self._core.physics_step(self._pack_action(action))

    def render(self, mode="human", close=False):
        """BaseAviary.render (BaseAviary.py:559-620): textual, first env only."""
        if self.first_render_call and not self.GUI:
            print("[WARNING] BaseAviary.render() is implemented as text-only, re-initialize the environment using "
                  "Aviary(gui=True) to use PyBullet's graphical interface")
            self.first_render_call = False
        it = self.step_counter
        print("\n[INFO] BaseAviary.render() ——— it {:04d}".format(it),
              "——— wall-clock time {:.1f}s,".format(time.time() - self.RESET_TIME),
              "simulation time {:.1f}s@{:d}Hz ({:.2f}x)".format(it * self.TIMESTEP, self.SIM_FREQ,
                                                                  (it * self.TIMESTEP) / max(time.time() - self.RESET_TIME, 1e-9)))
        pos, vel, rpy, ang = (np.asarray(x).reshape(-1, self.NUM_DRONES, 3)[0] for x in (self.pos, self.vel, self.rpy, self.ang_v))
        for i in range(self.NUM_DRONES):
            print("[INFO] BaseAviary.render() ——— drone {:d}".format(i),
                  "——— x {:+06.2f}, y {:+06.2f}, z {:+06.2f}".format(*pos[i]),
                  "——— velocity {:+06.2f}, {:+06.2f}, {:+06.2f}".format(*vel[i]),
                  "——— roll {:+06.2f}, pitch {:+06.2f}, yaw {:+06.2f}".format(*(rpy[i] * self.RAD2DEG)),
                  "——— angular velocity {:+06.4f}, {:+06.4f}, {:+06.4f} ——— ".format(*ang[i]))

    def close(self):
        self._core.close()

    def getPyBulletClient(self):
        return self.CLIENT

    def getDroneIds(self):
        return self.DRONE_IDS

    # ------------------------------------------------------------------ internals
    def _housekeeping(self):
        """BaseAviary._housekeeping (BaseAviary.py:640-714) on the device."""
        self.RESET_TIME = time.time()
        self.step_counter = 0
        self.first_render_call = True
        E, D = self.NUM_ENVS, self.NUM_DRONES
        bc = lambda a: np.broadcast_to(np.asarray(a, dtype=float), (E, D, 3))  # noqa: E731
        vel0 = None
        if self.INIT_VELS is not None:
            vel0 = np.zeros((D, 3))
            for i in range(D):
                if self.INIT_VELS[i] is not None:
                    vel0[i] = self.INIT_VELS[i]
            vel0 = bc(vel0)
        self._core.reset(bc(self.INIT_XYZS), rpy0=bc(self.INIT_RPYS), vel0=vel0)
        self.last_clipped_action = {str(i): np.zeros(self._n_u[i]) for i in range(D)}
        import torch

        dev = self._core.device  # the initial poses stay resident for masked resets (reset_envs / auto_reset)
        f = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev).reshape(E * D, 3)  # noqa: E731
        self._init_dev = (f(bc(self.INIT_XYZS)), f(bc(self.INIT_RPYS)), f(vel0))

    def _pack_action(self, action, width: int = 6):
        """dict {str(i): [<= width]} (reference form, one env) or array/tensor [E, D, <= width] -> device [N, width]."""
        import torch

        E, D = self.NUM_ENVS, self.NUM_DRONES
        dev = self._core.device
        if isinstance(action, dict):
            a = np.zeros((D, width), dtype=np.float32)
            for k, v in action.items():
                v = np.asarray(v, dtype=np.float32).reshape(-1)
                a[int(k), : v.shape[0]] = v
            t = torch.from_numpy(a).to(dev)
            if E > 1:
                t = t.unsqueeze(0).expand(E, D, width)
            return t.reshape(E * D, width).contiguous()
        t = torch.as_tensor(action, dtype=torch.float32, device=dev)
        if t.shape[-1] < width:
            t = torch.nn.functional.pad(t, (0, width - t.shape[-1]))
        return t.reshape(E * D, width).contiguous()

    def _updateAndStoreKinematicInformation(self):
        """State cache ``pos quat rpy vel ang_v`` (BaseAviary.py:718-732) + clipped action + adjacency,
        produced by one observation kernel; kept on the device, mirrored to numpy lazily."""
        st, nb, dn, rw = self._core.get_obs(state=True, neighbors=True, done=True, reward=True)
        E, D = self.NUM_ENVS, self.NUM_DRONES
        self._state_dev = st.view(E, D, 22)
        self._neigh_dev = nb.view(E, D)
        self._done_dev, self._reward_dev = dn, rw
        self._host = None

    def _host_state(self):
        if self._host is None:
            self._host = (self._state_dev.cpu().numpy().astype(np.float64), self._neigh_dev.cpu().numpy())
        return self._host

    def _sq(self, a):
        return a[0] if self.NUM_ENVS == 1 else a

    @property
    def pos(self):
        return self._sq(self._host_state()[0][..., 0:3])

    @property
    def quat(self):
        return self._sq(self._host_state()[0][..., 3:7])

    @property
    def rpy(self):
        return self._sq(self._host_state()[0][..., 7:10])

    @property
    def vel(self):
        return self._sq(self._host_state()[0][..., 10:13])

    @property
    def ang_v(self):
        return self._sq(self._host_state()[0][..., 13:16])

    def _getDroneStateVector(self, nth_drone, env: int = 0):
        """(16 + n_u,) state of one drone (BaseAviary.py:764-790)."""
        return self._host_state()[0][env, nth_drone, : 16 + self._n_u[nth_drone]].copy()

    def _getAdjacencyMatrix(self, env: int = 0):
        """(NUM_DRONES, NUM_DRONES) 0/1 matrix (BaseAviary.py:901-921) unpacked from the device bitmask."""
        bits = self._host_state()[1][env].astype(np.uint32)
        D = self.NUM_DRONES
        return ((bits[:, None] >> np.arange(D, dtype=np.uint32)[None, :]) & 1).astype(float)

    # ---- batched (device-resident) views for RL-style consumers
    def state_tensor(self):
        """[E, D, 22] float32 device tensor (16 + n_u used per drone, zero padded)."""
        return self._state_dev

    def neighbors_tensor(self):
        """[E, D] int32 device tensor: bit j of entry (e, i) = adjacency[i, j] of env e."""
        return self._neigh_dev

    def _actionSpace(self):
        raise NotImplementedError

    def _observationSpace(self):
        raise NotImplementedError

    def _computeObs(self):
        raise NotImplementedError

    def _computeReward(self):
        raise NotImplementedError

    def _computeDone(self):
        raise NotImplementedError

    def _computeInfo(self):
        raise NotImplementedError
