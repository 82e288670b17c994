"""``sys.modules`` shims so the UNMODIFIED reference controllers import without pybullet/gym.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Only usable where the reference checkout
exists (this container: ``/root/reference``); never on the GPU box, never from the product.

The reference controller modules import ``pybullet`` only for three math functions
(INDIControl.py:225,301,388,428 ; INDIControl_6DOF.py:334,419,538,548,567) and import
``BaseAviary`` (-> ``gym``, ``pybullet_data``, ``PIL``) only for unused names
(INDIControl.py:18).  The shims provide exactly that much.
"""
import os
import sys
import types

from . import pyb_math

REFERENCE_ROOT = os.environ.get("DRONESIM_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dronesim", "control"))


def install():
    """Install the shims and put the reference on ``sys.path``.  Idempotent."""
    if not reference_available():
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_ROOT)
    if "pybullet" not in sys.modules:
        pb = types.ModuleType("pybullet")
        pb.getEulerFromQuaternion = pyb_math.getEulerFromQuaternion
        pb.getQuaternionFromEuler = pyb_math.getQuaternionFromEuler
        pb.getMatrixFromQuaternion = pyb_math.getMatrixFromQuaternion
        pb.__shim__ = True
        sys.modules["pybullet"] = pb
    if "pybullet_data" not in sys.modules:
        sys.modules["pybullet_data"] = types.ModuleType("pybullet_data")
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env(object):
            pass

        gym.Env = Env
        spaces = types.ModuleType("gym.spaces")
        for name in ("Box", "Dict", "MultiBinary", "Discrete"):
            setattr(spaces, name, type(name, (), {"__init__": lambda self, *a, **k: None}))
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    try:
        import PIL  # noqa: F401
    except Exception:
        pil = types.ModuleType("PIL")
        img = types.ModuleType("PIL.Image")
        pil.Image = img
        sys.modules["PIL"] = pil
        sys.modules["PIL.Image"] = img
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def quad_controller(drone_model: str):
    """Reference ``dronesim.control.INDIControl.INDIControl`` instance (quad / 4-output law)."""
    install()
    from dronesim.control.INDIControl import INDIControl

    return INDIControl(drone_model=drone_model)


def hexa_controller(drone_model: str):
    """Reference ``dronesim.control.INDIControl_6DOF.INDIControl`` instance (6-DOF law)."""
    install()
    import contextlib
    import io

    from dronesim.control.INDIControl_6DOF import INDIControl

    with contextlib.redirect_stdout(io.StringIO()):  # it prints a banner (INDIControl_6DOF.py:177)
        return INDIControl(drone_model=drone_model)


def wls_alloc():
    install()
    from dronesim.control.wls_alloc import wls_alloc as f

    return f
