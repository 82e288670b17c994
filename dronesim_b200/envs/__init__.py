"""Gym-style aviary facades over the CUDA core (drop-in for ``dronesim.envs``)."""
from .BaseAviary import BaseAviary, Physics  # noqa: F401
from .CtrlAviary import CtrlAviary  # noqa: F401
from .RPYTAviary import RPYTAviary  # noqa: F401
from .VelocityAviary import VelocityAviary  # noqa: F401
