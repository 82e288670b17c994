"""Host-side utilities mirroring ``dronesim.utils`` where the hot path touches them (the Logger's data format)."""
from .Logger import Logger  # noqa: F401
