"""nvp_b200 — B200-native implementation of NVP's per-coordinate encode path."""
from . import _lib  # noqa: F401
from .modules import NVP  # noqa: F401

__all__ = ["NVP"]
