"""regt_b200 -- host-side mirror of the RegT-GCN hot path over libregt_b200.so (sm_100a)."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
