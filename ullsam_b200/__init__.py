"""ullsam_b200 — B200-native (sm_100a) SAM segmentation branch of uLLSAM behind the reference's module API."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
