"""B200-native batched trajectory evaluation behind the reference's Trajectory interface."""
from . import abi  # noqa: F401

__all__ = ["abi"]
