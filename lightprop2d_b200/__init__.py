"""lightprop2d_b200 — B200 (sm_100a) backend for the lightprop2d angular-spectrum hot path.

Same public names as ``lightprop2d/__init__.py:7-8``: ``Beam2D``, the length units
and the profile generators.
"""
from .beam2d import Beam2D, nm, um, mm, cm, m
from .profiles import *
from .profiles import __all__ as _profiles_all
from . import device_profiles          # the same generators evaluated on the GPU (lp2d_generate)

__all__ = ["Beam2D", "nm", "um", "mm", "cm", "m", *_profiles_all]
__version__ = "0.1.0"
