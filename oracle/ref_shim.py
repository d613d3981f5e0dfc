"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference from /root/reference.

Works only in the build container (the reference mount does not exist on the GPU
box).  Used by ``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by
``tests/test_oracle_vs_reference.py`` (skipped when the mount is absent).

The reference cannot be imported as is without cupy (``NameError: cp`` at
beam2d.py:135, annotations are evaluated at def time), so a stub ``cupy`` module
is placed in ``sys.modules`` first; after import ``_using_cupy`` is forced False so
that every code path is the reference's own numpy path (beam2d.py:106-109, 174-175).
No reference file is modified or copied.
"""
import os
import sys
import types
import warnings

REF_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lightprop2d"))


def load():
    """Return the reference ``lightprop2d`` module (numpy path)."""
    if not available():
        raise RuntimeError("reference mount /root/reference is not present")
    import numpy as np
    sys.dont_write_bytecode = True            # the mount is read-only
    if "cupy" not in sys.modules:
        stub = types.ModuleType("cupy")
        stub.ndarray = type("ndarray", (), {})
        stub.get_array_module = lambda *a: np
        sys.modules["cupy"] = stub
    import scipy.interpolate as si
    if not hasattr(si, "interp2d"):           # removed in recent SciPy; only crop/lens_image use it
        si.interp2d = None
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import lightprop2d as ref
        import lightprop2d.beam2d as refb
    refb._using_cupy = False
    return ref
