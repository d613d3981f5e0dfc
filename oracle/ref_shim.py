"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference from /root/reference.

Source of the package, in this order: the read-only mount ``/root/reference`` (build container;
``oracle/make_golden.py`` generates ``tests/golden/*.npz`` from it), else ``oracle/_ref`` (the pip
install of the same unmodified package made by ``oracle/build_ref.py``; git-ignored, it travels to
the GPU box with the snapshot so that ``bench.py --impl reference`` / ``cpu_baseline`` can time the
reference's own ``Beam2D`` there).

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

REF_MOUNT = "/root/reference"
REF_INSTALL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def ref_root():
    for root in (REF_MOUNT, REF_INSTALL):
        if os.path.isfile(os.path.join(root, "lightprop2d", "beam2d.py")):
            return root
    return None


def available() -> bool:
    return ref_root() is not None


def mount_available() -> bool:
    return os.path.isfile(os.path.join(REF_MOUNT, "lightprop2d", "beam2d.py"))


def load():
    """Return the reference ``lightprop2d`` module (numpy path)."""
    REF_ROOT = ref_root()
    if REF_ROOT is None:
        raise RuntimeError("neither the reference mount /root/reference nor oracle/_ref is present")
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
