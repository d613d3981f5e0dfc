"""TEST / BENCH INFRASTRUCTURE ONLY — install the UNMODIFIED reference into oracle/_ref/.

    python oracle/build_ref.py

Runs in the build container (needs /root/reference).  ``pip install --no-deps --target oracle/_ref``
from a scratch copy (the mount is read-only and setuptools writes egg-info next to setup.py).
``oracle/_ref/`` is git-ignored (no reference source enters the history) but travels to the GPU box
with the snapshot, like the built ``.so``, so that ``bench.py --impl reference`` and the
``cpu_baseline`` leg can time the reference's own ``lightprop2d.Beam2D`` there
(``cpu_baseline.kind = "reference"``).  Import it with ``oracle.ref_shim.load()``: the package imports
``cupy`` at module scope, the shim provides the stub.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
DEST = os.path.join(HERE, "_ref")


def build(verbose=True):
    if not os.path.isdir(os.path.join(REF_ROOT, "lightprop2d")):
        return False
    if os.path.exists(os.path.join(DEST, "lightprop2d", "beam2d.py")):
        return True
    tmp = tempfile.mkdtemp(prefix="lp2d_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_ROOT, src, ignore=shutil.ignore_patterns("tests", "docs", ".git*"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DEST, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"pip install of the reference failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[oracle] reference installed into {DEST}", file=sys.stderr)
        return True
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    ok = build()
    print("installed" if ok else "reference mount absent: nothing to do")
