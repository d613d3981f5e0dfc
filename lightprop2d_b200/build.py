"""Build liblp2d.so (the sm_100a CUDA backend) in-tree with nvcc.

    python -m lightprop2d_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so lands next to this file so that it
travels with the source tree; it is git-ignored.
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "liblp2d.so")
UNITS = ["lp2d_api.cu", "lp2d_aux.cu", "lp2d_f32.cu", "lp2d_f64.cu", "lp2d_p2_f32.cu", "lp2d_p2_f64.cu", "lp2d_modes_tc.cu", "lp2d_xl.cu", "lp2d_generic.cu", "lp2d_gen.cu", "lp2d_colstep.cu"]
NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
# tuning builds only: extra -D definitions (LP2D_Q4_CLOCKS=1 phase clocks, LP2D_ABLATION=1 no-transform variants)
NVCC_FLAGS += os.environ.get("LP2D_BUILD_DEFS", "").split()


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: liblp2d.so cannot be built on this machine")
    return exe


def _source_digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(unit):
    out = os.path.join(OBJ, unit.replace(".cu", ".o"))
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, unit), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {unit}:\n{r.stdout}\n{r.stderr}")
    return out


def build(force=False, verbose=True):
    digest = _source_digest()
    stamp = LIB + ".stamp"
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    if verbose:
        print(f"[lp2d] compiling {len(UNITS)} units for sm_100a ...", file=sys.stderr)
    with cf.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(_compile, UNITS))
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print(f"[lp2d] built {LIB}", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
