"""Build libdis_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m depthinspace_b200.build [--force] [--verbose]

The k x k window kernels are instantiated once per window radius (photometric_inst.cu with
-DDIS_R=0..7); the objects are compiled in parallel and linked into
depthinspace_b200/lib/libdis_b200.so.  The .so is git-ignored but travels with the tree.
"""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "lib", "obj")
LIB_PATH = os.path.join(HERE, "lib", "libdis_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-I", CSRC, "-I", INCLUDE,
]
PLAIN_UNITS = ["api.cu", "lcn.cu", "misc.cu", "smooth.cu", "flow_warp.cu", "flow_consistency.cu", "conv3d_gather.cu", "resize.cu", "ext_misc.cu", "point_loss.cu"]
RADII = range(8)


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libdis_b200.so cannot be built")
    return exe


def _units():
    units = [(f, os.path.join(OBJ_DIR, f.replace(".cu", ".o")), []) for f in PLAIN_UNITS
             if os.path.exists(os.path.join(CSRC, f))]
    units += [("photometric_inst.cu", os.path.join(OBJ_DIR, f"photometric_r{r}.o"), [f"-DDIS_R={r}"]) for r in RADII]
    return units


def source_digest():
    h = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode() + b"\0" + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh():
    stamp = LIB_PATH + ".digest"
    return os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == source_digest()


def build(force=False, verbose=False):
    if not force and is_fresh():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(unit):
        src, obj, extra = unit
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src} {extra}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max(1, min(8, os.cpu_count() or 1))) as ex:
        objs = list(ex.map(compile_one, _units()))
    link = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(LIB_PATH + ".digest", "w") as f:
        f.write(source_digest())
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
