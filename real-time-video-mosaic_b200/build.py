"""Build libb200mosaic.so in-tree with nvcc for sm_100a (no torch involved; the library is a plain C ABI).

    python real-time-video-mosaic_b200/build.py [--force] [--verbose]

Objects go to csrc/_obj/, the library to lib/libb200mosaic.so (git-ignored, but it travels with gpurun).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = CSRC / "_obj"
LIBDIR = HERE / "lib"
LIB = LIBDIR / "libb200mosaic.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v", "--fmad=false"] + os.environ.get("BM_EXTRA_NVCC_FLAGS", "").split()
# --fmad=false: the parity-critical kernels spell every FMA explicitly (__fmaf_rn); the compiler must not contract the
# rest (OpenCV's scalar code paths are not contracted either).


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src: Path, verbose: bool) -> tuple[Path, str]:
    obj = OBJ / (src.stem + ".o")
    cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{log}")
    return obj, log


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    LIBDIR.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [HERE.parent / "include" / "b200mosaic.h"]
    stamp = OBJ / "stamp.txt"
    dig = _digest(srcs + hdrs)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    logs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = []
        for obj, log in ex.map(lambda s: _compile(s, verbose), srcs):
            objs.append(obj)
            logs.append(log)
    (OBJ / "ptxas.log").write_text("\n".join(logs))
    cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static", "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    stamp.write_text(dig)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
