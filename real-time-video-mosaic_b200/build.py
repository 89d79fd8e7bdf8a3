"""Build libb200mosaic.so in-tree with nvcc for sm_100a (no torch involved; the library is a plain C ABI).

    python real-time-video-mosaic_b200/build.py [--force] [--verbose]

Objects go to csrc/_obj/, the library and its source-digest stamp to lib/ (git-ignored, but the directory travels with gpurun).
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
STAMP = LIBDIR / "stamp.txt"
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


def _up_to_date(dig: str) -> bool:
    return LIB.exists() and STAMP.exists() and STAMP.read_text() == dig


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu and link lib/libb200mosaic.so unless lib/stamp.txt already carries the digest of the sources + flags.
    The stamp lives NEXT TO the library (lib/ travels to the GPU box with the tree, csrc/_obj/ does not): a box that received an
    up-to-date library loads it as is instead of recompiling in every process.  Builders are serialised by a file lock, so the N ranks of
    a torchrun launch that all find a stale library compile once, not N times into the same object files."""
    import fcntl
    OBJ.mkdir(exist_ok=True)
    LIBDIR.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [HERE.parent / "include" / "b200mosaic.h"]
    dig = _digest(srcs + hdrs)
    if not force and _up_to_date(dig):
        return LIB
    with open(LIBDIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(dig):          # another process built it while this one waited for the lock
                return LIB
            logs = []
            with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
                objs = []
                for obj, log in ex.map(lambda s: _compile(s, verbose), srcs):
                    objs.append(obj)
                    logs.append(log)
            (OBJ / "ptxas.log").write_text("\n".join(logs))
            tmp = LIBDIR / (LIB.name + f".tmp{os.getpid()}")
            cmd = [NVCC, "-shared", "-o", str(tmp), *map(str, objs), "-cudart", "static", "-lcuda"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                tmp.unlink(missing_ok=True)
                raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
            if STAMP.exists():
                STAMP.unlink()                          # never a stamp that vouches for a library it does not describe
            os.replace(tmp, LIB)                        # atomic: a process that already mapped the old file keeps its inode
            STAMP.write_text(dig)
            if verbose:
                print("\n".join(logs))
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
