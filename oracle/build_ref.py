"""Build the compiled checker pieces of the oracle into oracle/_ref/ (TEST INFRASTRUCTURE ONLY, git-ignored output).

    python oracle/build_ref.py

* oracle/_ref/libstlorder.so  <-  oracle/csrc/stl_order.cpp: the REAL libstdc++ std::nth_element / std::partition /
  std::sort that decide cv2's keypoint order (cv::KeyPointsFilter::retainBest), used to pin oracle/cvorder.py and the CUDA
  order emulation.  The reference itself is pure Python (no C sources to compile), so this is the only compiled piece.
"""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"


def build(force: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    src = HERE / "csrc" / "stl_order.cpp"
    lib = OUT / "libstlorder.so"
    if force or not lib.exists() or lib.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", str(src), "-o", str(lib)], check=True)
    return lib


if __name__ == "__main__":
    print(build(force=True))
