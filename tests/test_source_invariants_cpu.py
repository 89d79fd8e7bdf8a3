"""Static guards over csrc/ that need no GPU."""
import re
from pathlib import Path

CSRC = Path(__file__).resolve().parent.parent / "real-time-video-mosaic_b200" / "csrc"


def test_every_pdl_launched_kernel_waits_for_its_predecessor():
    """a kernel launched with programmatic stream serialization may start while its predecessor still runs: it must execute
    griddepcontrol.wait (BM_PDL_WAIT) before it touches global memory (common.cuh)"""
    src = {p.name: p.read_text() for p in CSRC.glob("*.cu")}
    launched = set()
    for text in src.values():
        launched |= set(re.findall(r"bm_launch_pdl\((k_[a-z0-9_]+)", text))
    assert len(launched) >= 10
    for k in sorted(launched):
        bodies = [m for text in src.values() for m in re.finditer(r"__global__[^{;]*\b" + k + r"\([^{;]*\)\s*\{([^\n]*\n){1,4}", text)]
        assert bodies, k
        for m in bodies:
            assert "BM_PDL_WAIT()" in m.group(0), f"{k} is launched with bm_launch_pdl but does not start with BM_PDL_WAIT()"

