"""Debug helper (GPU box): run the full clip, find the frames whose relative homography differs from the reference-run golden,
and dump their matched point sets to gpurun_out/diverge_<det>.npz for offline comparison with cv2.findHomography."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import b200mosaic  # noqa: E402
from test_clip_gpu import _decode, _reproj  # noqa: E402


def main(det):
    gd = ROOT / "tests" / "golden"
    g = np.load(gd / f"clip01_full_{det}.npz")
    frames = _decode(gd, g)
    vm = b200mosaic.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False)
    out = {}
    bad = []
    for t in range(1, len(frames)):
        kp_prev = np.array([k.pt for k in vm.kp_prev], np.float32)
        vm.process_frame(frames[t], t)
        info = vm.last_info
        Hrel = np.array(info.H_rel).reshape(3, 3)
        e = _reproj(Hrel, g["H_rel"][t])
        if e > 1e-3:
            bad.append((t, e))
            kp_cur = np.array([k.pt for k in vm.kp_prev], np.float32)      # the frame was accepted: cur became prev
            m = np.array([[mm.queryIdx, mm.trainIdx, mm.distance] for mm in vm.matches])
            out[f"src_{t}"] = kp_cur[m[:, 0].astype(int)]
            out[f"dst_{t}"] = kp_prev[m[:, 1].astype(int)]
            out[f"H_{t}"] = Hrel
            out[f"it_{t}"] = np.array([info.ransac_iters, info.n_inliers, info.n_matches])
    print(det, "diverging frames:", [(t, round(e, 4)) for t, e in bad])
    out["bad"] = np.array([t for t, _ in bad])
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    np.savez_compressed(ROOT / "gpurun_out" / f"diverge_{det}.npz", **out)


if __name__ == "__main__":
    for d in sys.argv[1:] or ["orb", "sift"]:
        main(d)
