"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck): ORB and SIFT pipelines with
pipelined calls, the stage entry points, finalisation, canvas tiles."""
import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200mosaic, b200mosaic.ops as ops
from b200mosaic.synth import DroneSweep

fr = DroneSweep(427, 240, seed=5, ground_size=1024, max_step=7.0).frames(5)
for det in ("orb", "sift"):
    vm = b200mosaic.VideMosaic(fr[0], detector_type=det, show_intermediate=False, visualize=False)
    for t in range(1, 5):
        vm.process_frame(fr[t], t, next_frame=fr[t + 1] if t + 1 < 5 else None)
    print(det, vm.last_info.status, vm.last_info.n_matches, vm.output_img.shape, vm.finalize().shape, len(vm.matches))
    vm.close()
m = (np.random.default_rng(0).random((97, 131)) > 0.01).astype(np.uint8) * 255
print(float(ops.distance_transform(torch.from_numpy(m).cuda()).max()))
q = np.random.default_rng(1).integers(0, 200, (130, 128)).astype(np.float32)
print(len(ops.match_l2_ratio(q, q[::-1].copy(), 1.5)))
print("done")
