"""warp+blend chain only (no features): drives VideMosaic.warp with the sweep's true homographies. For ncu captures."""
import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, time
import b200mosaic
from b200mosaic.synth import DroneSweep
n=int(sys.argv[1]) if len(sys.argv)>1 else 8
sw=DroneSweep(1920,1080,seed=1234,ground_size=4096,max_step=12.0,max_travel=860)
fr=sw.frames(n+1)
vm=b200mosaic.VideMosaic(fr[0],detector_type='orb',show_intermediate=False,visualize=False)
H=vm.H_old.copy()
vm.timing(enable=True,reset=True)
for t in range(1,n+1):
    H=H@sw.D_true[t-1]
    vm.warp(fr[t],H)
ms,by,frs=vm.timing(reset=True)
print('chain ms/frame',ms/frs,'GB/s',by/1e9/(ms/1e3), 'frames', frs)
