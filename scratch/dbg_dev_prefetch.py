import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
import b200mosaic
from b200mosaic import _lib
from b200mosaic.synth import DroneSweep
lib=_lib.load()
sw=DroneSweep(1920,1080,seed=1234,ground_size=4096,max_step=12.0,max_travel=860)
fr=sw.frames(40)
dev=torch.from_numpy(np.stack(fr)).cuda(); fb=fr[0].nbytes; base=dev.data_ptr()
for mode in ("plain","next"):
    vm=b200mosaic.VideMosaic(fr[0],detector_type='orb',show_intermediate=False,visualize=False)
    for i in range(1,6): vm.process_frame_device(base+i*fb, base+(i+1)*fb if mode=="next" else None)
    vm.sync(); l0=lib.bm_kernel_launches(); t0=time.perf_counter(); st=[]
    for i in range(6,39): st.append(vm.process_frame_device(base+i*fb, base+(i+1)*fb if mode=="next" else None))
    vm.sync(); dt=time.perf_counter()-t0
    print(mode, 33/dt, 'fps', 'launches/frame', (lib.bm_kernel_launches()-l0)/33, set(st))
