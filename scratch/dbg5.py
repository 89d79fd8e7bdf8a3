import sys; sys.path.insert(0,'.')
import numpy as np, cv2, torch
import b200mosaic.ops as ops
from oracle import orb as oorb
from b200mosaic.synth import DroneSweep
g=cv2.cvtColor(DroneSweep(1920,1080,seed=21,ground_size=2048).next(),cv2.COLOR_BGR2GRAY)
kp,des=ops.orb_detect_and_compute(torch.from_numpy(g).cuda())
kc,dc=oorb.cv_detect_and_compute(g)
a,ad=oorb.canon(kp.astype(np.float64),des); b,bd=oorb.canon(kc,dc)
bad=np.nonzero((a!=b).any(axis=1)|(ad!=bd).any(axis=1))[0]
for i in bad:
    print(a[i],b[i], 'desc bits differ', int(np.unpackbits(ad[i]^bd[i]).sum()))
    print(repr(np.float32(a[i][3])), repr(np.float32(b[i][3])))
