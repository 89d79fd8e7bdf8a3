import sys; sys.path.insert(0,'.')
import numpy as np, cv2, torch
import b200mosaic.ops as ops
from oracle import sift as osift
from b200mosaic.synth import DroneSweep
for size in [(640,360),(1920,1080)]:
    g=cv2.cvtColor(DroneSweep(size[0],size[1],seed=9,ground_size=2048).next(),cv2.COLOR_BGR2GRAY)
    kp,des=ops.sift_detect_and_compute(torch.from_numpy(g).cuda())
    kc,dc=osift.cv_detect_and_compute(g)
    print(size,len(kp),len(kc),'resp ours min/max',kp[:,4].min(),kp[:,4].max(),'cv',kc[:,4].min(),kc[:,4].max())
    print(' ours resp pct',np.percentile(kp[:,4],[0,10,50,90,100]),' cv',np.percentile(kc[:,4],[0,10,50,90,100]))
