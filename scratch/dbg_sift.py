import sys; sys.path.insert(0,'.')
import numpy as np, cv2, torch
import b200mosaic.ops as ops
from oracle import sift as osift
fr=np.load('tests/golden/clip01_frames.npz')['frames']
g=cv2.cvtColor(fr[0],cv2.COLOR_BGR2GRAY)
gp,dp=osift.build_pyramids(g); dev=torch.from_numpy(g).cuda()
_,noct=ops.sift_debug_level(dev,0,0)
for o in range(noct):
    ds=[]
    for l in range(6):
        img,_=ops.sift_debug_level(dev,o,l); ds.append(float(np.abs(img-gp[o][l]).max()))
    print(o, gp[o][0].shape, ['%.2e'%d for d in ds], 'frac exact l1', float(np.mean(ops.sift_debug_level(dev,o,1)[0]==gp[o][1])))
kp,des=ops.sift_detect_and_compute(dev); kc,dc=osift.cv_detect_and_compute(g)
pairs=osift.match_keypoints(kc,kp.astype(np.float64))
print('kp',len(kp),len(kc),'matched',len(pairs))
ia,ib=pairs[:,0],pairs[:,1]
dang=np.abs(((kp[ib,3]-kc[ia,3])+180)%360-180); print('angle <0.01:',np.mean(dang<0.01),' <0.5:',np.mean(dang<0.5), 'max',dang.max())
l2=np.linalg.norm(des[ib].astype(float)-dc[ia].astype(float),axis=1); print('desc exact',np.mean(l2==0),'median',np.median(l2),'p99',np.percentile(l2,99),'max',l2.max())
print('pt maxdiff',np.abs(kp[ib,:2]-kc[ia,:2]).max())
