import sys; sys.path.insert(0,'.')
import numpy as np, cv2, torch
cv2.ipp.setUseIPP(False)
import b200mosaic, b200mosaic.ops as ops
from oracle import orb as oorb, ransac as ors, matching as omt
from oracle.mosaic_ref import RefMosaic
from b200mosaic.synth import DroneSweep
g=cv2.cvtColor(DroneSweep(1920,1080,seed=21,ground_size=2048).next(),cv2.COLOR_BGR2GRAY)
try:
    kp,des=ops.orb_detect_and_compute(torch.from_numpy(g).cuda())
    kc,dc=oorb.cv_detect_and_compute(g)
    print('1080p', kp.shape, kc.shape)
    A={(r[5],r[1],r[0]):(r,d) for r,d in zip(kp.astype(np.float64),des)}
    B={(r[5],r[1],r[0]):(r,d) for r,d in zip(kc,dc)}
    print('extra',len(set(A)-set(B)),'missing',len(set(B)-set(A)))
    for l in range(8): print(l, sum(1 for k in A if k[0]==l), sum(1 for k in B if k[0]==l))
    nb=0
    for k in set(A)&set(B):
        if not np.array_equal(A[k][0],B[k][0]) or not np.array_equal(A[k][1],B[k][1]): nb+=1
    print('common differing',nb)
except Exception as e:
    print('ERR',e)
# e2e per-step
fr=np.load('tests/golden/clip01_frames.npz')['frames']
gd=np.load('tests/golden/clip01_orb.npz')
def reproj(Ha,Hb,w=427,h=240):
    ys,xs=np.mgrid[0:h:16,0:w:16]; p=np.stack([xs.ravel(),ys.ravel(),np.ones(xs.size)])
    a=Ha@p; b=Hb@p; return np.abs(a[:2]/a[2]-b[:2]/b[2]).max()
vm=b200mosaic.VideMosaic(fr[0],detector_type='orb',show_intermediate=False,visualize=False)
ref=RefMosaic(fr[0],detector_type='orb')
for t in range(1,5):
    vm.process_frame(fr[t],t); ref.process_frame(fr[t],t)
    Hr=np.array(vm.last_info.H_rel).reshape(3,3)
    print(t,'n_matches',vm.last_info.n_matches,len(ref.matches),'iters',vm.last_info.ransac_iters,'inl',vm.last_info.n_inliers,'Hrel reproj',reproj(Hr,ref.H_rel),'Habs reproj',reproj(vm.H,ref.H))
    # run oracle ransac on GPU's own matches/points
    m=vm.matches; kpc=vm._fetch_kp(0)[0]
