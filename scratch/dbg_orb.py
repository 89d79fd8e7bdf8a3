import sys; sys.path.insert(0,'.')
import numpy as np, cv2, torch
cv2.ipp.setUseIPP(False)
import b200mosaic.ops as ops
from oracle import orb as oorb
fr=np.load('tests/golden/clip01_frames.npz')['frames']
g=cv2.cvtColor(fr[0],cv2.COLOR_BGR2GRAY)
kp,des=ops.orb_detect_and_compute(torch.from_numpy(g).cuda())
kc,dc=oorb.cv_detect_and_compute(g)
A={(r[5],r[1],r[0]):(r,d) for r,d in zip(kp.astype(np.float64),des)}
B={(r[5],r[1],r[0]):(r,d) for r,d in zip(kc,dc)}
print(len(A),len(B))
for k in sorted(set(A)-set(B)): print('extra',k,A[k][0])
for k in sorted(set(B)-set(A)): print('missing',k,B[k][0])
nb=0
for k in set(A)&set(B):
    if not np.array_equal(A[k][0],B[k][0]): nb+=1; print('diff',A[k][0],B[k][0]) if nb<5 else None
    
print('nbad',nb, 'desc bad', sum(not np.array_equal(A[k][1],B[k][1]) for k in set(A)&set(B)))
# per level counts
for l in range(8): print(l, sum(1 for k in A if k[0]==l), sum(1 for k in B if k[0]==l))
