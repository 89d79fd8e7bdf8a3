import sys; sys.path.insert(0,'.')
import numpy as np, cv2, torch
import b200mosaic.ops as ops
from oracle import orb as oorb
fr=np.load('tests/golden/clip01_frames.npz')['frames']
g=cv2.cvtColor(fr[0],cv2.COLOR_BGR2GRAY)
img,sc=ops.orb_debug_level(torch.from_numpy(g).cuda(), 0)
rs=oorb.fast_score_map(g)
ys,xs=np.nonzero(sc!=rs)
print('n',len(ys),'gpu nonzero',np.count_nonzero(sc),'ref nonzero',np.count_nonzero(rs))
print('gpu==0 where ref>0:',np.count_nonzero((sc==0)&(rs>0)),' gpu>0 where ref==0:',np.count_nonzero((sc>0)&(rs==0)), 'both>0 differ', np.count_nonzero((sc>0)&(rs>0)&(sc!=rs)))
for y,x in list(zip(ys,xs))[:6]:
    ring=[int(g[y+dy,x+dx]) for dx,dy in oorb.RING]
    print((y,x),'gpu',sc[y,x],'ref',rs[y,x],'c',g[y,x],'d',[int(g[y,x])-r for r in ring])
