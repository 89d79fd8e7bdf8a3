import sys; sys.path.insert(0,'.')
import numpy as np, cv2, torch
import b200mosaic.ops as ops
from oracle import orb as oorb
fr=np.load('tests/golden/clip01_frames.npz')['frames']
g=cv2.cvtColor(fr[0],cv2.COLOR_BGR2GRAY)
lev=oorb.build_pyramid(g)
for l in range(8):
    img,sc=ops.orb_debug_level(torch.from_numpy(g).cuda(), l)
    ref=lev[l]; rs=oorb.fast_score_map(ref)
    print(l, img.shape, ref.shape, 'img mism', np.count_nonzero(img!=ref) if img.shape==ref.shape else 'shape', 'score mism', np.count_nonzero(sc!=rs) if sc.shape==rs.shape else 'shape')
    if img.shape==ref.shape and np.count_nonzero(img!=ref):
        ys,xs=np.nonzero(img!=ref); print('  first', list(zip(ys[:5],xs[:5])), img[ys[:5],xs[:5]], ref[ys[:5],xs[:5]], 'rows', np.unique(ys)[:10], 'cols', np.unique(xs)[:10])
