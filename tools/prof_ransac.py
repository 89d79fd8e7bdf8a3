import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, ctypes as C
from b200mosaic import _lib
lib=_lib.load()
g=np.load(str(pathlib.Path(__file__).resolve().parent.parent / 'tests/golden/clip01_orb.npz'))
mm=g['matches1']; src=g['kp1'][mm[:,0].astype(int),:2].astype(np.float32); dst=g['kp0'][mm[:,1].astype(int),:2].astype(np.float32)
def run(src,dst):
    H=np.zeros(9); cyc=np.zeros(8,np.int64); lm=C.c_int(0); js=C.c_int(0)
    for _ in range(2):
        _lib.check(lib.bm_ransac_profile(src.ctypes.data_as(C.c_void_p),dst.ctypes.data_as(C.c_void_p),len(src),2.0,2000,0.995,H.ctypes.data_as(C.POINTER(C.c_double)),cyc.ctypes.data_as(C.c_void_p),C.byref(lm),C.byref(js)))
    print('n',len(src),'cycles: subsets %d hyp %d sel %d refit-sums %d DLT %d LM %d total %d param %d'%tuple(cyc[:8]),'lm_iters',lm.value,'of which by eigen-decomposition',(js.value>>8)&255,'(sweeps of the last one %d)'%(js.value>>16),'DLT inverse-iteration steps',js.value&255)
run(src,dst)
rng=np.random.default_rng(3); n=500
s2=(rng.random((n,2))*[1920,1080]).astype(np.float32); Ht=np.array([[1.0,0.002,3],[-0.002,1.0,-11],[1e-7,2e-7,1]])
p=np.c_[s2,np.ones(n)]@Ht.T; d2=(p[:,:2]/p[:,2:]+rng.normal(0,0.3,(n,2))).astype(np.float32)
run(s2,d2)
# a consensus set in one corner of the frame (ill-conditioned): some LM iterations take the eigen-decomposition route (bits 8-15 of "jacobi sweeps")
s3=(rng.random((300,2))*[154,60]+[700,0]).astype(np.float32); p=np.c_[s3,np.ones(300)]@Ht.T; d3=(p[:,:2]/p[:,2:]+rng.normal(0,0.3,(300,2))).astype(np.float32)
run(s3,d3)
lib.bm_debug_lm_force_eig(1); run(s2,d2); lib.bm_debug_lm_force_eig(0)
