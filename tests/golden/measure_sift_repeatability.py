"""How repeatable is the UNMODIFIED reference's own SIFT run?  Runs /root/reference/main.py `VideMosaic(detector_type="sift")` over the
full clip a second time and compares its homographies with the first run stored in clip01_full_sift.npz (build container only):

    python tests/golden/measure_sift_repeatability.py        ->  tests/golden/clip01_sift_repeatability.json

cv2's SIFT is not bit-repeatable from call to call (orientation angles jitter by an ulp, a keypoint at the 0.8-of-maximum orientation
threshold comes or goes), which reshuffles `retainBest`'s output order, hence the cv::RNG subsets of findHomography.  The result is the
noise floor any SIFT trajectory comparison against the goldens has to be read against (tests/test_clip_gpu.py)."""
import sys, io, contextlib, numpy as np, cv2
sys.path.insert(0,'/root/repo/tests/golden'); sys.path.insert(0,'/root/repo/tests')
from make_golden import load_reference
cv2.ipp.setUseIPP(False)
ref=load_reference()
g=np.load('/root/repo/tests/golden/clip01_full_sift.npz')
cap=cv2.VideoCapture('/root/repo/tests/golden/clip01.mp4'); frames=[]
while True:
    ok,f=cap.read()
    if not ok: break
    frames.append(f)
C=np.array([[0,0,1],[853,0,1],[853,479,1],[0,479,1]],float).T
def rp(a,b):
    x=a@C; y=b@C; return float(np.abs(x[:2]/x[2]-y[:2]/y[2]).max())
vm=ref.VideMosaic(frames[0],detector_type='sift',show_intermediate=False,visualize=False)
rel=[];ab=[];nk=[]
for t in range(1,len(frames)):
    with contextlib.redirect_stdout(io.StringIO()): vm.process_frame(frames[t],t)
    rel.append(rp(vm.last_valid_H,g['H_rel'][t])); ab.append(rp(vm.H_old,g['H'][t])); nk.append(len(vm.kp_cur)-g['n_kp'][t])
rel=np.array(rel);ab=np.array(ab)
import json
json.dump({"rel_H_max_px": float(rel.max()), "frames_above_1e-3_px": int((rel>1e-3).sum()), "frames_above_0.5_px": int((rel>0.5).sum()), "frames": int(len(rel)), "abs_drift_max_px": float(ab.max()), "n_kp_diff_max": int(np.abs(nk).max()), "cv2": cv2.__version__}, open("/root/repo/tests/golden/clip01_sift_repeatability.json","w"), indent=1)
print("cv2 SIFT run 2 vs run 1 (goldens): rel H max %.4f px, >1e-3: %d, >0.5: %d; abs drift max %.3f px; n_kp diff max %d"%(rel.max(),(rel>1e-3).sum(),(rel>0.5).sum(),ab.max(),np.abs(nk).max()))
