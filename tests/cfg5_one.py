"""Diagnostic: ONE config-5-sized BA (7 x 5000) for an ncu launch list."""
import sys, os, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvus_b200 import synth
BA_KW = dict(rs=True, motion_reg=True, motion_weights=1e4)
s = synth.make_flight(nc=7, det_per_cam=5000, seed=7, rolling_shutter=True, distortion=True, motion_type='F', motion_weights=1e4, uncovered=0.0)[0]
with contextlib.redirect_stdout(io.StringIO()):
    r = s.BA(7, max_iter=9, **BA_KW)
print(r.stats['ms_total'], r.stats['launches'])
