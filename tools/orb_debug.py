import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccm_slam_b200 import api
from ccm_slam_b200.frontend import ORBextractor
from ccm_slam_b200.synth_images import make_image
from oracle import pyoracle as po
api.init(0)
img = make_image(0)
ex = ORBextractor()
kps, desc = ex(img)
xys = np.zeros((60000, 3), np.float32); lv = np.zeros(60000, np.int32); n = C.c_int32()
api.lib().ccm_orb_debug_candidates(ex._h, xys.ctypes.data_as(C.c_void_p), lv.ctypes.data_as(C.c_void_p), 60000, C.byref(n))
print("gpu total candidates", n.value)
cfg = po.orb_cfg()
for l in range(8):
    ref = po.orb_level_candidates(img, cfg, l)
    got = xys[:n.value][lv[:n.value] == l]
    same = len(ref) == len(got) and np.array_equal(ref, got)
    print("level", l, "ref", len(ref), "gpu", len(got), "equal", same)
    if not same:
        m = min(len(ref), len(got))
        d = np.nonzero((ref[:m] != got[:m]).any(axis=1))[0]
        print("  first diff idx", d[:5], ref[d[:3]], got[d[:3]])
rk, rd = po.orb_extract(img)
print("kps gpu", len(kps), "ref", len(rk))
for l in range(8):
    a = kps[kps["octave"] == l]; b = rk[rk["octave"] == l]
    print(" lvl", l, len(a), len(b), "same set", set(zip(a["x"], a["y"])) == set(zip(b["x"], b["y"])), "same order", len(a) == len(b) and np.array_equal(a["x"], b["x"]))
cur = img
for l in range(1, 8):
    lv = ex.image_pyramid(l)
    cur = po.resize_linear_u8(cur, lv.shape[1], lv.shape[0])
    print("pyr level", l, lv.shape, "equal", np.array_equal(lv, cur), "ndiff", int((lv != cur).sum()))
