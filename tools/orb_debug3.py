import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccm_slam_b200 import api
from ccm_slam_b200.frontend import ORBextractor
from ccm_slam_b200.synth_images import make_image
api.init(0)
img = make_image(0)
ex = ORBextractor()
api.lib().ccm_orb_debug_dump(ex._h, 0, None, 0)
kps, desc = ex(img)
out = np.zeros(16384, np.int32)
api.lib().ccm_orb_debug_dump(ex._h, 0, out.ctypes.data_as(C.c_void_p), 16384)
n = 36 * 38
tile = out[:n].reshape(38, 36); score = out[n:2 * n].reshape(38, 36); flag = out[2 * n:3 * n].reshape(38, 36)
print("n_ini,tw,th", out[3 * n:3 * n + 3])
print("tile equals image ROI:", np.array_equal(tile, img[16:54, 16:52].astype(np.int32)))
if not np.array_equal(tile, img[16:54, 16:52].astype(np.int32)):
    print(tile[:4, :12]); print(img[16:20, 16:28])
np.set_printoptions(linewidth=250)
print("score rows 3..9:"); print(score[3:10])
print("flag nonzero:", list(zip(*np.nonzero(flag)))[:12])
