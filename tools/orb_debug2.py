import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccm_slam_b200 import api
from ccm_slam_b200.frontend import ORBextractor
from ccm_slam_b200.synth_images import make_image
api.init(0)
img = make_image(0)
ex = ORBextractor()
kps, desc = ex(img)
for cell in (0, 1, 24, 335, 336):
    g = np.zeros(8, np.int32); ent = np.zeros((4096, 4), np.uint16); n = C.c_int32()
    rc = api.lib().ccm_orb_debug_cell(ex._h, cell, g.ctypes.data_as(C.c_void_p), ent.ctypes.data_as(C.c_void_p), C.byref(n))
    print("cell", cell, "rc", rc, "geom(level,tile_cap,max_per_cell,x0,y0,x1,y1,out_off)", g.tolist(), "count", n.value)
    print("   entries", ent[:min(n.value, 8)].tolist())
