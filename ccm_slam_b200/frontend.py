"""Host-side mirror of the reference's ORB front-end interface on top of the C ABI (ctypes).

  ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)  ~ cslam::ORBextractor  (I/ORBextractor.h:103-138)
  ORBmatcher(nnratio, checkOri).SearchByBoW / SearchForTriangulation   ~ cslam::ORBmatcher    (I/ORBmatcher.h:97-145)

Same names, argument meaning and outputs as the reference classes, on flat numpy arrays instead of cv::Mat / KeyFrame.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api
from .api import FeatureVectorC, KeyPointC, ORBConfigC, TriViewC, _chk, _p, lib

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"), ("octave", "i4")])


class ORBextractor:
    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, width=752, height=480, blur_2413=False):
        self.cfg = ORBConfigC(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, int(blur_2413))
        self.width, self.height, self.nlevels = width, height, nlevels
        self._h = C.c_void_p()
        _chk(lib().ccm_orb_create(C.byref(self.cfg), width, height, C.byref(self._h)))
        self.max_kp = nfeatures + 4 * nlevels + 64

    def __call__(self, image):
        """operator()(image, mask, keypoints, descriptors): returns (keypoints structured array, descriptors N x 32 u8)."""
        img = np.ascontiguousarray(image, np.uint8)
        assert img.shape == (self.height, self.width)
        kps = np.zeros(self.max_kp, KP_DTYPE); desc = np.zeros((self.max_kp, 32), np.uint8); n = C.c_int32()
        _chk(lib().ccm_orb_extract(self._h, _p(img), img.strides[0], _p(kps), self.max_kp, C.byref(n), _p(desc)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def image_pyramid(self, level):
        """mvImagePyramid[level]"""
        w = C.c_int32(); h = C.c_int32()
        _chk(lib().ccm_orb_get_level(self._h, level, None, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        _chk(lib().ccm_orb_get_level(self._h, level, _p(out), C.byref(w), C.byref(h)))
        return out

    def close(self):
        if self._h:
            lib().ccm_orb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FeatureVector:
    """DBoW2::FeatureVector flattened (nodes ascending; features per node in insertion order)."""

    def __init__(self, node_of_feature):
        node_of_feature = np.asarray(node_of_feature)
        order = np.argsort(node_of_feature, kind="stable")
        nodes, counts = np.unique(node_of_feature, return_counts=True)
        self.node_id = nodes.astype(np.uint32)
        self.node_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.feat = order.astype(np.uint32)

    def c(self):
        return FeatureVectorC(len(self.node_id), _p(self.node_id), _p(self.node_ptr), _p(self.feat))


class ORBmatcher:
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30

    def __init__(self, nnratio=0.6, checkOri=True):
        self.nnratio, self.checkOri = float(nnratio), bool(checkOri)

    @staticmethod
    def DescriptorDistance(a, b):
        return int(api.hamming_matrix(np.asarray(a).reshape(1, 32), np.asarray(b).reshape(1, 32))[0, 0])

    def SearchByBoW_KF_Frame(self, desc_kf, kf_has_mp, angle_kf, fv_kf, desc_f, angle_f, fv_f):
        desc_kf = np.ascontiguousarray(desc_kf, np.uint8); desc_f = np.ascontiguousarray(desc_f, np.uint8)
        has = np.ascontiguousarray(kf_has_mp, np.uint8); ak = np.ascontiguousarray(angle_kf, np.float32); af = np.ascontiguousarray(angle_f, np.float32)
        out = np.empty(desc_f.shape[0], np.int32); n = C.c_int32()
        fk, ff = fv_kf.c(), fv_f.c()
        _chk(lib().ccm_match_bow_kf_frame(_p(desc_kf), desc_kf.shape[0], _p(has), _p(ak), C.byref(fk), _p(desc_f), desc_f.shape[0], _p(af),
                                          C.byref(ff), C.c_float(self.nnratio), int(self.checkOri), _p(out), C.byref(n)))
        return out, n.value

    def SearchByBoW_KF_KF(self, d1, has1, a1, fv1, d2, has2, a2, fv2):
        d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
        has1 = np.ascontiguousarray(has1, np.uint8); has2 = np.ascontiguousarray(has2, np.uint8)
        a1 = np.ascontiguousarray(a1, np.float32); a2 = np.ascontiguousarray(a2, np.float32)
        out = np.empty(d1.shape[0], np.int32); n = C.c_int32()
        f1, f2 = fv1.c(), fv2.c()
        _chk(lib().ccm_match_bow_kf_kf(_p(d1), d1.shape[0], _p(has1), _p(a1), C.byref(f1), _p(d2), d2.shape[0], _p(has2), _p(a2), C.byref(f2),
                                       C.c_float(self.nnratio), int(self.checkOri), _p(out), C.byref(n)))
        return out, n.value

    def SearchForTriangulation(self, v1, v2, F12, ex, ey, level_sigma2, scale_factors):
        keep = []

        def view(v):
            arrs = dict(desc=np.ascontiguousarray(v["desc"], np.uint8), has=np.ascontiguousarray(v["has_mp"], np.uint8),
                        xy=np.ascontiguousarray(v["kp_xy"], np.float32), oc=np.ascontiguousarray(v["octave"], np.int32),
                        an=np.ascontiguousarray(v["angle"], np.float32))
            fv = v["fv"].c(); keep.extend([arrs, fv])
            fx, fy, cx, cy = v["intr"]
            return TriViewC(_p(arrs["desc"]), arrs["desc"].shape[0], _p(arrs["has"]), _p(arrs["xy"]), _p(arrs["oc"]), _p(arrs["an"]),
                            C.pointer(fv), fx, fy, cx, cy)
        a, b = view(v1), view(v2)
        F = np.ascontiguousarray(F12, np.float32); ls = np.ascontiguousarray(level_sigma2, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
        pairs = np.empty((min(a.n, b.n) + 1, 2), np.int32); n = C.c_int32()
        _chk(lib().ccm_match_triangulation(C.byref(a), C.byref(b), _p(F), C.c_float(ex), C.c_float(ey), _p(ls), _p(sf), len(ls),
                                           int(self.checkOri), _p(pairs), C.byref(n)))
        return pairs[:n.value].copy()
