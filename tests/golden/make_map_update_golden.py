"""Generates tests/golden/map_update_cv2.npz: the map update after a global BA (Map::RunGBA, S/Map.cpp:1441-1570) evaluated with
OpenCV's own arithmetic — every cv::Mat product of the reference loop is a cv2.gemm call of the same shape (Python cv2 4.13, the
OpenCV generation SURVEY.md §8(c') pins), sums and negations as the MatExpr of the statement fuses them.  This is what pins the
oracle's restatement of cv::gemm's small-matrix rounding (oracle/map_update_oracle.cpp).  Run from the repo root:
    python tests/golden/make_map_update_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from ccm_slam_b200 import synth  # noqa: E402

CASES = [dict(K=60, P=800, seed=21), dict(K=120, P=1500, seed=22, chain=1.0, n_origins=1, new_kf_frac=0.35), dict(K=40, P=600, seed=23, chain=0.0, n_origins=3, outside_frac=0.2)]


def cv2_update(sc):
    import cv2

    def mul(A, B):                       # cv::Mat * cv::Mat
        return cv2.gemm(np.ascontiguousarray(A), np.ascontiguousarray(B), 1.0, None, 0.0)

    def pose_inverse(Tcw):               # KeyFrame::SetPose, S/KeyFrame.cpp:298-306
        Rwc = np.ascontiguousarray(Tcw[:3, :3].T)
        Ow = cv2.gemm(Rwc, np.ascontiguousarray(Tcw[:3, 3:4]), -1.0, None, 0.0)          # Ow = -Rwc*tcw: one gemm with alpha = -1
        Twc = np.eye(4, dtype=np.float32); Twc[:3, :3] = Rwc; Twc[:3, 3:4] = Ow
        return Twc
    K = len(sc["kf_parent"]); P = len(sc["mp_state"])
    pose = sc["kf_Tcw"].copy(); gba = sc["kf_TcwGBA"].copy(); bef = np.full((K, 4, 4), np.nan, np.float32)
    flag = sc["kf_optimized"].astype(bool).copy(); vis = np.zeros(K, np.uint8)
    children = [[] for _ in range(K)]
    for k in range(K):
        if sc["kf_parent"][k] >= 0:
            children[sc["kf_parent"][k]].append(k)
    todo = [k for k in range(K) if sc["kf_parent"][k] == -1]
    while todo:                          # S/Map.cpp:1455-1490
        kf = todo[0]
        Twc = pose_inverse(pose[kf])
        for c in children[kf]:
            if not flag[c]:
                gba[c] = mul(mul(pose[c], Twc), gba[kf]); flag[c] = True
            todo.append(c)
        bef[kf] = pose[kf]; pose[kf] = gba[kf]; vis[kf] = 1
        todo.pop(0)
    out = sc["mp_pos"].copy(); corr = np.zeros(P, np.uint8)
    for i in range(P):                   # S/Map.cpp:1497-1563
        st = sc["mp_state"][i]
        if st == 0:
            continue
        if st == 1:
            out[i] = sc["mp_pos_gba"][i]; corr[i] = 1
            continue
        r = sc["mp_ref"][i]
        if r < 0 or not flag[r] or not vis[r]:
            continue
        Xc = cv2.gemm(np.ascontiguousarray(bef[r][:3, :3]), np.ascontiguousarray(out[i].reshape(3, 1)), 1.0, np.ascontiguousarray(bef[r][:3, 3:4]), 1.0)   # Rcw*Xw + tcw
        Twc = pose_inverse(pose[r])
        out[i] = cv2.gemm(np.ascontiguousarray(Twc[:3, :3]), Xc, 1.0, np.ascontiguousarray(Twc[:3, 3:4]), 1.0).reshape(3)                                    # Rwc*Xc + twc
        corr[i] = 1
    return dict(kf_TcwGBA=gba, kf_visited=vis, mp_pos=out, mp_corrected=corr)


if __name__ == "__main__":
    import cv2
    store = {"cv2_version": np.array(cv2.__version__)}
    for n, kw in enumerate(CASES):
        r = cv2_update(synth.make_map_update(**kw))
        for k, v in r.items():
            store["case%d_%s" % (n, k)] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "map_update_cv2.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")
