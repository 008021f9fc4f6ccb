"""Front-end timings for bench.py: ms per frame / per call of the ORB extractor and the two BoW-guided matchers of the north-star path,
through the C ABI with HOST buffers (copies inside the timed region — these are latency-bound per-frame calls), next to
the CPU oracle on the same inputs when an oracle module is passed in.

  ORBextractor::operator()        S/ORBextractor.cpp:1216-1278   752 x 480, 1000 features, 8 levels (cslam/conf/config.yaml:38-51)
  ORBmatcher::SearchByBoW         S/ORBmatcher.cpp:178-306       keyframe (≈1000 features) against frame
  ORBmatcher::SearchForTriangulation  S/ORBmatcher.cpp:700-852   keyframe pair, epipolar gate

Plumbing only: every timed call is the product's own entry point.
"""
from __future__ import annotations

import statistics
import threading
import time

import numpy as np

from . import synth
from .frontend import FeatureVector, ORBextractor, ORBmatcher
from .synth_images import make_image


def _median_ms(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


def run(oracle=None, reps=30, cpu_reps=3):
    out = {"unit": "ms (median)", "image": "752x480 synthetic (seeded value noise + rectangles), 1000 features, 8 levels, scale 1.2",
           "note": "host buffers, H2D/D2H inside every timed call; cpu = the oracle port, single thread"}
    imgs = [make_image(s) for s in range(4)]
    ex = ORBextractor()
    k1, d1 = ex(imgs[0])
    it = [0]

    def one():
        it[0] += 1
        ex(imgs[it[0] % 4])
    out["orb_extract_ms_per_frame"] = _median_ms(one, reps, warm=3)
    out["orb_keypoints"] = int(len(k1))
    # four agents at once: one extractor (own stream) per agent, one host thread each, as the four client front ends would call it
    exs = [ORBextractor() for _ in range(4)]
    for e, im in zip(exs, imgs):
        e(im)

    def four():
        th = [threading.Thread(target=lambda e=e, im=im: [e(im) for _ in range(5)]) for e, im in zip(exs, imgs)]
        for t in th: t.start()
        for t in th: t.join()
    out["orb_extract_4_agents_ms_per_frame"] = _median_ms(four, max(3, reps // 6), warm=1) / 20.0
    for e in exs:
        e.close()
    # matchers on two views of one scene (a shifted copy: many true matches)
    b = np.roll(imgs[0], (3, 5), axis=(0, 1))
    k2, d2 = ex(b)
    ex.close()
    rng = np.random.default_rng(1)
    node = lambda d: (d[:, 0].astype(np.int64) * 7 + d[:, 1] // 64) % 97   # ~100 vocabulary nodes like DBoW2 at levelsup = 4
    fv1, fv2 = FeatureVector(node(d1)), FeatureVector(node(d2))
    has1 = (rng.random(len(k1)) < 0.7).astype(np.uint8); has2 = (rng.random(len(k2)) < 0.7).astype(np.uint8)
    m = ORBmatcher(0.7, True)
    got_b, n_b = m.SearchByBoW_KF_Frame(d1, has1, k1["angle"], fv1, d2, k2["angle"], fv2)
    out["search_by_bow_ms_per_call"] = _median_ms(lambda: m.SearchByBoW_KF_Frame(d1, has1, k1["angle"], fv1, d2, k2["angle"], fv2), reps)
    out["search_by_bow_matches"] = int(n_b)
    fx, fy, cx, cy = [np.float32(v) for v in synth.EUROC_INTR]
    Kinv = np.linalg.inv(np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64))
    tx = np.array([[0, 0, 0], [0, 0, -1.0], [0, 1.0, 0]])
    F12 = (Kinv.T @ tx @ Kinv).astype(np.float32)
    sf = (1.2 ** np.arange(8)).astype(np.float32); ls2 = (sf * sf).astype(np.float32)
    v = lambda k, d, has, fv: dict(desc=d, has_mp=has, kp_xy=np.stack([k["x"], k["y"]], 1), octave=k["octave"], angle=k["angle"], fv=fv,
                                   intr=(fx, fy, cx, cy))
    mt = ORBmatcher(0.6, False)
    got_t = mt.SearchForTriangulation(v(k1, d1, has1, fv1), v(k2, d2, has2, fv2), F12, -5000.0, float(cy), ls2, sf)
    out["search_for_triangulation_ms_per_call"] = _median_ms(
        lambda: mt.SearchForTriangulation(v(k1, d1, has1, fv1), v(k2, d2, has2, fv2), F12, -5000.0, float(cy), ls2, sf), reps)
    out["search_for_triangulation_matches"] = int(len(got_t))
    if oracle is not None:
        rk, rd = oracle.orb_extract(imgs[0])
        out["cpu_orb_extract_ms_per_frame"] = _median_ms(lambda: oracle.orb_extract(imgs[1]), cpu_reps, warm=0)
        ofv1, ofv2 = oracle.FeatureVector(node(d1)), oracle.FeatureVector(node(d2))
        ref_b, rn = oracle.match_bow_kf_frame(d1, has1, k1["angle"], ofv1, d2, k2["angle"], ofv2, 0.7, True)
        out["cpu_search_by_bow_ms_per_call"] = _median_ms(
            lambda: oracle.match_bow_kf_frame(d1, has1, k1["angle"], ofv1, d2, k2["angle"], ofv2, 0.7, True), cpu_reps * 3, warm=1)
        ref_t = oracle.match_triangulation(v(k1, d1, has1, ofv1), v(k2, d2, has2, ofv2), F12, -5000.0, float(cy), ls2, sf, False)
        out["cpu_search_for_triangulation_ms_per_call"] = _median_ms(
            lambda: oracle.match_triangulation(v(k1, d1, has1, ofv1), v(k2, d2, has2, ofv2), F12, -5000.0, float(cy), ls2, sf, False),
            cpu_reps * 3, warm=1)
        out["parity"] = {"orb_bit_exact": bool(len(rk) == len(k1) and np.array_equal(rd, d1) and np.array_equal(rk["x"], k1["x"])),
                         "bow_indices_equal": bool(rn == n_b and np.array_equal(ref_b, got_b)),
                         "triangulation_indices_equal": bool(np.array_equal(ref_t, got_t))}
    return out
