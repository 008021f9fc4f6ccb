"""Seeded synthetic inputs for the projection-guided matchers and the DBoW2 transform.

A *grid* is the image side of a matcher (Frame or KeyFrame: undistorted keypoints, octaves, angles, descriptors and the
64-ish x 48 lookup grid geometry); *queries* are map points after the caller's projection prelude (valid flag, projected
pixel, search radius, predicted level, descriptor).  Plain dicts of numpy arrays: the product binding
(ccm_slam_b200.frontend) and the test-side checker each build their own C structs from them.
"""
from __future__ import annotations

import numpy as np

SCALE_FACTORS = (np.float32(1.2) ** np.arange(8)).astype(np.float32)          # ORBextractor scale table
LEVEL_SIGMA2 = (SCALE_FACTORS * SCALE_FACTORS).astype(np.float32)
INV_LEVEL_SIGMA2 = (np.float32(1.0) / LEVEL_SIGMA2).astype(np.float32)
_QUOTA = np.array([217, 181, 151, 126, 105, 87, 73, 60], np.float64)


def flip_bits(desc, nbits, rng):
    """copies of `desc` (n x 32 u8) with `nbits[i]` random bits flipped in row i"""
    out = desc.copy()
    for i in range(out.shape[0]):
        if nbits[i] <= 0:
            continue
        pos = rng.choice(256, size=int(nbits[i]), replace=False)
        np.bitwise_xor.at(out[i], pos // 8, (1 << (pos % 8)).astype(np.uint8))
    return out


def make_grid(n=1000, seed=0, bounds=(-10.5, -8.25, 761.0, 489.5), cols=75, rows=48, clustered=True):
    """Features of one image.  A few keypoints lie outside the bounds (undistortion can do that; PosInGrid drops them)."""
    rng = np.random.default_rng(seed)
    x0, y0, x1, y1 = bounds
    if clustered:  # textured regions: many features per window so that best/second-best and ties matter
        centres = rng.uniform([x0 + 30, y0 + 30], [x1 - 30, y1 - 30], size=(max(4, n // 25), 2))
        xy = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 9.0, size=(n, 2))
    else:
        xy = rng.uniform([x0, y0], [x1, y1], size=(n, 2))
    n_out = max(1, n // 50)
    xy[:n_out] = rng.uniform([x0 - 20, y0 - 20], [x0 - 1, y0 - 1], size=(n_out, 2))   # out of the grid
    xy[n_out:2 * n_out, 0] = x1 + rng.uniform(0.0, 3.0, n_out)                        # right at / past the right edge
    xy = np.round(xy * 4) / 4                                                          # quarter-pixel positions: exact ties in |dx| < r
    octave = rng.choice(8, size=n, p=_QUOTA / _QUOTA.sum()).astype(np.int32)
    angle = rng.uniform(0, 360, n).astype(np.float32)
    desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    # near-duplicate descriptors inside clusters (repetitive texture): forces the ratio test and first-wins ties
    for i in range(0, n - 1, 7):
        desc[i + 1] = flip_bits(desc[i:i + 1], [rng.integers(0, 3)], rng)[0]
    return dict(desc=desc, kp_xy=xy.astype(np.float32), octave=octave, angle=angle, bounds=bounds, cols=cols, rows=rows)


def make_queries(grid, m=1500, seed=1, th=3.0, noise_px=2.0, invalid_frac=0.1, max_flip=70, dup_frac=0.15):
    """Map points projected into `grid`'s image.  Most are noisy copies of a feature (true matches); some are duplicates of
    an earlier query (two points competing for one feature); some are random (no match); some are invalid."""
    rng = np.random.default_rng(seed)
    n = grid["desc"].shape[0]
    src = rng.integers(0, n, m)
    dup = rng.random(m) < dup_frac
    for i in range(1, m):
        if dup[i]:
            src[i] = src[rng.integers(0, i)]
    uv = grid["kp_xy"][src] + rng.normal(0, noise_px, size=(m, 2)).astype(np.float32)
    uv = (np.round(uv * 4) / 4).astype(np.float32)
    level = np.clip(grid["octave"][src] + rng.choice([0, 0, 0, 1, 1, -1, 2], size=m), 0, 7).astype(np.int32)
    nflip = rng.integers(0, max_flip, m)
    desc = flip_bits(grid["desc"][src], nflip, rng)
    rnd = rng.random(m) < 0.1
    desc[rnd] = rng.integers(0, 256, size=(int(rnd.sum()), 32), dtype=np.uint8)
    radius = (np.float32(th) * SCALE_FACTORS[level]).astype(np.float32)
    valid = (rng.random(m) >= invalid_frac).astype(np.uint8)
    angle = ((grid["angle"][src] + rng.normal(0, 4.0, m) + np.where(rng.random(m) < 0.1, rng.uniform(0, 360, m), 0.0)) % 360).astype(np.float32)
    return dict(valid=valid, uv=uv, radius=radius, level=level, desc=desc, angle=angle, src=src)


def tie_storm(grid, queries=None, pool=6, seed=0):
    """Replace every descriptor by one of `pool` patterns that lie 8..30 bits from a common base (copies of the inputs are returned).
    Every window then holds many candidates at exactly the same distance, below the acceptance thresholds: what is selected is decided
    by the visiting order alone (first minimum, second best on the same level, who keeps a contested feature)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(1, 32), dtype=np.uint8)
    pats = np.concatenate([flip_bits(base, [int(rng.integers(8, 31))], rng) for _ in range(pool)])
    g = dict(grid); g["desc"] = pats[rng.integers(0, pool, grid["desc"].shape[0])]
    if queries is None:
        return g
    q = dict(queries); q["desc"] = pats[rng.integers(0, pool, queries["desc"].shape[0])]
    return g, q


def make_vocabulary(k=10, L=3, seed=0, scoring=0, weighting=0, early_leaf_frac=0.05, tie_frac=0.05, stop_frac=0.02):
    """A DBoW2-style vocabulary tree as the rows of its text file (row 0 = root): parent, leaf flag, descriptor, weight.
    Hierarchical: children are perturbed copies of their parent, so descents are decisive at the top and close at the
    bottom; some siblings are exact duplicates (first-minimum-wins ties); some inner nodes stop early (ragged depth);
    a few words have weight 0 (stopped words)."""
    rng = np.random.default_rng(seed)
    parent, is_leaf, desc, weight = [0], [0], [np.zeros(32, np.uint8)], [0.0]

    def grow(pid, pdesc, level):
        flips = max(6, 110 >> (level - 1))
        prev = None
        for c in range(k):
            nid = len(parent)
            d = flip_bits(pdesc[None, :], [flips], rng)[0] if level > 1 else rng.integers(0, 256, 32, dtype=np.uint8)
            if prev is not None and rng.random() < tie_frac:
                d = prev.copy()
            prev = d
            leaf = level == L or (level > 1 and rng.random() < early_leaf_frac)
            parent.append(pid); is_leaf.append(1 if leaf else 0); desc.append(d)
            weight.append(0.0 if (not leaf or rng.random() < stop_frac) else float(rng.uniform(0.5, 12.0)))
            if not leaf:
                grow(nid, d, level + 1)

    grow(0, desc[0], 1)
    return dict(k=k, L=L, scoring=scoring, weighting=weighting, parent=np.array(parent, np.int32), is_leaf=np.array(is_leaf, np.uint8),
                desc=np.stack(desc).astype(np.uint8), weight=np.array(weight, np.float64))


def make_voc_features(voc, n=1000, seed=0):
    """descriptors near random leaves of `voc` (plus some uniformly random ones)"""
    rng = np.random.default_rng(seed)
    leaves = np.flatnonzero(voc["is_leaf"])
    pick = leaves[rng.integers(0, len(leaves), n)]
    d = flip_bits(voc["desc"][pick], rng.integers(0, 40, n), rng)
    rnd = rng.random(n) < 0.1
    d[rnd] = rng.integers(0, 256, size=(int(rnd.sum()), 32), dtype=np.uint8)
    return d


def make_init_pair(n=1000, seed=0, window=100.0, shift=(6.0, -4.0)):
    """Two frames for SearchForInitialization: F2 holds shifted, slightly perturbed copies of most of F1's keypoints.  Returns
    (grid of F2, queries = one per keypoint of F1 with level = its octave, uv = its own position (vbPrevMatched at the first call),
    radius = windowSize)."""
    rng = np.random.default_rng(seed)
    g1 = make_grid(n=n, seed=seed + 100)
    g1["octave"][: n // 2] = 0                      # the matcher only looks at octave 0
    g2 = make_grid(n=n, seed=seed + 200)
    keep = rng.permutation(n)[: (3 * n) // 4]
    g2["kp_xy"][keep] = g1["kp_xy"][keep] + np.float32(shift) + rng.normal(0, 0.75, (len(keep), 2)).astype(np.float32)
    g2["octave"][keep] = g1["octave"][keep]
    g2["angle"][keep] = (g1["angle"][keep] + rng.normal(0, 3.0, len(keep))).astype(np.float32) % np.float32(360)
    g2["desc"][keep] = flip_bits(g1["desc"][keep], rng.integers(0, 45, len(keep)), rng)
    q = dict(valid=np.ones(n, np.uint8), uv=g1["kp_xy"].copy(), radius=np.full(n, window, np.float32), level=g1["octave"].copy(),
             desc=g1["desc"], angle=g1["angle"])
    return g2, q
