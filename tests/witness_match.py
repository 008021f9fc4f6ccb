"""Independent pure-Python witnesses for the projection-guided matchers and the DBoW2 transform (test infrastructure).

Written in a different style from oracle/proj_oracle.cpp and oracle/bow_oracle.cpp on purpose: the lookup grid is not
walked cell by cell — every in-grid feature is tested against the window and the survivors are ordered by
(cell column, cell row, feature index), which is the order Frame/KeyFrame::GetFeaturesInArea emits (S/Frame.cpp:200-253,
S/KeyFrame.cpp:1162-1201); distances come from one numpy bit-count matrix.  All pixel arithmetic is float32.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30


def hamming(A, B):
    return np.unpackbits(A[:, None, :] ^ B[None, :, :], axis=2).sum(axis=2).astype(np.int64)


def _cround(v):  # C round(): half away from zero
    return int(math.floor(abs(float(v)) + 0.5) * (1 if v >= 0 else -1))


class GridW:
    def __init__(self, g):
        self.g = g
        self.x0, self.y0, self.x1, self.y1 = [f32(v) for v in g["bounds"]]
        self.wi = f32(g["cols"]) / f32(self.x1 - self.x0)
        self.hi = f32(g["rows"]) / f32(self.y1 - self.y0)
        xy = np.asarray(g["kp_xy"], f32)
        self.px = np.array([_cround(f32(f32(x - self.x0) * self.wi)) for x in xy[:, 0]])
        self.py = np.array([_cround(f32(f32(y - self.y0) * self.hi)) for y in xy[:, 1]])
        self.ingrid = (self.px >= 0) & (self.px < g["cols"]) & (self.py >= 0) & (self.py < g["rows"])
        self.xy = xy
        self.order = np.lexsort((np.arange(len(xy)), self.py, self.px))  # by px, then py, then index

    def in_area(self, x, y, r, min_level=None, max_level=None):
        x, y, r = f32(x), f32(y), f32(r)
        out = []
        for i in self.order:
            if not self.ingrid[i]:
                continue
            if min_level is not None and not (min_level <= self.g["octave"][i] <= max_level):
                continue
            if abs(f32(self.xy[i, 0] - x)) < r and abs(f32(self.xy[i, 1] - y)) < r:
                out.append(int(i))
        return out


def _rot_bin(a1, a2):
    rot = f32(f32(a1) - f32(a2))
    if rot < 0.0:
        rot = f32(rot + f32(360.0))
    b = _cround(f32(rot * f32(f32(1.0) / f32(HISTO_LENGTH))))
    return 0 if b == HISTO_LENGTH else b


def _three_maxima(sizes):
    m = [0, 0, 0]; ind = [-1, -1, -1]
    for i, s in enumerate(sizes):
        if s > m[0]:
            m = [s, m[0], m[1]]; ind = [i, ind[0], ind[1]]
        elif s > m[1]:
            m = [m[0], s, m[1]]; ind = [ind[0], i, ind[1]]
        elif s > m[2]:
            m[2] = s; ind[2] = i
    if m[1] < f32(0.1) * f32(m[0]):
        ind[1] = ind[2] = -1
    elif m[2] < f32(0.1) * f32(m[0]):
        ind[2] = -1
    return ind


def search_track(g, q, query_has_obs, feat_blocked, nnratio):
    G = GridW(g); D = hamming(np.asarray(q["desc"]), np.asarray(g["desc"]))
    blocked = np.array(feat_blocked, bool).copy(); out = np.full(len(blocked), -1, np.int64); n = 0
    for i in range(len(q["valid"])):
        if not q["valid"][i]:
            continue
        L = int(q["level"][i])
        cand = [j for j in G.in_area(q["uv"][i, 0], q["uv"][i, 1], q["radius"][i], L - 1, L) if not blocked[j]]
        if not cand:
            continue
        best = (256, -1, -1); second = (256, -1)
        for j in cand:
            d = int(D[i, j])
            if d < best[0]:
                second = (best[0], best[1]); best = (d, int(g["octave"][j]), j)
            elif d < second[0]:
                second = (d, int(g["octave"][j]))
        if best[0] <= TH_HIGH:
            if best[1] == second[1] and f32(best[0]) > f32(nnratio) * f32(second[0]):
                continue
            out[best[2]] = i; blocked[best[2]] = bool(query_has_obs[i]); n += 1
    return out, n


def search_frame(g, q, query_has_obs, feat_blocked, reloc, orb_dist, check_ori):
    G = GridW(g); D = hamming(np.asarray(q["desc"]), np.asarray(g["desc"]))
    blocked = np.array(feat_blocked, bool).copy(); out = np.full(len(blocked), -1, np.int64); n = 0
    hist = [[] for _ in range(HISTO_LENGTH)]
    th = orb_dist if reloc else TH_HIGH
    for i in range(len(q["valid"])):
        if not q["valid"][i]:
            continue
        L = int(q["level"][i])
        cand = [j for j in G.in_area(q["uv"][i, 0], q["uv"][i, 1], q["radius"][i], L - 1, L + 1) if not blocked[j]]
        if not cand:
            continue
        d, j = min((int(D[i, j]), k) for k, j in enumerate(cand))   # first minimum in visiting order
        j = cand[j]
        if d < 256 and d <= th:
            out[j] = i; blocked[j] = True if reloc else bool(query_has_obs[i]); n += 1
            if check_ori:
                hist[_rot_bin(q["angle"][i], g["angle"][j])].append(j)
    if check_ori:
        keep = _three_maxima([len(h) for h in hist])
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for j in hist[b]:
                    out[j] = -2; n -= 1
    return out, n


def _best(G, D, g, q, i, blocked=None, inv_sigma2=None):
    L = int(q["level"][i]); u, v = f32(q["uv"][i, 0]), f32(q["uv"][i, 1])
    best = (1 << 30, -1)
    for j in G.in_area(u, v, q["radius"][i]):
        if blocked is not None and blocked[j]:
            continue
        o = int(g["octave"][j])
        if o < L - 1 or o > L:
            continue
        if inv_sigma2 is not None:
            ex = f32(u - G.xy[j, 0]); ey = f32(v - G.xy[j, 1])
            e2 = f32(f32(ex * ex) + f32(ey * ey))
            if float(f32(e2 * f32(inv_sigma2[o]))) > 5.99:
                continue
        if int(D[i, j]) < best[0]:
            best = (int(D[i, j]), j)
    return best


def search_sim3proj(g, q, feat_matched, existing_idx):
    G = GridW(g); D = hamming(np.asarray(q["desc"]), np.asarray(g["desc"]))
    matched = np.array(feat_matched, bool).copy(); out = np.full(len(matched), -1, np.int64); best_idx = np.full(len(q["valid"]), -1, np.int64); n = 0
    for i in range(len(q["valid"])):
        if not q["valid"][i]:
            continue
        d, j = _best(G, D, g, q, i, blocked=matched)
        if d <= TH_LOW:
            best_idx[i] = j
            if existing_idx[i] == -1:
                matched[j] = True; out[j] = i; n += 1
    return best_idx, out, n


def fuse_search(g, q, inv_sigma2=None):
    G = GridW(g); D = hamming(np.asarray(q["desc"]), np.asarray(g["desc"]))
    best_idx = np.full(len(q["valid"]), -1, np.int64)
    for i in range(len(q["valid"])):
        if q["valid"][i]:
            d, j = _best(G, D, g, q, i, inv_sigma2=inv_sigma2)
            if d <= TH_LOW:
                best_idx[i] = j
    return best_idx, int((best_idx >= 0).sum())


def search_by_sim3(g1, g2, q12, q21):
    def one_way(g, q):
        G = GridW(g); D = hamming(np.asarray(q["desc"]), np.asarray(g["desc"])); m = np.full(len(q["valid"]), -1, np.int64)
        for i in range(len(m)):
            if q["valid"][i]:
                d, j = _best(G, D, g, q, i)
                if d <= TH_HIGH:
                    m[i] = j
        return m
    m1 = one_way(g2, q12); m2 = one_way(g1, q21)
    out = np.full(len(m1), -1, np.int64)
    for i1, j in enumerate(m1):
        if j >= 0 and m2[j] == i1:
            out[i1] = j
    return out, int((out >= 0).sum())


def search_init(g2, q, nnratio, check_ori):
    G = GridW(g2); D = hamming(np.asarray(q["desc"]), np.asarray(g2["desc"]))
    m12 = np.full(len(q["valid"]), -1, np.int64); holder = {}; held = {}; n = 0
    hist = [[] for _ in range(HISTO_LENGTH)]
    for i in range(len(m12)):
        if q["level"][i] > 0:
            continue
        cand = [(int(D[i, j]), j) for j in G.in_area(q["uv"][i, 0], q["uv"][i, 1], q["radius"][i], 0, 0) if held.get(j, 1 << 30) > int(D[i, j])]
        if not cand:
            continue
        order = sorted(range(len(cand)), key=lambda k: (cand[k][0], k))     # stable: first minimum first
        d1, j1 = cand[order[0]]
        d2 = cand[order[1]][0] if len(cand) > 1 else (1 << 31) - 1
        if d1 <= TH_LOW and f32(d1) < f32(f32(d2) * f32(nnratio)):
            if j1 in holder:
                m12[holder[j1]] = -1; n -= 1
            m12[i] = j1; holder[j1] = i; held[j1] = d1; n += 1
            if check_ori:
                hist[_rot_bin(q["angle"][i], g2["angle"][j1])].append(i)
    if check_ori:
        keep = _three_maxima([len(h) for h in hist])
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for i in hist[b]:
                    if m12[i] >= 0:
                        m12[i] = -1; n -= 1
    return m12, n


# ---- DBoW2 transform ----------------------------------------------------------------------------------------------
def voc_transform(voc, feat, levelsup):
    """D/TemplatedVocabulary.h:1127-1192, 1219-1260 with python containers (dict of children lists, dict accumulators)."""
    N = len(voc["parent"])
    children = {i: [] for i in range(N)}
    word_id = {}; nw = 0
    for nid in range(1, N):
        children[int(voc["parent"][nid])].append(nid)
        if voc["is_leaf"][nid]:
            word_id[nid] = nw; nw += 1
    bow = {}; fv = {}; per = []
    nid_level = voc["L"] - levelsup
    desc = np.asarray(voc["desc"])
    for i, f in enumerate(np.asarray(feat)):
        cur, level, nid = 0, 0, 0
        while True:
            level += 1
            ch = children[cur]
            d = np.unpackbits(desc[ch] ^ f[None, :], axis=1).sum(axis=1)
            cur = ch[int(np.argmin(d))]             # argmin returns the first minimum
            if level == nid_level:
                nid = cur
            if not children[cur]:
                break
        w = float(voc["weight"][cur]); wid = word_id.get(cur, 0)
        per.append((wid, nid, w))
        if w > 0:
            if voc["weighting"] in (0, 1):
                bow[wid] = bow.get(wid, 0.0) + w if wid in bow else w
            else:
                bow.setdefault(wid, w)
            fv.setdefault(nid, []).append(i)
    ids = sorted(bow)
    vals = [bow[k] for k in ids]
    must = voc["scoring"] != 5
    if voc["weighting"] in (0, 1) and ids and not must:
        vals = [v / float(len(ids)) for v in vals]
    if must:
        norm = 0.0
        if voc["scoring"] == 1:
            for v in vals:
                norm += v * v
            norm = math.sqrt(norm)
        else:
            for v in vals:
                norm += abs(v)
        if norm > 0.0:
            vals = [v / norm for v in vals]
    return dict(per=per, bow_id=ids, bow_val=vals, fv={k: fv[k] for k in sorted(fv)})
