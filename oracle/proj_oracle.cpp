// proj_oracle.cpp — CPU oracle for the projection-guided matchers (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// SURVEY.md §8(f) rank 3.  Restates, on flat arrays, the part of each matcher that starts at GetFeaturesInArea:
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea          S/Frame.cpp:103-119, 200-253, 255-265
//   KeyFrame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea       S/KeyFrame.cpp:206-226, 265-275, 1162-1201
//   ORBmatcher::SearchByProjection(Frame&, vector<mpptr>&, th)           S/ORBmatcher.cpp:71-148   ("track")
//   ORBmatcher::SearchByProjection(kfptr, Scw, vpPoints, vpMatched, th)  S/ORBmatcher.cpp:308-446  ("sim3")
//   ORBmatcher::Fuse(kfptr, vector<mpptr>&, th)                          S/ORBmatcher.cpp:854-993
//   ORBmatcher::Fuse(kfptr, Scw, vpPoints, th, vpReplacePoint)           S/ORBmatcher.cpp:995-1122
//   ORBmatcher::SearchBySim3                                             S/ORBmatcher.cpp:1124-1348
//   ORBmatcher::SearchByProjection(Frame&, const Frame& LastFrame, th)   S/ORBmatcher.cpp:1350-1476 ("last")
//   ORBmatcher::SearchByProjection(Frame&, kfptr, sAlreadyFound, th, d)  S/ORBmatcher.cpp:1478-1605 ("reloc")
//   ORBmatcher::SearchForInitialization                                  S/ORBmatcher.cpp:448-563
// Everything BEFORE GetFeaturesInArea in those functions (isBad, the cv::Mat projection, IsInImage, the distance /
// viewing-angle gates, PredictScale) is the caller's prelude: it is f32 cv::Mat arithmetic whose rounding belongs to
// OpenCV, it is O(#points), and the drop-in shim keeps it verbatim in the reference's own types.  A query arrives as
// (valid, u, v, r, level, descriptor).  TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30 (S/ORBmatcher.cpp:63-65).
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
const int TH_HIGH = 100;
const int TH_LOW = 50;
const int HISTO_LENGTH = 30;

int descriptor_distance(const uint8_t* a, const uint8_t* b) {  // S/ORBmatcher.cpp:1653-1669
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t pa, pb;
    memcpy(&pa, a + 4 * i, 4); memcpy(&pb, b + 4 * i, 4);
    unsigned int v = pa ^ pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

void three_maxima(std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {  // S/ORBmatcher.cpp:1607-1648
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

int rot_bin(float a1, float a2) {
  const float factor = 1.0f / HISTO_LENGTH;
  float rot = a1 - a2;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)round(rot * factor);
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}
}  // namespace

extern "C" {

struct orc_grid {
  int32_t n;
  const uint8_t* desc;
  const float* kp_xy;
  const int32_t* octave;
  const float* angle;
  float min_x, min_y, max_x, max_y, grid_w_inv, grid_h_inv;
  int32_t grid_cols, grid_rows;
};

struct orc_queries {
  int32_t m;
  const uint8_t* valid;
  const float* uv;
  const float* radius;
  const int32_t* level;
  const uint8_t* desc;
  const float* angle;
};

}  // extern "C"

namespace {

// mGrid[cols][rows] of feature indices, filled in feature order (AssignFeaturesToGrid)
struct Grid {
  const orc_grid& g;
  std::vector<std::vector<int>> cell;
  explicit Grid(const orc_grid& gg) : g(gg), cell((size_t)gg.grid_cols * gg.grid_rows) {
    for (int i = 0; i < g.n; i++) {
      const int px = (int)round((g.kp_xy[2 * i] - g.min_x) * g.grid_w_inv);
      const int py = (int)round((g.kp_xy[2 * i + 1] - g.min_y) * g.grid_h_inv);
      if (px < 0 || px >= g.grid_cols || py < 0 || py >= g.grid_rows) continue;
      cell[(size_t)px * g.grid_rows + py].push_back(i);
    }
  }
  // GetFeaturesInArea; check_levels follows Frame's (minLevel>0)||(maxLevel>=0); KeyFrame's overload has no levels
  std::vector<int> in_area(float x, float y, float r, int minLevel = -1, int maxLevel = -1) const {
    std::vector<int> out;
    const int nMinCellX = std::max(0, (int)floor((x - g.min_x - r) * g.grid_w_inv));
    if (nMinCellX >= g.grid_cols) return out;
    const int nMaxCellX = std::min(g.grid_cols - 1, (int)ceil((x - g.min_x + r) * g.grid_w_inv));
    if (nMaxCellX < 0) return out;
    const int nMinCellY = std::max(0, (int)floor((y - g.min_y - r) * g.grid_h_inv));
    if (nMinCellY >= g.grid_rows) return out;
    const int nMaxCellY = std::min(g.grid_rows - 1, (int)ceil((y - g.min_y + r) * g.grid_h_inv));
    if (nMaxCellY < 0) return out;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
      for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
        for (int idx : cell[(size_t)ix * g.grid_rows + iy]) {
          if (bCheckLevels) {
            if (g.octave[idx] < minLevel) continue;
            if (maxLevel >= 0 && g.octave[idx] > maxLevel) continue;
          }
          const float distx = g.kp_xy[2 * idx] - x;
          const float disty = g.kp_xy[2 * idx + 1] - y;
          if (fabs(distx) < r && fabs(disty) < r) out.push_back(idx);
        }
    return out;
  }
};

// the "best keypoint in the window at the predicted level or one below" loop shared by Fuse x2, SearchBySim3 and
// SearchByProjection(kf, Scw): returns bestIdx, sets bestDist.  blocked (may be null) = vpMatched[idx] of :391.
// inv_sigma2 (may be null) enables Fuse's chi-square gate of :941-947.
int best_in_window(const Grid& G, const orc_queries& q, int i, const uint8_t* blocked, const float* inv_sigma2, int& bestDist) {
  const float u = q.uv[2 * i], v = q.uv[2 * i + 1];
  const int nPredictedLevel = q.level[i];
  const std::vector<int> vIndices = G.in_area(u, v, q.radius[i]);
  bestDist = INT_MAX;
  int bestIdx = -1;
  for (int idx : vIndices) {
    if (blocked && blocked[idx]) continue;
    const int kpLevel = G.g.octave[idx];
    if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
    if (inv_sigma2) {
      const float ex = u - G.g.kp_xy[2 * idx];
      const float ey = v - G.g.kp_xy[2 * idx + 1];
      const float e2 = ex * ex + ey * ey;
      if (e2 * inv_sigma2[kpLevel] > 5.99) continue;
    }
    const int dist = descriptor_distance(q.desc + 32 * (size_t)i, G.g.desc + 32 * (size_t)idx);
    if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
  }
  return bestIdx;
}

}  // namespace

extern "C" {

// GetFeaturesInArea on its own (for the witness tests); returns the count, writes at most cap indices
int orc_features_in_area(const orc_grid* g, float x, float y, float r, int32_t minLevel, int32_t maxLevel, int32_t* out, int32_t cap) {
  Grid G(*g);
  const std::vector<int> v = G.in_area(x, y, r, minLevel, maxLevel);
  for (size_t k = 0; k < v.size() && (int)k < cap; k++) out[k] = v[k];
  return (int)v.size();
}

// S/ORBmatcher.cpp:71-148.  q.radius = RadiusByViewingCos(mTrackViewCos) [* th] * mvScaleFactors[level], q.level =
// mnTrackScaleLevel, q.valid = mbTrackInView && !isBad().  query_has_obs[i] = pMP->Observations()>0 (decides whether an
// assignment shields the feature from later queries, :107-109); feat_blocked = the same test on the frame's initial
// mvpMapPoints.  match_of_feat[j] = query index last written into F.mvpMapPoints[j], -1 = untouched.
int orc_search_by_projection_track(const orc_grid* g, const orc_queries* q, const uint8_t* query_has_obs, const uint8_t* feat_blocked,
                                   float nnratio, int32_t* match_of_feat) {
  Grid G(*g);
  std::vector<uint8_t> blocked(feat_blocked, feat_blocked + g->n);
  for (int j = 0; j < g->n; j++) match_of_feat[j] = -1;
  int nmatches = 0;
  for (int iMP = 0; iMP < q->m; iMP++) {
    if (!q->valid[iMP]) continue;
    const int nPredictedLevel = q->level[iMP];
    const std::vector<int> vIndices = G.in_area(q->uv[2 * iMP], q->uv[2 * iMP + 1], q->radius[iMP], nPredictedLevel - 1, nPredictedLevel);
    if (vIndices.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : vIndices) {
      if (blocked[idx]) continue;
      const int dist = descriptor_distance(q->desc + 32 * (size_t)iMP, g->desc + 32 * (size_t)idx);
      if (dist < bestDist) {
        bestDist2 = bestDist; bestDist = dist;
        bestLevel2 = bestLevel; bestLevel = g->octave[idx];
        bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = g->octave[idx];
        bestDist2 = dist;
      }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
      match_of_feat[bestIdx] = iMP;
      blocked[bestIdx] = query_has_obs[iMP];
      nmatches++;
    }
  }
  return nmatches;
}

// S/ORBmatcher.cpp:1350-1476 (reloc = 0: LastFrame; block test = map point with observations, threshold TH_HIGH) and
// :1478-1605 (reloc = 1: keyframe; block test = any map point, threshold ORBdist).  Levels [L-1, L+1] through
// Frame::GetFeaturesInArea.  match_of_feat: >= 0 query index, -1 untouched, -2 cleared by the orientation check.
int orc_search_by_projection_frame(const orc_grid* g, const orc_queries* q, const uint8_t* query_has_obs, const uint8_t* feat_blocked,
                                   int32_t reloc, int32_t orb_dist, int32_t check_orientation, int32_t* match_of_feat) {
  Grid G(*g);
  std::vector<uint8_t> blocked(feat_blocked, feat_blocked + g->n);
  for (int j = 0; j < g->n; j++) match_of_feat[j] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  const int th = reloc ? orb_dist : TH_HIGH;
  int nmatches = 0;
  for (int i = 0; i < q->m; i++) {
    if (!q->valid[i]) continue;
    const int L = q->level[i];
    const std::vector<int> vIndices2 = G.in_area(q->uv[2 * i], q->uv[2 * i + 1], q->radius[i], L - 1, L + 1);
    if (vIndices2.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : vIndices2) {
      if (blocked[i2]) continue;
      const int dist = descriptor_distance(q->desc + 32 * (size_t)i, g->desc + 32 * (size_t)i2);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= th) {
      match_of_feat[bestIdx2] = i;
      blocked[bestIdx2] = reloc ? 1 : query_has_obs[i];
      nmatches++;
      if (check_orientation) rotHist[rot_bin(q->angle[i], g->angle[bestIdx2])].push_back(bestIdx2);
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int j : rotHist[i]) { match_of_feat[j] = -2; nmatches--; }
    }
  }
  return nmatches;
}

// S/ORBmatcher.cpp:308-446.  feat_matched = (vpMatched[idx] != nullptr) on entry; existing_idx[i] =
// pMP->GetIndexInKeyFrame(pKF) (-1 = not observed).  best_idx[i] = bestIdx when bestDist <= TH_LOW else -1; a query with
// existing_idx != -1 never touches vpMatched (the caller remaps, :418-432 — dist_newplace equals bestDist so
// bDoNotReplace is never set); otherwise vpMatched[bestIdx] = pMP (match_of_feat) and the match counts.
int orc_search_by_projection_sim3(const orc_grid* g, const orc_queries* q, const uint8_t* feat_matched, const int32_t* existing_idx,
                                  int32_t* best_idx, int32_t* match_of_feat) {
  Grid G(*g);
  std::vector<uint8_t> matched(feat_matched, feat_matched + g->n);
  for (int j = 0; j < g->n; j++) match_of_feat[j] = -1;
  int nmatches = 0;
  for (int iMP = 0; iMP < q->m; iMP++) {
    best_idx[iMP] = -1;
    if (!q->valid[iMP]) continue;
    int bestDist;
    const int bestIdx = best_in_window(G, *q, iMP, matched.data(), nullptr, bestDist);
    if (bestDist <= TH_LOW) {
      best_idx[iMP] = bestIdx;
      if (existing_idx[iMP] == -1) {
        matched[bestIdx] = 1;
        match_of_feat[bestIdx] = iMP;
        nmatches++;
      }
    }
  }
  return nmatches;
}

// the search half of both Fuse overloads (S/ORBmatcher.cpp:854-993 with inv_level_sigma2, :995-1122 without):
// best_idx[i] = bestIdx when bestDist <= TH_LOW, else -1.  The map surgery that follows (:955-990 / :1103-1118) is the
// caller's and runs in query order on these indices.  Returns the number of queries with a best_idx.
int orc_fuse_search(const orc_grid* g, const orc_queries* q, const float* inv_level_sigma2, int32_t* best_idx) {
  Grid G(*g);
  int n = 0;
  for (int i = 0; i < q->m; i++) {
    best_idx[i] = -1;
    if (!q->valid[i]) continue;
    int bestDist;
    const int bestIdx = best_in_window(G, *q, i, nullptr, inv_level_sigma2, bestDist);
    if (bestDist <= TH_LOW) { best_idx[i] = bestIdx; n++; }
  }
  return n;
}

// S/ORBmatcher.cpp:1124-1348.  q12 = map points of KF1 (one query per KF1 feature, valid = has a usable point that
// passed the projection gates into KF2), searched in g2; q21 the converse in g1.  match12[i1] = idx2 when both
// directions agree (:1327-1343), else -1.  Returns nFound.
int orc_search_by_sim3(const orc_grid* g1, const orc_grid* g2, const orc_queries* q12, const orc_queries* q21, int32_t* match12) {
  Grid G1(*g1), G2(*g2);
  const int N1 = q12->m, N2 = q21->m;
  std::vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1);
  for (int i1 = 0; i1 < N1; i1++) {
    if (!q12->valid[i1]) continue;
    int bestDist;
    const int bestIdx = best_in_window(G2, *q12, i1, nullptr, nullptr, bestDist);
    if (bestDist <= TH_HIGH) vnMatch1[i1] = bestIdx;
  }
  for (int i2 = 0; i2 < N2; i2++) {
    if (!q21->valid[i2]) continue;
    int bestDist;
    const int bestIdx = best_in_window(G1, *q21, i2, nullptr, nullptr, bestDist);
    if (bestDist <= TH_HIGH) vnMatch2[i2] = bestIdx;
  }
  int nFound = 0;
  for (int i1 = 0; i1 < N1; i1++) {
    match12[i1] = -1;
    const int idx2 = vnMatch1[i1];
    if (idx2 >= 0 && idx2 < N2) {
      if (vnMatch2[idx2] == i1) { match12[i1] = idx2; nFound++; }
    }
  }
  return nFound;
}

// S/ORBmatcher.cpp:448-563 (monocular initialisation).  Queries = the keypoints of F1 (q.level = their octave, q.uv =
// vbPrevMatched, q.radius = windowSize), searched in F2's grid at octave 0 only; a keypoint of F2 goes to the closest F1 keypoint
// seen so far (vMatchedDistance), an earlier holder loses it.  match12[i1] = i2 or -1.
int orc_search_for_initialization(const orc_grid* g2, const orc_queries* q, float nnratio, int32_t check_orientation, int32_t* match12) {
  Grid G(*g2);
  int nmatches = 0;
  std::vector<int> vnMatches12(q->m, -1), vnMatches21(g2->n, -1), vMatchedDistance(g2->n, INT_MAX);
  std::vector<int> rotHist[HISTO_LENGTH];
  for (int i1 = 0; i1 < q->m; i1++) {
    const int level1 = q->level[i1];
    if (level1 > 0) continue;
    const std::vector<int> vIndices2 = G.in_area(q->uv[2 * i1], q->uv[2 * i1 + 1], q->radius[i1], level1, level1);
    if (vIndices2.empty()) continue;
    int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
    for (int i2 : vIndices2) {
      const int dist = descriptor_distance(q->desc + 32 * (size_t)i1, g2->desc + 32 * (size_t)i2);
      if (vMatchedDistance[i2] <= dist) continue;
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
      else if (dist < bestDist2) bestDist2 = dist;
    }
    if (bestDist <= TH_LOW) {
      if (bestDist < (float)bestDist2 * nnratio) {
        if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
        vnMatches12[i1] = bestIdx2;
        vnMatches21[bestIdx2] = i1;
        vMatchedDistance[bestIdx2] = bestDist;
        nmatches++;
        if (check_orientation) rotHist[rot_bin(q->angle[i1], g2->angle[bestIdx2])].push_back(i1);
      }
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i])
        if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
    }
  }
  for (int i1 = 0; i1 < q->m; i1++) match12[i1] = vnMatches12[i1];
  return nmatches;
}

}  // extern "C"
