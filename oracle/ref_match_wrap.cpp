// ref_match_wrap.cpp — C entry points over the REFERENCE's own ORBmatcher.cpp (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// oracle/Makefile compiles cslam/src/ORBmatcher.cpp where it lies under /root/reference, together with this file, into
// oracle/_ref/libmatch_ref.so.  Frame / KeyFrame / MapPoint are the stand-ins of oracle/ref_stub/cslam/Frame.h (plain structs with the
// members the matcher source names); cv::Mat and friends are oracle/ref_stub/opencv2.  The search methods themselves — candidate
// walks, best / second-best bookkeeping, thresholds, ratio tests, rotation histograms, mutual checks, the geometric gates in front of
// them — are the reference's object code.  Each entry point builds the stand-in objects from flat arrays, calls ONE reference method
// and flattens what it did.  tests/test_oracle_vs_reference_matchers.py holds oracle/match_oracle.cpp and oracle/proj_oracle.cpp to it.
#include <cslam/ORBmatcher.h>

#include <cstdint>

namespace cslam {
float Frame::fx = 0, Frame::fy = 0, Frame::cx = 0, Frame::cy = 0, Frame::mnMinX = 0, Frame::mnMinY = 0, Frame::mnMaxX = 0, Frame::mnMaxY = 0,
      Frame::mfGridElementWidthInv = 0, Frame::mfGridElementHeightInv = 0;
}

using namespace cslam;
typedef boost::shared_ptr<KeyFrame> kfptr;
typedef boost::shared_ptr<MapPoint> mpptr;
typedef boost::shared_ptr<Frame> frameptr;

extern "C" {

struct ref_fv { int32_t n_nodes; const uint32_t* node_id; const int32_t* node_ptr; const uint32_t* feat; };

// one image (Frame or KeyFrame): keypoints + lookup grid + camera; pose as row-major R (9) and t (3)
struct ref_image {
  int32_t n;
  const uint8_t* desc; const float* kp_xy; const int32_t* octave; const float* angle;
  float min_x, min_y, max_x, max_y, grid_w_inv, grid_h_inv;
  int32_t grid_cols, grid_rows;
  float fx, fy, cx, cy;
  const float* R; const float* t;          // Tcw = [R | t]; may be NULL (identity, zero)
  int32_t nlevels; const float* scale_factors; const float* level_sigma2; const float* inv_level_sigma2; float log_scale_factor;
  const ref_fv* fv;                         // may be NULL
};

// map points (the query side)
struct ref_points {
  int32_t m;
  const float* pos; const float* normal; const float* min_dist; const float* max_dist;   // mfMinDistance / mfMaxDistance (un-scaled)
  const uint8_t* desc; const uint8_t* bad; const uint8_t* do_not_replace; const int32_t* n_obs;
  const int32_t* index_in_kf;               // GetIndexInKeyFrame(pKF) for the keyframe of the call, -1 = none; may be NULL
  // tracking fields (SearchByProjection(Frame&, vpMapPoints)); may be NULL
  const uint8_t* track_in_view; const float* track_xy; const int32_t* track_level; const float* track_view_cos;
};

}  // extern "C"

namespace {

cv::Mat mat_u8(const uint8_t* d, int rows, int cols) {
  cv::Mat m(rows, cols, CV_8U);
  for (int r = 0; r < rows; r++) memcpy(m.ptr(r), d + (size_t)r * cols, cols);
  return m;
}
cv::Mat mat_f32(const float* d, int rows, int cols) {
  cv::Mat m(rows, cols, CV_32F);
  for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.at<float>(r, c) = d[r * cols + c];
  return m;
}
std::vector<cv::KeyPoint> keypoints(const ref_image& im) {
  std::vector<cv::KeyPoint> k(im.n);
  for (int i = 0; i < im.n; i++) {
    k[i].pt.x = im.kp_xy[2 * i]; k[i].pt.y = im.kp_xy[2 * i + 1];
    k[i].octave = im.octave ? im.octave[i] : 0; k[i].angle = im.angle ? im.angle[i] : 0.f;
  }
  return k;
}
void feature_vector(const ref_fv* f, DBoW2::FeatureVector& out) {
  out.clear();
  if (!f) return;
  for (int a = 0; a < f->n_nodes; a++)
    for (int k = f->node_ptr[a]; k < f->node_ptr[a + 1]; k++) out.addFeature(f->node_id[a], f->feat[k]);
}
FeatureGridStandIn grid_of(const ref_image& im, const std::vector<cv::KeyPoint>& keys) {
  FeatureGridStandIn g;
  g.minX = im.min_x; g.minY = im.min_y; g.maxX = im.max_x; g.maxY = im.max_y; g.wInv = im.grid_w_inv; g.hInv = im.grid_h_inv;
  g.cols = im.grid_cols; g.rows = im.grid_rows;
  if (g.cols > 0 && g.rows > 0) g.build(keys);
  return g;
}
const float IDENT[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, ZERO3[3] = {0, 0, 0};

kfptr make_kf(const ref_image& im) {
  kfptr k(new KeyFrame);
  k->N = im.n; k->mvKeysUn = keypoints(im);
  k->mDescriptors = mat_u8(im.desc, im.n, 32);
  k->mvpMapPoints.assign(im.n, mpptr());
  feature_vector(im.fv, k->mFeatVec);
  k->fx = im.fx; k->fy = im.fy; k->cx = im.cx; k->cy = im.cy;
  k->mnMinX = (int)im.min_x; k->mnMinY = (int)im.min_y; k->mnMaxX = (int)im.max_x; k->mnMaxY = (int)im.max_y;
  k->mnGridCols = im.grid_cols; k->mnGridRows = im.grid_rows; k->mfGridElementWidthInv = im.grid_w_inv; k->mfGridElementHeightInv = im.grid_h_inv;
  k->grid = grid_of(im, k->mvKeysUn);
  k->Rcw = mat_f32(im.R ? im.R : IDENT, 3, 3); k->tcw = mat_f32(im.t ? im.t : ZERO3, 3, 1);
  k->Ow = -k->Rcw.t() * k->tcw;
  k->mnScaleLevels = im.nlevels; k->mfLogScaleFactor = im.log_scale_factor;
  if (im.scale_factors) k->mvScaleFactors.assign(im.scale_factors, im.scale_factors + im.nlevels);
  if (im.level_sigma2) k->mvLevelSigma2.assign(im.level_sigma2, im.level_sigma2 + im.nlevels);
  if (im.inv_level_sigma2) k->mvInvLevelSigma2.assign(im.inv_level_sigma2, im.inv_level_sigma2 + im.nlevels);
  return k;
}
frameptr make_frame(const ref_image& im) {
  frameptr f(new Frame);
  f->N = im.n; f->mvKeysUn = keypoints(im); f->mvKeys = f->mvKeysUn;
  f->mDescriptors = mat_u8(im.desc, im.n, 32);
  f->mvpMapPoints.assign(im.n, mpptr()); f->mvbOutlier.assign(im.n, false);
  feature_vector(im.fv, f->mFeatVec);
  Frame::fx = im.fx; Frame::fy = im.fy; Frame::cx = im.cx; Frame::cy = im.cy;
  Frame::mnMinX = im.min_x; Frame::mnMinY = im.min_y; Frame::mnMaxX = im.max_x; Frame::mnMaxY = im.max_y;
  Frame::mfGridElementWidthInv = im.grid_w_inv; Frame::mfGridElementHeightInv = im.grid_h_inv;
  f->grid = grid_of(im, f->mvKeysUn);
  cv::Mat T(4, 4, CV_32F);
  T = cv::Mat::zeros(4, 4, CV_32F);
  const float* R = im.R ? im.R : IDENT; const float* t = im.t ? im.t : ZERO3;
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T.at<float>(r, c) = R[3 * r + c]; T.at<float>(r, 3) = t[r]; }
  T.at<float>(3, 3) = 1.f;
  f->mTcw = T;
  f->mnScaleLevels = im.nlevels; f->mfLogScaleFactor = im.log_scale_factor;
  if (im.scale_factors) f->mvScaleFactors.assign(im.scale_factors, im.scale_factors + im.nlevels);
  return f;
}
std::vector<mpptr> make_points(const ref_points& p, const KeyFrame* kf_for_index) {
  std::vector<mpptr> v(p.m);
  for (int i = 0; i < p.m; i++) {
    mpptr m(new MapPoint);
    m->tag = i;
    m->bad = p.bad && p.bad[i];
    if (p.pos) m->pos = mat_f32(p.pos + 3 * i, 3, 1);
    if (p.normal) m->normal = mat_f32(p.normal + 3 * i, 3, 1);
    if (p.desc) m->desc = mat_u8(p.desc + 32 * (size_t)i, 1, 32);
    if (p.min_dist) m->mfMinDistance = p.min_dist[i];
    if (p.max_dist) m->mfMaxDistance = p.max_dist[i];
    m->nObs = p.n_obs ? p.n_obs[i] : 1;
    m->mbDoNotReplace = p.do_not_replace && p.do_not_replace[i];
    if (p.index_in_kf && kf_for_index && p.index_in_kf[i] >= 0) m->indexIn[kf_for_index] = p.index_in_kf[i];
    if (p.track_in_view) {
      m->mbTrackInView = p.track_in_view[i] != 0; m->mTrackProjX = p.track_xy[2 * i]; m->mTrackProjY = p.track_xy[2 * i + 1];
      m->mnTrackScaleLevel = p.track_level[i]; m->mTrackViewCos = p.track_view_cos[i];
    }
    v[i] = m;
  }
  return v;
}
// a placeholder map point for "this keypoint already holds a point" flags
mpptr holder(int n_obs) { mpptr m(new MapPoint); m->tag = -100; m->nObs = n_obs; return m; }

}  // namespace

extern "C" {

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(mat_u8(a, 1, 32), mat_u8(b, 1, 32)); }

// SearchByBoW(kfptr, Frame&, vpMapPointMatches): kf_has_mp[i] = the keyframe's feature i holds a good map point
int ref_match_bow_kf_frame(const ref_image* kf, const uint8_t* kf_has_mp, const ref_image* fr, float nnratio, int check_ori, int32_t* match_kf_of_f) {
  kfptr K = make_kf(*kf);
  for (int i = 0; i < kf->n; i++) if (kf_has_mp[i]) { K->mvpMapPoints[i] = mpptr(new MapPoint); K->mvpMapPoints[i]->tag = i; }
  frameptr F = make_frame(*fr);
  ORBmatcher matcher(nnratio, check_ori != 0);
  std::vector<mpptr> out;
  const int n = matcher.SearchByBoW(K, *F, out);
  for (int j = 0; j < fr->n; j++) match_kf_of_f[j] = out[j] ? out[j]->tag : -1;
  return n;
}

// SearchByBoW(kfptr, kfptr, vpMatches12)
int ref_match_bow_kf_kf(const ref_image* k1, const uint8_t* has1, const ref_image* k2, const uint8_t* has2, float nnratio, int check_ori, int32_t* match12) {
  kfptr K1 = make_kf(*k1), K2 = make_kf(*k2);
  for (int i = 0; i < k1->n; i++) if (has1[i]) { K1->mvpMapPoints[i] = mpptr(new MapPoint); K1->mvpMapPoints[i]->tag = i; }
  for (int i = 0; i < k2->n; i++) if (has2[i]) { K2->mvpMapPoints[i] = mpptr(new MapPoint); K2->mvpMapPoints[i]->tag = i; }
  ORBmatcher matcher(nnratio, check_ori != 0);
  std::vector<mpptr> out;
  const int n = matcher.SearchByBoW(K1, K2, out);
  for (int i = 0; i < k1->n; i++) match12[i] = out[i] ? out[i]->tag : -1;
  return n;
}

// SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs); keyframe 1's camera centre = -R1^T t1, keyframe 2's pose from k2
int ref_match_triangulation(const ref_image* k1, const uint8_t* has1, const ref_image* k2, const uint8_t* has2, const float* F12, int check_ori,
                            int32_t* pairs) {
  kfptr K1 = make_kf(*k1), K2 = make_kf(*k2);
  for (int i = 0; i < k1->n; i++) if (has1[i]) K1->mvpMapPoints[i] = holder(1);
  for (int i = 0; i < k2->n; i++) if (has2[i]) K2->mvpMapPoints[i] = holder(1);
  ORBmatcher matcher(0.6f, check_ori != 0);
  std::vector<std::pair<size_t, size_t> > v;
  const int n = matcher.SearchForTriangulation(K1, K2, mat_f32(F12, 3, 3), v);
  for (size_t i = 0; i < v.size(); i++) { pairs[2 * i] = (int32_t)v[i].first; pairs[2 * i + 1] = (int32_t)v[i].second; }
  return n;
}

// SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)
int ref_search_for_initialization(const ref_image* f1, const ref_image* f2, float* prev_matched /*2 per F1 keypoint, in/out*/, int window, float nnratio,
                                  int check_ori, int32_t* match12) {
  frameptr F2 = make_frame(*f2);                     // the grid statics belong to the frame that is searched
  Frame F1;
  F1.N = f1->n; F1.mvKeysUn = keypoints(*f1); F1.mvKeys = F1.mvKeysUn; F1.mDescriptors = mat_u8(f1->desc, f1->n, 32);
  std::vector<cv::Point2f> prev(f1->n);
  for (int i = 0; i < f1->n; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
  std::vector<int> m;
  ORBmatcher matcher(nnratio, check_ori != 0);
  const int n = matcher.SearchForInitialization(F1, *F2, prev, m, window);
  for (int i = 0; i < f1->n; i++) { match12[i] = m[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
  return n;
}

// SearchByProjection(Frame&, const vector<mpptr>&, th): feat_blocked[j] = the frame's keypoint j already holds a point with observations
int ref_search_by_projection_track(const ref_image* fr, const ref_points* pts, const uint8_t* feat_blocked, float th, float nnratio, int32_t* match_of_feat) {
  frameptr F = make_frame(*fr);
  for (int j = 0; j < fr->n; j++) if (feat_blocked[j]) F->mvpMapPoints[j] = holder(1);
  std::vector<mpptr> P = make_points(*pts, nullptr);
  ORBmatcher matcher(nnratio, true);
  const int n = matcher.SearchByProjection(*F, P, th);
  for (int j = 0; j < fr->n; j++) match_of_feat[j] = (F->mvpMapPoints[j] && F->mvpMapPoints[j]->tag >= 0) ? F->mvpMapPoints[j]->tag : -1;
  return n;
}


// ---- the overloads with a geometric prelude.  Keypoints of the keyframe that already hold a map point get a placeholder whose tag
// encodes the keypoint (-1000 - j), so that Replace / vpReplacePoint outcomes can be read back as keypoint indices. ----------------
static mpptr kf_holder(int j, int n_obs) { mpptr m(new MapPoint); m->tag = -1000 - j; m->nObs = n_obs; return m; }
static int holder_feat(const mpptr& m) { return (m && m->tag <= -1000) ? -1000 - m->tag : -1; }

// Fuse(pKF, vpMapPoints, th): out best_idx[i] = keypoint the reference fused point i with (-1: not fused); returns nFused.
// kf_mp_obs[j] >= 0: keypoint j holds a map point with that many observations, -1: none
int ref_fuse(const ref_image* kf, const int32_t* kf_mp_obs, const ref_points* pts, float th, int32_t* best_idx) {
  kfptr K = make_kf(*kf);
  for (int j = 0; j < kf->n; j++) if (kf_mp_obs[j] >= 0) K->mvpMapPoints[j] = kf_holder(j, kf_mp_obs[j]);
  std::vector<mpptr> held = K->mvpMapPoints;
  std::vector<mpptr> P = make_points(*pts, K.get());
  ORBmatcher matcher(0.6f, true);
  const int n = matcher.Fuse(K, P, th);
  // where a map point sits in the keyframe: its placeholder's keypoint, or the keypoint it was added at
  auto sits = [&](const mpptr& x) { return x->tag <= -1000 ? holder_feat(x) : (x->added.empty() ? -1 : (int)x->added.back().second); };
  for (int i = 0; i < pts->m; i++) {
    best_idx[i] = -1;
    if (!P[i]->added.empty()) best_idx[i] = (int)P[i]->added.back().second;        // AddObservation(pKF, bestIdx)
    else if (P[i]->replacedBy) best_idx[i] = sits(P[i]->replacedBy);                // pMP->Replace(pMPinKF)
  }
  std::vector<mpptr> all = held;                                                    // pMPinKF->Replace(pMP): pMPinKF is a placeholder or an
  all.insert(all.end(), P.begin(), P.end());                                        // earlier point of the list that was added to the keyframe
  for (size_t k = 0; k < all.size(); k++)
    if (all[k] && all[k]->replacedBy && all[k]->replacedBy->tag >= 0 && all[k]->replacedBy->added.empty() && !all[k]->replacedBy->replacedBy) {
      const int i = all[k]->replacedBy->tag;
      if (best_idx[i] < 0) best_idx[i] = sits(all[k]);
    }
  return n;
}

// Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)
int ref_fuse_sim3(const ref_image* kf, const int32_t* kf_mp_obs, const float* Scw /*16*/, const ref_points* pts, float th, int32_t* best_idx) {
  kfptr K = make_kf(*kf);
  for (int j = 0; j < kf->n; j++) if (kf_mp_obs[j] >= 0) K->mvpMapPoints[j] = kf_holder(j, kf_mp_obs[j]);
  std::vector<mpptr> P = make_points(*pts, K.get());
  std::vector<mpptr> repl(pts->m);
  ORBmatcher matcher(0.6f, true);
  const int n = matcher.Fuse(K, mat_f32(Scw, 4, 4), P, th, repl);
  for (int i = 0; i < pts->m; i++) {
    best_idx[i] = -1;
    // vpReplacePoint[i] is whatever sat at the keypoint: a placeholder, or an earlier point of the list that was added there
    if (repl[i]) best_idx[i] = repl[i]->tag <= -1000 ? holder_feat(repl[i]) : (repl[i]->added.empty() ? -1 : (int)repl[i]->added.back().second);
    else if (!P[i]->added.empty()) best_idx[i] = (int)P[i]->added.back().second;
  }
  return n;
}

// SearchByProjection(pKF, Scw, vpPoints, vpMatched, th): feat_matched[j] = vpMatched[j] != nullptr on entry.
// out: match_of_feat[j] = point newly written to vpMatched[j] (-1 otherwise); remap[3*k..] = (point, idx_now, idx_new); returns nmatches
int ref_search_by_projection_sim3(const ref_image* kf, const float* Scw, const ref_points* pts, const uint8_t* feat_matched, int th,
                                  int32_t* match_of_feat, int32_t* remap, int32_t* n_remap) {
  kfptr K = make_kf(*kf);
  std::vector<mpptr> P = make_points(*pts, K.get());
  for (int i = 0; i < pts->m; i++)                                 // a point the keyframe observes sits at that keypoint
    if (pts->index_in_kf && pts->index_in_kf[i] >= 0) K->mvpMapPoints[pts->index_in_kf[i]] = P[i];
  std::vector<mpptr> matched(kf->n);
  for (int j = 0; j < kf->n; j++) if (feat_matched[j]) matched[j] = holder(1);
  ORBmatcher matcher(0.75f, true);
  const int n = matcher.SearchByProjection(K, mat_f32(Scw, 4, 4), P, matched, th);
  for (int j = 0; j < kf->n; j++) match_of_feat[j] = (matched[j] && matched[j]->tag >= 0) ? matched[j]->tag : -1;
  *n_remap = (int)K->remapped.size() / 3;
  for (size_t k = 0; k < K->remapped.size(); k++) remap[k] = K->remapped[k];
  return n;
}

// SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th): p1_of_feat / p2_of_feat = map point index held by each keypoint (-1 none)
int ref_search_by_sim3(const ref_image* k1, const ref_image* k2, const ref_points* pts1, const int32_t* p1_of_feat, const ref_points* pts2,
                       const int32_t* p2_of_feat, float s12, const float* R12, const float* t12, float th, int32_t* match12) {
  kfptr K1 = make_kf(*k1), K2 = make_kf(*k2);
  std::vector<mpptr> P1 = make_points(*pts1, nullptr), P2 = make_points(*pts2, nullptr);
  for (int j = 0; j < k1->n; j++) if (p1_of_feat[j] >= 0) K1->mvpMapPoints[j] = P1[p1_of_feat[j]];
  for (int j = 0; j < k2->n; j++) if (p2_of_feat[j] >= 0) { K2->mvpMapPoints[j] = P2[p2_of_feat[j]]; P2[p2_of_feat[j]]->tag = 100000 + j; }
  std::vector<mpptr> m12(k1->n);
  ORBmatcher matcher(0.75f, true);
  const int n = matcher.SearchBySim3(K1, K2, m12, s12, mat_f32(R12, 3, 3), mat_f32(t12, 3, 1), th);
  for (int j = 0; j < k1->n; j++) match12[j] = m12[j] ? m12[j]->tag - 100000 : -1;   // the keypoint of KF2 whose point was matched
  return n;
}

// SearchByProjection(CurrentFrame, LastFrame, th): last-frame keypoint i holds point last_point[i] (-1 none); last_outlier flags
int ref_search_by_projection_last(const ref_image* cur, const ref_image* last, const ref_points* pts, const int32_t* last_point,
                                  const uint8_t* last_outlier, const uint8_t* feat_blocked, float th, int check_ori, int32_t* match_of_feat) {
  frameptr L = make_frame(*last);
  frameptr F = make_frame(*cur);                      // the statics (bounds, intrinsics, grid) must be the current frame's: build it last
  std::vector<mpptr> P = make_points(*pts, nullptr);
  for (int i = 0; i < last->n; i++) { if (last_point[i] >= 0) { L->mvpMapPoints[i] = P[last_point[i]]; P[last_point[i]]->tag = i; } L->mvbOutlier[i] = last_outlier[i] != 0; }
  for (int j = 0; j < cur->n; j++) if (feat_blocked[j]) F->mvpMapPoints[j] = holder(1);
  ORBmatcher matcher(0.9f, check_ori != 0);
  const int n = matcher.SearchByProjection(*F, *L, th);
  for (int j = 0; j < cur->n; j++) {
    const mpptr& m = F->mvpMapPoints[j];
    match_of_feat[j] = (m && m->tag >= 0) ? m->tag : -1;   // tag = the last-frame keypoint the point came from; cleared = null = -1
  }
  return n;
}

// SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist): keyframe keypoint i holds point kf_point[i] (-1 none)
int ref_search_by_projection_reloc(const ref_image* cur, const ref_image* kf, const ref_points* pts, const int32_t* kf_point,
                                   const uint8_t* already_found, const uint8_t* feat_blocked, float th, int orb_dist, int check_ori,
                                   int32_t* match_of_feat) {
  kfptr K = make_kf(*kf);
  frameptr F = make_frame(*cur);
  std::vector<mpptr> P = make_points(*pts, nullptr);
  std::set<mpptr> found;
  for (int i = 0; i < kf->n; i++) if (kf_point[i] >= 0) { K->mvpMapPoints[i] = P[kf_point[i]]; P[kf_point[i]]->tag = i; }
  for (int i = 0; i < pts->m; i++) if (already_found[i]) found.insert(P[i]);
  for (int j = 0; j < cur->n; j++) if (feat_blocked[j]) F->mvpMapPoints[j] = holder(1);
  ORBmatcher matcher(0.9f, check_ori != 0);
  const int n = matcher.SearchByProjection(*F, K, found, th, orb_dist);
  for (int j = 0; j < cur->n; j++) {
    const mpptr& m = F->mvpMapPoints[j];
    match_of_feat[j] = (m && m->tag >= 0) ? m->tag : -1;
  }
  return n;
}

}  // extern "C"
