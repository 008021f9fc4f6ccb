// orb_oracle.cpp — CPU oracle for the ORB extractor (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Restates cslam::ORBextractor (S/ORBextractor.cpp) and the OpenCV primitives it calls, without OpenCV:
//   ORBextractor::ORBextractor ........ S/ORBextractor.cpp:579-639   (scale tables, per-level quotas, umax)
//   ComputePyramid .................... S/ORBextractor.cpp:1280-1304 (chained cv::resize INTER_LINEAR)
//   ComputeKeyPointsOctTree ........... S/ORBextractor.cpp:933-1024  (30 px cells, FAST 20 -> 7 fallback)
//   DistributeOctTree / DivideNode .... S/ORBextractor.cpp:650-931
//   IC_Angle .......................... S/ORBextractor.cpp:68-95
//   computeOrbDescriptor .............. S/ORBextractor.cpp:98-316
//   operator() ........................ S/ORBextractor.cpp:1216-1278
//   cv::FAST 9/16 + cornerScore<16> ... OCV/features2d/src/fast.cpp:55-250, fast_score.cpp:50-200
//   cv::resize INTER_LINEAR u8 ........ OCV/imgproc/src/imgwarp.cpp (fixed point, 11-bit coefficients)
//   cv::GaussianBlur 7x7 sigma 2 u8 ... OpenCV 4.x fixed-point path (taps 18 34 48 56 48 34 18, 8.8 -> 16.16)
//                                       or 2.4.13 taps (18 34 49 55 49 34 18) with cfg.blur_2413
//   cv::fastAtan2 ..................... OCV/core/src/mathfuncs.cpp:51-77
// Pinned against Python cv2 4.13 in tests/test_oracle_orb.py (the reference ships no tests for this path).
//
// Two implementation-defined spots of the reference are pinned here and documented in DESIGN.md:
//  * DistributeOctTree sorts (size, node pointer) pairs; ties on size are therefore ordered by heap address.  The
//    oracle orders ties by node creation order (monotone allocation).
//  * x*b + y*a in computeOrbDescriptor is evaluated without FMA contraction; cos/sin are the float libm functions.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <vector>

#include "orb_pattern.h"

namespace {

const int PATCH_SIZE = 31;
const int HALF_PATCH_SIZE = 15;
const int EDGE_THRESHOLD = 19;

inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round_d(double v) { return (int)lrint(v); }

struct Img {
  int w = 0, h = 0;
  std::vector<uint8_t> d;
  uint8_t at(int y, int x) const { return d[(size_t)y * w + x]; }
};

struct KP { float x, y, size, angle, response; int octave; };

// ---- cv::resize INTER_LINEAR, 8-bit, fixed point ----------------------------------------------------------------
void resize_linear(const Img& src, Img& dst, int dw, int dh) {
  dst.w = dw; dst.h = dh; dst.d.assign((size_t)dw * dh, 0);
  const double sx = (double)src.w / dw, sy = (double)src.h / dh;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> xa(2 * dw), ya(2 * dh);
  auto coeffs = [](int d, double scale, int n, int& s, short& a0, short& a1) {
    float f = (float)((d + 0.5) * scale - 0.5);
    s = (int)std::floor(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= n - 1) { s = n - 1; f = 0.f; }
    auto sat = [](int v) { return (short)std::max(-32768, std::min(32767, v)); };
    a0 = sat(cv_round((1.f - f) * 2048.f));
    a1 = sat(cv_round(f * 2048.f));
  };
  for (int x = 0; x < dw; x++) coeffs(x, sx, src.w, xofs[x], xa[2 * x], xa[2 * x + 1]);
  for (int y = 0; y < dh; y++) coeffs(y, sy, src.h, yofs[y], ya[2 * y], ya[2 * y + 1]);
  std::vector<int> r0(dw), r1(dw);
  for (int y = 0; y < dh; y++) {
    const int s0 = yofs[y], s1 = std::min(s0 + 1, src.h - 1);
    for (int x = 0; x < dw; x++) {
      const int x0 = xofs[x], x1 = std::min(x0 + 1, src.w - 1);
      r0[x] = src.at(s0, x0) * xa[2 * x] + src.at(s0, x1) * xa[2 * x + 1];
      r1[x] = src.at(s1, x0) * xa[2 * x] + src.at(s1, x1) * xa[2 * x + 1];
    }
    const int b0 = ya[2 * y], b1 = ya[2 * y + 1];
    for (int x = 0; x < dw; x++)
      dst.d[(size_t)y * dw + x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
  }
}

inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    else p = 2 * (n - 1) - p;
  }
  return p;
}

// ---- cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on u8 ----------------------------------------------------
void gaussian_blur7(const Img& src, Img& dst, bool taps2413) {
  static const int k4[7] = {18, 34, 48, 56, 48, 34, 18};
  static const int k2[7] = {18, 34, 49, 55, 49, 34, 18};
  const int* k = taps2413 ? k2 : k4;
  dst.w = src.w; dst.h = src.h; dst.d.assign(src.d.size(), 0);
  std::vector<uint32_t> row((size_t)src.w * src.h);
  for (int y = 0; y < src.h; y++)
    for (int x = 0; x < src.w; x++) {
      uint32_t s = 0;
      for (int i = -3; i <= 3; i++) s += (uint32_t)k[i + 3] * src.at(y, reflect101(x + i, src.w));
      row[(size_t)y * src.w + x] = s;
    }
  for (int y = 0; y < src.h; y++)
    for (int x = 0; x < src.w; x++) {
      uint32_t s = 0;
      for (int j = -3; j <= 3; j++) s += (uint32_t)k[j + 3] * row[(size_t)reflect101(y + j, src.h) * src.w + x];
      uint32_t v = (s + 32768u) >> 16;
      dst.d[(size_t)y * src.w + x] = (uint8_t)std::min<uint32_t>(v, 255u);
    }
}

// ---- cv::FAST (TYPE_9_16, nonmaxSuppression = true) on a sub-rectangle --------------------------------------------
static const int kRing[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                                 {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// cornerScore<16>: the largest t for which the pixel is still a 9/16 corner (-1 .. 254)
int corner_score(const Img& im, int x, int y) {
  int d[25];
  const int v = im.at(y, x);
  for (int k = 0; k < 25; k++) d[k] = v - im.at(y + kRing[k % 16][1], x + kRing[k % 16][0]);
  int best = -1000;
  for (int k = 0; k < 16; k++) {
    int mn = d[k], mx = d[k];
    for (int i = 1; i < 9; i++) { mn = std::min(mn, d[k + i]); mx = std::max(mx, d[k + i]); }
    best = std::max(best, std::max(mn, -mx));
  }
  return best - 1;
}

// keypoints of cv::FAST(img(rows [y0,y1), cols [x0,x1)), threshold, nms=true): ROI-relative coords, row-major order
void fast_roi(const Img& im, int x0, int y0, int x1, int y1, int threshold, std::vector<KP>& out) {
  const int cols = x1 - x0, rows = y1 - y0;
  if (cols < 7 || rows < 7) return;
  std::vector<int> score((size_t)cols * rows, 0);
  for (int i = 3; i < rows - 3; i++)
    for (int j = 3; j < cols - 3; j++) {
      const int s = corner_score(im, x0 + j, y0 + i);
      if (s >= threshold) score[(size_t)i * cols + j] = s;  // corner at this threshold; buffer holds its score
    }
  for (int i = 3; i < rows - 3; i++)
    for (int j = 3; j < cols - 3; j++) {
      const int s = score[(size_t)i * cols + j];
      if (s < threshold) continue;
      bool mx = true;
      for (int dy = -1; dy <= 1 && mx; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          if (!dx && !dy) continue;
          if (!(s > score[(size_t)(i + dy) * cols + (j + dx)])) { mx = false; break; }
        }
      if (mx) out.push_back(KP{(float)j, (float)i, 7.f, -1.f, (float)s, 0});
    }
}

float fast_atan2(float y, float x) {  // OCV/core/src/mathfuncs.cpp:51-77, scalar, no FMA
  static const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
  static const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
  static const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
  static const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
  float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

struct Extractor {
  int nfeatures, nlevels, iniTh, minTh;
  float scaleFactor;
  bool blur2413;
  std::vector<float> scale, invScale;
  std::vector<int> nPerLevel, umax;
  std::vector<Img> pyr;

  Extractor(int nf, float sf, int nl, int ini, int mn, bool b2413)
      : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf), blur2413(b2413) {
    scale.resize(nl); invScale.resize(nl);
    scale[0] = 1.0f;
    for (int i = 1; i < nl; i++) scale[i] = scale[i - 1] * sf;
    for (int i = 0; i < nl; i++) invScale[i] = 1.0f / scale[i];
    nPerLevel.resize(nl);
    float factor = 1.0f / sf;
    float nDesired = nf * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
      nPerLevel[l] = cv_round(nDesired);
      sum += nPerLevel[l];
      nDesired *= factor;
    }
    nPerLevel[nl - 1] = std::max(nf - sum, 0);
    umax.resize(HALF_PATCH_SIZE + 1);
    int v, v0, vmax = (int)std::floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) umax[v] = cv_round_d(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
  }

  void compute_pyramid(const Img& image) {
    pyr.resize(nlevels);
    for (int l = 0; l < nlevels; l++) {
      const float s = invScale[l];
      const int w = cv_round((float)image.w * s), h = cv_round((float)image.h * s);
      if (l == 0) pyr[0] = image;
      else resize_linear(pyr[l - 1], pyr[l], w, h);
    }
  }

  float ic_angle(const Img& im, float px, float py, int* m01_out = nullptr, int* m10_out = nullptr) const {
    int m_01 = 0, m_10 = 0;
    const int cx = cv_round(px), cy = cv_round(py);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * im.at(cy, cx + u);
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
      int v_sum = 0;
      const int d = umax[v];
      for (int u = -d; u <= d; ++u) {
        const int val_plus = im.at(cy + v, cx + u), val_minus = im.at(cy - v, cx + u);
        v_sum += (val_plus - val_minus);
        m_10 += u * (val_plus + val_minus);
      }
      m_01 += v * v_sum;
    }
    if (m01_out) *m01_out = m_01;
    if (m10_out) *m10_out = m_10;
    return fast_atan2((float)m_01, (float)m_10);
  }

  static void descriptor(const Img& im, float kx, float ky, float kangle, uint8_t* desc) {
    const float factorPI = (float)(M_PI / 180.f);
    const float angle = kangle * factorPI;
    const float a = cosf(angle), b = sinf(angle);
    const int cx = cv_round(kx), cy = cv_round(ky);
    auto get = [&](int idx) {
      const float x = (float)kOrcOrbPattern[2 * idx], y = (float)kOrcOrbPattern[2 * idx + 1];
      volatile float xb = x * b, ya = y * a, xa = x * a, yb = y * b;  // volatile: no contraction into FMA
      const int iy = cv_round(xb + ya), ix = cv_round(xa - yb);
      return (int)im.at(cy + iy, cx + ix);
    };
    for (int i = 0; i < 32; i++) {
      int val = 0;
      for (int bit = 0; bit < 8; bit++) {
        const int t0 = get(16 * i + 2 * bit), t1 = get(16 * i + 2 * bit + 1);
        val |= (t0 < t1) << bit;
      }
      desc[i] = (uint8_t)val;
    }
  }

  struct Node {
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::vector<KP> keys;
    bool noMore = false;
    std::list<Node>::iterator lit;
    long seq = 0;
  };

  static void divide(const Node& n, Node& n1, Node& n2, Node& n3, Node& n4) {
    const int halfX = (int)std::ceil(static_cast<float>(n.URx - n.ULx) / 2);
    const int halfY = (int)std::ceil(static_cast<float>(n.BRy - n.ULy) / 2);
    n1.ULx = n.ULx; n1.ULy = n.ULy; n1.URx = n.ULx + halfX; n1.URy = n.ULy;
    n1.BLx = n.ULx; n1.BLy = n.ULy + halfY; n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = n.URx; n2.URy = n.URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = n.BLx; n3.BLy = n.BLy; n3.BRx = n1.BRx; n3.BRy = n.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = n.BRx; n4.BRy = n.BRy;
    for (const KP& kp : n.keys) {
      if (kp.x < n1.URx) {
        if (kp.y < n1.BRy) n1.keys.push_back(kp); else n3.keys.push_back(kp);
      } else if (kp.y < n1.BRy) n2.keys.push_back(kp);
      else n4.keys.push_back(kp);
    }
    if (n1.keys.size() == 1) n1.noMore = true;
    if (n2.keys.size() == 1) n2.noMore = true;
    if (n3.keys.size() == 1) n3.noMore = true;
    if (n4.keys.size() == 1) n4.noMore = true;
  }

  std::vector<KP> distribute(const std::vector<KP>& cand, int minX, int maxX, int minY, int maxY, int N) const {
    const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<Node> nodes;
    std::vector<Node*> ini(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; i++) {
      Node ni;
      ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
      ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
      ni.BLx = ni.ULx; ni.BLy = maxY - minY;
      ni.BRx = ni.URx; ni.BRy = maxY - minY;
      ni.seq = seq++;
      nodes.push_back(ni);
      ini[i] = &nodes.back();
    }
    for (const KP& kp : cand) ini[(int)(kp.x / hX)]->keys.push_back(kp);
    for (auto it = nodes.begin(); it != nodes.end();) {
      if (it->keys.size() == 1) { it->noMore = true; ++it; }
      else if (it->keys.empty()) it = nodes.erase(it);
      else ++it;
    }
    bool finish = false;
    typedef std::pair<int, Node*> SP;
    auto cmp = [](const SP& a, const SP& b) { return a.first != b.first ? a.first < b.first : a.second->seq < b.second->seq; };
    std::vector<SP> sizeAndNode;
    auto push_children = [&](Node* kids[4], int& nToExpand) {
      for (int c = 0; c < 4; c++) {
        Node& ch = *kids[c];
        if (ch.keys.size() > 0) {
          ch.seq = seq++;
          nodes.push_front(ch);
          if (ch.keys.size() > 1) {
            nToExpand++;
            sizeAndNode.push_back(std::make_pair((int)ch.keys.size(), &nodes.front()));
            nodes.front().lit = nodes.begin();
          }
        }
      }
    };
    while (!finish) {
      const int prevSize = (int)nodes.size();
      auto lit = nodes.begin();
      int nToExpand = 0;
      sizeAndNode.clear();
      while (lit != nodes.end()) {
        if (lit->noMore) { ++lit; continue; }
        Node n1, n2, n3, n4;
        divide(*lit, n1, n2, n3, n4);
        Node* kids[4] = {&n1, &n2, &n3, &n4};
        push_children(kids, nToExpand);
        lit = nodes.erase(lit);
      }
      if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
        finish = true;
      } else if (((int)nodes.size() + nToExpand * 3) > N) {
        while (!finish) {
          const int prev2 = (int)nodes.size();
          std::vector<SP> prevList = sizeAndNode;
          sizeAndNode.clear();
          std::sort(prevList.begin(), prevList.end(), cmp);
          for (int j = (int)prevList.size() - 1; j >= 0; j--) {
            Node n1, n2, n3, n4;
            divide(*prevList[j].second, n1, n2, n3, n4);
            Node* kids[4] = {&n1, &n2, &n3, &n4};
            int dummy = 0;
            push_children(kids, dummy);
            nodes.erase(prevList[j].second->lit);
            if ((int)nodes.size() >= N) break;
          }
          if ((int)nodes.size() >= N || (int)nodes.size() == prev2) finish = true;
        }
      }
    }
    std::vector<KP> res;
    for (auto& nd : nodes) {
      const KP* best = &nd.keys[0];
      float maxResp = best->response;
      for (size_t k = 1; k < nd.keys.size(); k++)
        if (nd.keys[k].response > maxResp) { best = &nd.keys[k]; maxResp = nd.keys[k].response; }
      res.push_back(*best);
    }
    return res;
  }

  // per-level candidates in the reference's order (coords relative to minBorder)
  void level_candidates(int level, std::vector<KP>& cand, int& minBX, int& maxBX, int& minBY, int& maxBY) const {
    const Img& im = pyr[level];
    const float W = 30;
    minBX = EDGE_THRESHOLD - 3; minBY = minBX;
    maxBX = im.w - EDGE_THRESHOLD + 3; maxBY = im.h - EDGE_THRESHOLD + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    for (int i = 0; i < nRows; i++) {
      const float iniY = (float)(minBY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < nCols; j++) {
        const float iniX = (float)(minBX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        std::vector<KP> cell;
        fast_roi(im, (int)iniX, (int)iniY, (int)maxX, (int)maxY, iniTh, cell);
        if (cell.empty()) fast_roi(im, (int)iniX, (int)iniY, (int)maxX, (int)maxY, minTh, cell);
        for (KP& k : cell) {
          k.x += j * wCell;
          k.y += i * hCell;
          cand.push_back(k);
        }
      }
    }
  }

  void extract(const Img& image, std::vector<KP>& kps, std::vector<uint8_t>& desc) {
    compute_pyramid(image);
    std::vector<std::vector<KP>> all(nlevels);
    for (int l = 0; l < nlevels; l++) {
      std::vector<KP> cand;
      int minBX, maxBX, minBY, maxBY;
      level_candidates(l, cand, minBX, maxBX, minBY, maxBY);
      all[l] = distribute(cand, minBX, maxBX, minBY, maxBY, nPerLevel[l]);
      const int scaledPatch = (int)(PATCH_SIZE * scale[l]);
      for (KP& k : all[l]) { k.x += minBX; k.y += minBY; k.octave = l; k.size = (float)scaledPatch; }
    }
    for (int l = 0; l < nlevels; l++)
      for (KP& k : all[l]) k.angle = ic_angle(pyr[l], k.x, k.y);
    kps.clear(); desc.clear();
    for (int l = 0; l < nlevels; l++) {
      if (all[l].empty()) continue;
      Img blurred;
      gaussian_blur7(pyr[l], blurred, blur2413);
      const size_t off = desc.size();
      desc.resize(off + all[l].size() * 32);
      for (size_t i = 0; i < all[l].size(); i++) descriptor(blurred, all[l][i].x, all[l][i].y, all[l][i].angle, &desc[off + 32 * i]);
      if (l != 0) for (KP& k : all[l]) { k.x *= scale[l]; k.y *= scale[l]; }
      kps.insert(kps.end(), all[l].begin(), all[l].end());
    }
  }
};

Img make_img(const uint8_t* p, int w, int h, int stride) {
  Img im; im.w = w; im.h = h; im.d.resize((size_t)w * h);
  for (int y = 0; y < h; y++) memcpy(&im.d[(size_t)y * w], p + (size_t)y * stride, w);
  return im;
}

}  // namespace

extern "C" {

struct orc_orb_config { int32_t nfeatures; float scale_factor; int32_t nlevels, ini_th_fast, min_th_fast, blur_2413; };
struct orc_keypoint { float x, y, size, angle, response; int32_t octave; };

int orc_orb_extract(const uint8_t* img, int w, int h, int stride, const orc_orb_config* c, orc_keypoint* kps, int max_kp,
                    int* n, uint8_t* desc) {
  Extractor ex(c->nfeatures, c->scale_factor, c->nlevels, c->ini_th_fast, c->min_th_fast, c->blur_2413 != 0);
  std::vector<KP> k; std::vector<uint8_t> d;
  ex.extract(make_img(img, w, h, stride), k, d);
  const int cnt = std::min((int)k.size(), max_kp);
  for (int i = 0; i < cnt; i++) kps[i] = orc_keypoint{k[i].x, k[i].y, k[i].size, k[i].angle, k[i].response, k[i].octave};
  memcpy(desc, d.data(), (size_t)cnt * 32);
  *n = cnt;
  return 0;
}

// quotas / umax / level sizes for known-answer checks
void orc_orb_tables(const orc_orb_config* c, int w, int h, int* n_per_level, int* umax16, int* level_wh) {
  Extractor ex(c->nfeatures, c->scale_factor, c->nlevels, c->ini_th_fast, c->min_th_fast, false);
  for (int l = 0; l < c->nlevels; l++) {
    n_per_level[l] = ex.nPerLevel[l];
    level_wh[2 * l] = cv_round((float)w * ex.invScale[l]);
    level_wh[2 * l + 1] = cv_round((float)h * ex.invScale[l]);
  }
  for (int i = 0; i < 16; i++) umax16[i] = ex.umax[i];
}

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  Img s = make_img(src, sw, sh, sw), d;
  resize_linear(s, d, dw, dh);
  memcpy(dst, d.d.data(), d.d.size());
}

void orc_gaussian_blur7(const uint8_t* src, int w, int h, uint8_t* dst, int taps2413) {
  Img s = make_img(src, w, h, w), d;
  gaussian_blur7(s, d, taps2413 != 0);
  memcpy(dst, d.d.data(), d.d.size());
}

// cv::FAST on a whole image: returns count; xy (2 per kp), score
int orc_fast(const uint8_t* img, int w, int h, int threshold, int* xy, int* score, int max_out) {
  Img s = make_img(img, w, h, w);
  std::vector<KP> out;
  fast_roi(s, 0, 0, w, h, threshold, out);
  const int n = std::min((int)out.size(), max_out);
  for (int i = 0; i < n; i++) { xy[2 * i] = (int)out[i].x; xy[2 * i + 1] = (int)out[i].y; score[i] = (int)out[i].response; }
  return (int)out.size();
}

float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }

// per-level candidate list (before the quadtree), level coords relative to minBorder — for the GPU FAST kernel parity test
int orc_orb_level_candidates(const uint8_t* img, int w, int h, const orc_orb_config* c, int level, float* xys /*3 per cand*/, int max_out) {
  Extractor ex(c->nfeatures, c->scale_factor, c->nlevels, c->ini_th_fast, c->min_th_fast, false);
  ex.compute_pyramid(make_img(img, w, h, w));
  std::vector<KP> cand; int a, b, cc, d;
  ex.level_candidates(level, cand, a, b, cc, d);
  const int n = std::min((int)cand.size(), max_out);
  for (int i = 0; i < n; i++) { xys[3 * i] = cand[i].x; xys[3 * i + 1] = cand[i].y; xys[3 * i + 2] = cand[i].response; }
  return (int)cand.size();
}

void orc_orb_descriptor(const uint8_t* img, int w, int h, float x, float y, float angle, uint8_t* desc32) {
  Img s = make_img(img, w, h, w);
  Extractor::descriptor(s, x, y, angle, desc32);
}

float orc_ic_angle(const uint8_t* img, int w, int h, float x, float y, int* m01, int* m10) {
  orc_orb_config c{1000, 1.2f, 8, 20, 7, 0};
  Extractor ex(c.nfeatures, c.scale_factor, c.nlevels, c.ini_th_fast, c.min_th_fast, false);
  Img s = make_img(img, w, h, w);
  return ex.ic_angle(s, x, y, m01, m10);
}

}  // extern "C"
