// match_oracle.cpp — CPU oracle for ORB descriptor matching (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// Restates, on flat arrays, the three matchers BASELINE.json names:
//   ORBmatcher::SearchByBoW(kfptr, Frame&, ...)        S/ORBmatcher.cpp:178-306  (accept best <= TH_LOW)
//   ORBmatcher::SearchByBoW(kfptr, kfptr, ...)         S/ORBmatcher.cpp:565-698  (accept best <  TH_LOW, vbMatched2)
//   ORBmatcher::SearchForTriangulation(...)            S/ORBmatcher.cpp:700-852  (+ CheckDistEpipolarLine :159-176)
//   ORBmatcher::ComputeThreeMaxima                     S/ORBmatcher.cpp:1607-1648
//   ORBmatcher::DescriptorDistance                     S/ORBmatcher.cpp:1653-1669 (SWAR popcount over 8 x 32 bit)
// TH_LOW = 50, HISTO_LENGTH = 30 (S/ORBmatcher.cpp:63-65).  The rotation histogram keeps the reference's
// bin = round(rot * (1/30)) quirk (only bins 0..12 are ever hit).  The DBoW2 FeatureVector (std::map<NodeId,
// vector<unsigned>>) arrives flattened with ascending node ids; the merge-join with lower_bound is equivalent to a
// two-pointer walk.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
const int TH_LOW = 50;
const int HISTO_LENGTH = 30;

int descriptor_distance(const uint8_t* a, const uint8_t* b) {
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

void three_maxima(std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
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

struct orc_fv { int32_t n_nodes; const uint32_t* node_id; const int32_t* node_ptr; const uint32_t* feat; };

int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// out: match_kf_of_f[n_f] = KF feature whose MapPoint was assigned to frame feature, -1 otherwise
int orc_match_bow_kf_frame(const uint8_t* desc_kf, int n_kf, const uint8_t* kf_has_mp, const float* angle_kf, const orc_fv* fk,
                           const uint8_t* desc_f, int n_f, const float* angle_f, const orc_fv* ff, float nnratio,
                           int check_ori, int* match_kf_of_f) {
  (void)n_kf;
  for (int i = 0; i < n_f; i++) match_kf_of_f[i] = -1;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  int a = 0, b = 0;
  while (a < fk->n_nodes && b < ff->n_nodes) {
    if (fk->node_id[a] == ff->node_id[b]) {
      for (int ik = fk->node_ptr[a]; ik < fk->node_ptr[a + 1]; ik++) {
        const unsigned idxKF = fk->feat[ik];
        if (!kf_has_mp[idxKF]) continue;
        int best1 = 256, bestIdx = -1, best2 = 256;
        for (int jf = ff->node_ptr[b]; jf < ff->node_ptr[b + 1]; jf++) {
          const unsigned idxF = ff->feat[jf];
          if (match_kf_of_f[idxF] >= 0) continue;
          const int dist = descriptor_distance(desc_kf + 32 * (size_t)idxKF, desc_f + 32 * (size_t)idxF);
          if (dist < best1) { best2 = best1; best1 = dist; bestIdx = (int)idxF; }
          else if (dist < best2) best2 = dist;
        }
        if (best1 <= TH_LOW && static_cast<float>(best1) < nnratio * static_cast<float>(best2)) {
          match_kf_of_f[bestIdx] = (int)idxKF;
          if (check_ori) rotHist[rot_bin(angle_kf[idxKF], angle_f[bestIdx])].push_back(bestIdx);
          nmatches++;
        }
      }
      a++; b++;
    } else if (fk->node_id[a] < ff->node_id[b]) {
      while (a < fk->n_nodes && fk->node_id[a] < ff->node_id[b]) a++;
    } else {
      while (b < ff->n_nodes && ff->node_id[b] < fk->node_id[a]) b++;
    }
  }
  if (check_ori) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == i1 || i == i2 || i == i3) continue;
      for (int v : rotHist[i]) { match_kf_of_f[v] = -1; nmatches--; }
    }
  }
  return nmatches;
}

int orc_match_bow_kf_kf(const uint8_t* d1, int n1, const uint8_t* has1, const float* ang1, const orc_fv* f1,
                        const uint8_t* d2, int n2, const uint8_t* has2, const float* ang2, const orc_fv* f2, float nnratio,
                        int check_ori, int* match12) {
  for (int i = 0; i < n1; i++) match12[i] = -1;
  std::vector<char> matched2(n2, 0);
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  int a = 0, b = 0;
  while (a < f1->n_nodes && b < f2->n_nodes) {
    if (f1->node_id[a] == f2->node_id[b]) {
      for (int i1 = f1->node_ptr[a]; i1 < f1->node_ptr[a + 1]; i1++) {
        const unsigned idx1 = f1->feat[i1];
        if (!has1[idx1]) continue;
        int best1 = 256, bestIdx2 = -1, best2 = 256;
        for (int i2 = f2->node_ptr[b]; i2 < f2->node_ptr[b + 1]; i2++) {
          const unsigned idx2 = f2->feat[i2];
          if (matched2[idx2] || !has2[idx2]) continue;
          const int dist = descriptor_distance(d1 + 32 * (size_t)idx1, d2 + 32 * (size_t)idx2);
          if (dist < best1) { best2 = best1; best1 = dist; bestIdx2 = (int)idx2; }
          else if (dist < best2) best2 = dist;
        }
        if (best1 < TH_LOW && static_cast<float>(best1) < nnratio * static_cast<float>(best2)) {
          match12[idx1] = bestIdx2;
          matched2[bestIdx2] = 1;
          if (check_ori) rotHist[rot_bin(ang1[idx1], ang2[bestIdx2])].push_back((int)idx1);
          nmatches++;
        }
      }
      a++; b++;
    } else if (f1->node_id[a] < f2->node_id[b]) {
      while (a < f1->n_nodes && f1->node_id[a] < f2->node_id[b]) a++;
    } else {
      while (b < f2->n_nodes && f2->node_id[b] < f1->node_id[a]) b++;
    }
  }
  if (check_ori) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == i1 || i == i2 || i == i3) continue;
      for (int v : rotHist[i]) { match12[v] = -1; nmatches--; }
    }
  }
  return nmatches;
}

struct orc_tri_view { const uint8_t* desc; int32_t n; const uint8_t* has_mp; const float* kp_xy; const int32_t* octave;
                      const float* angle; const orc_fv* fv; float fx, fy, cx, cy; };

int orc_match_triangulation(const orc_tri_view* v1, const orc_tri_view* v2, const float F12[9], float ex, float ey,
                            const float* level_sigma2, const float* scale_factors, int check_ori, int* pairs) {
  std::vector<char> matched2(v2->n, 0);
  std::vector<int> m12(v1->n, -1);
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const orc_fv *f1 = v1->fv, *f2 = v2->fv;
  int a = 0, b = 0;
  while (a < f1->n_nodes && b < f2->n_nodes) {
    if (f1->node_id[a] == f2->node_id[b]) {
      for (int i1 = f1->node_ptr[a]; i1 < f1->node_ptr[a + 1]; i1++) {
        const unsigned idx1 = f1->feat[i1];
        if (v1->has_mp[idx1]) continue;
        const float k1x = v1->kp_xy[2 * idx1], k1y = v1->kp_xy[2 * idx1 + 1];
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int i2 = f2->node_ptr[b]; i2 < f2->node_ptr[b + 1]; i2++) {
          const unsigned idx2 = f2->feat[i2];
          if (matched2[idx2] || v2->has_mp[idx2]) continue;
          const int dist = descriptor_distance(v1->desc + 32 * (size_t)idx1, v2->desc + 32 * (size_t)idx2);
          if (dist > TH_LOW || dist > bestDist) continue;
          const float k2x = v2->kp_xy[2 * idx2], k2y = v2->kp_xy[2 * idx2 + 1];
          const float distex = ex - k2x, distey = ey - k2y;
          if (distex * distex + distey * distey < 100 * scale_factors[v2->octave[idx2]]) continue;
          // CheckDistEpipolarLine
          const float la = k1x * F12[0] + k1y * F12[3] + F12[6];
          const float lb = k1x * F12[1] + k1y * F12[4] + F12[7];
          const float lc = k1x * F12[2] + k1y * F12[5] + F12[8];
          const float num = la * k2x + lb * k2y + lc;
          const float den = la * la + lb * lb;
          if (den == 0) continue;
          const float dsqr = num * num / den;
          if (dsqr < 3.84 * level_sigma2[v2->octave[idx2]]) { bestIdx2 = (int)idx2; bestDist = dist; }
        }
        if (bestIdx2 >= 0) {
          m12[idx1] = bestIdx2;
          nmatches++;
          if (check_ori) rotHist[rot_bin(v1->angle[idx1], v2->angle[bestIdx2])].push_back((int)idx1);
        }
      }
      a++; b++;
    } else if (f1->node_id[a] < f2->node_id[b]) {
      while (a < f1->n_nodes && f1->node_id[a] < f2->node_id[b]) a++;
    } else {
      while (b < f2->n_nodes && f2->node_id[b] < f1->node_id[a]) b++;
    }
  }
  if (check_ori) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == i1 || i == i2 || i == i3) continue;
      for (int v : rotHist[i]) { m12[v] = -1; nmatches--; }
    }
  }
  int np = 0;
  for (int i = 0; i < v1->n; i++)
    if (m12[i] >= 0) { pairs[2 * np] = i; pairs[2 * np + 1] = m12[i]; np++; }
  return np;
}

}  // extern "C"
