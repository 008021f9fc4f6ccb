// ref_orb_wrap.cpp — command-line program around the REFERENCE's own ORBextractor.cpp (TEST INFRASTRUCTURE, NOT PRODUCT).
//
// oracle/Makefile compiles cslam/src/ORBextractor.cpp where it lies under /root/reference, together with this file, against the
// stand-in headers of oracle/ref_stub/ into the executable oracle/_ref/orb_ref_cli.  Nothing of the reference is copied: its scale tables and
// umax, the pyramid loop, the 30-px FAST cells with the 20 -> 7 threshold fallback, the quadtree distribution (including whatever
// its pointer-ordered sort does with ties), IC_Angle, the rotated BRIEF sampling, the per-level scaling and the output order are
// the reference's object code.  What this file supplies are the five OpenCV primitives that code calls, implemented on the oracle's
// restatements (liboracle.so: orc_fast, orc_resize_linear_u8, orc_gaussian_blur7, orc_fast_atan2) which tests/test_oracle_orb.py
// pins to cv2 4.13 — so tests/test_oracle_vs_reference_orb.py checks the oracle's EXTRACTOR LOGIC against the reference's, on
// identical primitives.
#include <cslam/ORBextractor.h>

#include <sys/mman.h>

#include <cstdint>
#include <cstdio>
#include <new>

// ---- a monotone allocator, switchable at run time --------------------------------------------------------------------------
// DistributeOctTree sorts (size, ExtractorNode*) pairs (ORBextractor.cpp:852): nodes of equal size are split in the order of their
// heap addresses, so with a general-purpose allocator the reference's own output depends on malloc's free lists (measured with glibc:
// the same image gives the same keypoints on most levels in a different order, and 2-3 different keypoints on some).  To compare
// logic with logic, this program can serve every allocation of the process (it is an executable: its operator new / delete replace
// libstdc++'s for all code, consistently) from a bump arena that never reuses memory: addresses then grow with allocation order, i.e.
// equal-size nodes are split latest-created first — the order the oracle documents for itself (orb_oracle.cpp: "ties by node creation
// order (monotone allocation)").  With --malloc the same program runs on glibc's allocator, as a deployed reference would.
namespace {
char* g_arena = nullptr;
size_t g_off = 0;
bool g_bump = false;
const size_t ARENA = (size_t)8 << 30;   // virtual, touched lazily
void* bump(size_t n) {
  if (!g_bump) { void* p = malloc(n ? n : 1); if (!p) abort(); return p; }
  if (!g_arena) {
    g_arena = static_cast<char*>(mmap(nullptr, ARENA, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
    if (g_arena == MAP_FAILED) abort();
  }
  void* p = g_arena + g_off;
  g_off += (n + 15) & ~(size_t)15;
  if (g_off > ARENA) abort();
  return p;
}
void unbump(void* p) {
  if (g_arena && p >= (void*)g_arena && p < (void*)(g_arena + ARENA)) return;   // arena memory is never reused
  free(p);
}
}  // namespace
void* operator new(size_t n) { return bump(n); }
void* operator new[](size_t n) { return bump(n); }
void operator delete(void* p) noexcept { unbump(p); }
void operator delete[](void* p) noexcept { unbump(p); }
void operator delete(void* p, size_t) noexcept { unbump(p); }
void operator delete[](void* p, size_t) noexcept { unbump(p); }

extern "C" {
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
void orc_gaussian_blur7(const uint8_t* src, int w, int h, uint8_t* dst, int taps2413);
int orc_fast(const uint8_t* img, int w, int h, int threshold, int* xy, int* score, int max_out);
float orc_fast_atan2(float y, float x);
}

namespace {
std::vector<uint8_t> packed(const cv::Mat& m) {   // a view's pixels as one contiguous w x h buffer
  std::vector<uint8_t> b((size_t)m.rows * m.cols);
  for (int r = 0; r < m.rows; r++) memcpy(b.data() + (size_t)r * m.cols, m.ptr(r), (size_t)m.cols);
  return b;
}
void unpack(const std::vector<uint8_t>& b, cv::Mat& m) {
  for (int r = 0; r < m.rows; r++) memcpy(m.ptr(r), b.data() + (size_t)r * m.cols, (size_t)m.cols);
}
int reflect101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
int g_blur_2413 = 0;
}  // namespace

namespace cv {

void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
  assert(nonmaxSuppression);
  keypoints.clear();
  if (image.rows < 7 || image.cols < 7) return;
  const std::vector<uint8_t> b = packed(image);
  const int cap = image.rows * image.cols;
  std::vector<int> xy(2 * (size_t)cap), score(cap);
  const int n = orc_fast(b.data(), image.cols, image.rows, threshold, xy.data(), score.data(), cap);
  for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xy[2 * i], (float)xy[2 * i + 1], 7.f, -1.f, (float)score[i]));
}

void resize(const Mat& src, Mat& dst, Size dsize, double, double, int interpolation) {
  assert(interpolation == INTER_LINEAR && src.type() == CV_8UC1);
  const std::vector<uint8_t> s = packed(src);
  std::vector<uint8_t> d((size_t)dsize.width * dsize.height);
  orc_resize_linear_u8(s.data(), src.cols, src.rows, d.data(), dsize.width, dsize.height);
  dst.create(dsize.height, dsize.width, src.type());   // an existing view of that shape is written in place, as OpenCV does
  unpack(d, dst);
}

void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
  assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
  const std::vector<uint8_t> s = packed(src);
  std::vector<uint8_t> d(s.size());
  orc_gaussian_blur7(s.data(), src.cols, src.rows, d.data(), g_blur_2413);
  dst.create(src.rows, src.cols, src.type());
  unpack(d, dst);
}

void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType) {
  assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
  const std::vector<uint8_t> s = packed(src);          // src may be the interior view of dst (ComputePyramid): read it first
  const int h = src.rows, w = src.cols;
  dst.create(h + top + bottom, w + left + right, src.type());
  for (int y = 0; y < dst.rows; y++) {
    const uint8_t* row = s.data() + (size_t)reflect101(y - top, h) * w;
    uchar* out = dst.ptr(y);
    for (int x = 0; x < dst.cols; x++) out[x] = row[reflect101(x - left, w)];
  }
}

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

void KeyPointsFilter::retainBest(std::vector<KeyPoint>& keypoints, int npoints) {   // only reachable from ComputeKeyPointsOld (unused)
  if ((int)keypoints.size() > npoints) keypoints.resize(npoints);
}

}  // namespace cv

extern "C" {

struct ref_keypoint { float x, y, size, angle, response; int32_t octave; };

// (*mpORBextractor)(im, cv::Mat(), mvKeys, mDescriptors) as Frame::ExtractORB calls it (cslam/src/Frame.cpp:120-123)
int ref_orb_extract(const uint8_t* img, int w, int h, int stride, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th,
                    int blur_2413, ref_keypoint* kps, int max_kp, uint8_t* desc) {
  g_blur_2413 = blur_2413;
  cv::Mat image(h, w, CV_8UC1);
  for (int r = 0; r < h; r++) memcpy(image.ptr(r), img + (size_t)r * stride, (size_t)w);
  cslam::ORBextractor ex(nfeatures, scale_factor, nlevels, ini_th, min_th);
  std::vector<cv::KeyPoint> keys;
  cv::Mat descriptors;
  ex(image, cv::Mat(), keys, descriptors);
  const int n = (int)keys.size();
  for (int i = 0; i < n && i < max_kp; i++) {
    kps[i].x = keys[i].pt.x; kps[i].y = keys[i].pt.y; kps[i].size = keys[i].size; kps[i].angle = keys[i].angle;
    kps[i].response = keys[i].response; kps[i].octave = keys[i].octave;
    memcpy(desc + 32 * (size_t)i, descriptors.ptr(i), 32);
  }
  return n;
}

}  // extern "C"

// orb_ref_cli <in.raw> <w> <h> <nfeatures> <scale> <nlevels> <iniTh> <minTh> <blur2413> <bump|malloc> <out.bin>
// in.raw: w*h bytes; out.bin: int32 n, then n x {6 x f32/i32 keypoint}, then n x 32 descriptor bytes
int main(int argc, char** argv) {
  if (argc != 12) { fprintf(stderr, "usage: %s in.raw w h nfeatures scale nlevels iniTh minTh blur2413 bump|malloc out.bin\n", argv[0]); return 2; }
  const int w = atoi(argv[2]), h = atoi(argv[3]);
  g_bump = std::string(argv[10]) == "bump";
  std::vector<uint8_t> img((size_t)w * h);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(img.data(), 1, img.size(), f) != img.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 3; }
  fclose(f);
  const int cap = 20000;
  std::vector<ref_keypoint> kps(cap);
  std::vector<uint8_t> desc((size_t)cap * 32);
  const int32_t n = ref_orb_extract(img.data(), w, h, w, atoi(argv[4]), (float)atof(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]),
                                    atoi(argv[9]), kps.data(), cap, desc.data());
  if (n > cap) return 4;
  f = fopen(argv[11], "wb");
  if (!f) return 5;
  fwrite(&n, 4, 1, f); fwrite(kps.data(), sizeof(ref_keypoint), n, f); fwrite(desc.data(), 32, n, f);
  fclose(f);
  return 0;
}
