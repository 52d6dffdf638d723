// Host-side (init-time) table builders for the compose path.
//
// Everything here runs ONCE per calibration / mask refresh on the host, exactly where the
// reference does it (ocvStitcher::initSeam, include/ocvstitcher.hpp:975-1139), and produces
// the static tables the CUDA kernels consume: rotation-warp remap tables, blender geometry,
// weight pyramids, feather weights, undistort maps.  The arithmetic follows OpenCV's CPU path
// bit-for-bit where that is achievable; this file must be compiled
// with -ffp-contract=off.
#pragma once
#include <cstdint>
#include <vector>

namespace pano {

struct Rect { int x = 0, y = 0, w = 0, h = 0; };

// ---- rotation warpers (cv::detail::RotationWarperBase<Spherical|CylindricalProjector>) ----
class RotationWarper {
public:
    RotationWarper(int kind, float scale) : kind_(kind), scale_(scale) {}
    void setCamera(const float K[9], const float R[9]);
    // warpRoi(): tl + size of the warped image (ocvstitcher.hpp:1057)
    Rect warpRoi(int src_w, int src_h) const;
    // buildMaps(): float32 backward maps over warpRoi()
    void buildMaps(int src_w, int src_h, const Rect &roi, float *xmap, float *ymap) const;
private:
    void forward(float x, float y, float &u, float &v) const;
    void backward(float u, float v, float &x, float &y) const;
    int kind_;
    float scale_;
    float k_[9], rinv_[9], r_kinv_[9], k_rinv_[9];
};

// cv::remap's fixed-point view of a float map entry: sx = cvRound(32*x) etc.
// ix/iy are saturated to int16 exactly like cv::convertMaps does.
struct FixedCoord { int ix, iy, fx, fy; };
FixedCoord toFixed(float mx, float my);

// Fold BORDER_REFLECT into an equivalent in-range sample position (see DESIGN.md "map folding").
// Returns 32*i + f with i in [0, n-1], f in [0,31]; taps are (i, min(i+1, n-1)).
uint32_t foldReflect(int i, int f, int n);

// ---- blender geometry -----------------------------------------------------------------
Rect resultRoi(const std::vector<Rect> &rois);
struct FeedRect {
    Rect rect;                      // in padded-dst coordinates (level 0)
    int top, bottom, left, right;   // copyMakeBorder amounts
};
// MultiBandBlender::prepare: effective band count and padded size
int multibandPrepare(const Rect &dst_roi, int num_bands, int &padded_w, int &padded_h);
// MultiBandBlender::feed rect arithmetic
FeedRect multibandFeedRect(const Rect &dst_roi, int padded_w, int padded_h, int nb, const Rect &img_roi);

// ---- weights ----------------------------------------------------------------------------
// cv::pyrDown on CV_32F, bit-exact with OpenCV 4.x's 4-lane universal-intrinsics build: the vector bodies and the
// scalar border / tail loops of pyramids.cpp sum in different orders, pyrDownColumnRule says which columns get which
// (horizontal pass: vector order for 1 <= x < hv_end; vertical pass: vector order for x < vv_end)
void pyrDownColumnRule(int sw, int *hv_end, int *vv_end);
void pyrDownF32(const float *src, int sw, int sh, float *dst);
// cv::distanceTransform(mask, DIST_L1, 3) followed by min(d*sharpness, 1)
void featherWeight(const uint8_t *mask, int w, int h, int stride, float sharpness, float *out);

// ---- nvCam undistort maps (cv::initUndistortRectifyMap, R = I, CV_32FC1) ---------------------
void undistortMaps(const double K[9], const double D[4], const double newK[9], int w, int h,
                   float *mapx, float *mapy);
// 15-bit bicubic weight table of cv::remap (1024 x 16 shorts)
void cubicTable(int16_t *tab);
// cv::resize INTER_LINEAR u8 coefficient tables for one axis
void resizeAxis(int ssize, int dsize, bool clamp_frac, std::vector<int> &ofs,
                std::vector<int16_t> &a0, std::vector<int16_t> &a1, double inv_scale = 0.0);

// cv::resize INTER_LINEAR_EXACT coefficient table for one axis (8UC1 path: ufixedpoint16, 8 fractional bits):
// sample d reads src[ofs[d]] * (256 - c1[d]) + src[min(ofs[d] + 1, ssize - 1)] * c1[d]
void linearExactAxis(int ssize, int dsize, std::vector<int> &ofs, std::vector<int> &c1, double inv_scale = 0.0);

// ---- the seam finder's inputs (initSeam :985-1017, updateMask :1228-1242; init-time, host) ----------------
double seamWorkAspect(int src_w, int src_h);
// cv::resize(src, dst, Size(), fx, fy, INTER_LINEAR_EXACT), 8-bit, `ch` interleaved channels; dw x dh = cvRound(w * fx) x cvRound(h * fy)
void resizeLinearExactU8(const uint8_t *src, int w, int h, int stride, int ch, double fx, double fy, int dw, int dh,
                         uint8_t *dst, int dstride);
// cv::remap(INTER_LINEAR, BORDER_REFLECT) through float maps
void remapBilinearReflectU8(const uint8_t *src, int w, int h, int stride, int ch, const float *xmap, const float *ymap,
                            int mw, int mh, uint8_t *dst, int dstride);
// warp of an all-255 mask, INTER_NEAREST / BORDER_CONSTANT
void warpedFullMask(const float *xmap, const float *ymap, int mw, int mh, int src_w, int src_h, uint8_t *dst, int dstride);

}  // namespace pano
