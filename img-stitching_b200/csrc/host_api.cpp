// Host-only entry points of the C ABI: the init-time table builders, callable without a CUDA
// device (used by callers that want geometry before creating a handle, and by the CPU tests).
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/panob200.h"
#include "geometry.hpp"

using namespace pano;

extern "C" {

int pano_host_warp_roi(int warp_kind, float scale, const float *K, const float *R, int src_w, int src_h, int *roi)
{
    if (!K || !R || !roi || src_w < 1 || src_h < 1) return PANO_ERR;
    RotationWarper w(warp_kind, scale);
    w.setCamera(K, R);
    const Rect r = w.warpRoi(src_w, src_h);
    roi[0] = r.x; roi[1] = r.y; roi[2] = r.w; roi[3] = r.h;
    return PANO_OK;
}

int pano_host_build_maps(int warp_kind, float scale, const float *K, const float *R, int src_w, int src_h,
                         float *xmap, float *ymap)
{
    if (!K || !R || !xmap || !ymap) return PANO_ERR;
    RotationWarper w(warp_kind, scale);
    w.setCamera(K, R);
    const Rect r = w.warpRoi(src_w, src_h);
    w.buildMaps(src_w, src_h, r, xmap, ymap);
    return PANO_OK;
}

int pano_host_blend_geometry(int n, const int *corners, const int *sizes, int num_bands, int *dst_roi,
                             int *eff_bands, int *padded_wh, int *feed_rects, int *borders)
{
    if (n < 1 || !corners || !sizes) return PANO_ERR;
    std::vector<Rect> rois(n);
    for (int i = 0; i < n; ++i) { rois[i].x = corners[2 * i]; rois[i].y = corners[2 * i + 1]; rois[i].w = sizes[2 * i]; rois[i].h = sizes[2 * i + 1]; }
    const Rect roi = resultRoi(rois);
    int pw, ph;
    const int nb = multibandPrepare(roi, num_bands, pw, ph);
    if (dst_roi) { dst_roi[0] = roi.x; dst_roi[1] = roi.y; dst_roi[2] = roi.w; dst_roi[3] = roi.h; }
    if (eff_bands) *eff_bands = nb;
    if (padded_wh) { padded_wh[0] = pw; padded_wh[1] = ph; }
    for (int i = 0; i < n; ++i) {
        const FeedRect f = multibandFeedRect(roi, pw, ph, nb, rois[i]);
        if (feed_rects) { feed_rects[4 * i] = f.rect.x; feed_rects[4 * i + 1] = f.rect.y; feed_rects[4 * i + 2] = f.rect.w; feed_rects[4 * i + 3] = f.rect.h; }
        if (borders) { borders[4 * i] = f.top; borders[4 * i + 1] = f.bottom; borders[4 * i + 2] = f.left; borders[4 * i + 3] = f.right; }
    }
    return PANO_OK;
}

/* folded sample position for integer coordinate i, fraction f (0..31), axis length n */
unsigned pano_host_fold_reflect(int i, int f, int n) { return foldReflect(i, f, n); }

int pano_host_fixed_maps(const float *xmap, const float *ymap, size_t count, int16_t *ixy, uint16_t *frac)
{
    if (!xmap || !ymap) return PANO_ERR;
    for (size_t p = 0; p < count; ++p) {
        const FixedCoord fc = toFixed(xmap[p], ymap[p]);
        if (ixy) { ixy[2 * p] = (int16_t)fc.ix; ixy[2 * p + 1] = (int16_t)fc.iy; }
        if (frac) frac[p] = (uint16_t)(fc.fy * 32 + fc.fx);
    }
    return PANO_OK;
}

int pano_host_pyrdown_f32(const float *src, int w, int h, float *dst)
{
    if (!src || !dst || w < 1 || h < 1) return PANO_ERR;
    pyrDownF32(src, w, h, dst);
    return PANO_OK;
}

int pano_host_feather_weight(const uint8_t *mask, int w, int h, int stride, float sharpness, float *out)
{
    if (!mask || !out || w < 1 || h < 1 || stride < w) return PANO_ERR;
    featherWeight(mask, w, h, stride, sharpness, out);
    return PANO_OK;
}

int pano_host_undistort_maps(const double *K, const double *D, const double *newK, int w, int h, float *mapx, float *mapy)
{
    if (!K || !D || !newK || !mapx || !mapy || w < 1 || h < 1) return PANO_ERR;
    undistortMaps(K, D, newK, w, h, mapx, mapy);
    return PANO_OK;
}

int pano_host_cubic_table(int16_t *tab)
{
    if (!tab) return PANO_ERR;
    cubicTable(tab);
    return PANO_OK;
}

int pano_host_resize_axis(int ssize, int dsize, int clamp_frac, int *ofs, int16_t *a0, int16_t *a1)
{
    if (ssize < 1 || dsize < 1 || !ofs || !a0 || !a1) return PANO_ERR;
    std::vector<int> o;
    std::vector<int16_t> x0, x1;
    resizeAxis(ssize, dsize, clamp_frac != 0, o, x0, x1);
    std::memcpy(ofs, o.data(), sizeof(int) * dsize);
    std::memcpy(a0, x0.data(), sizeof(int16_t) * dsize);
    std::memcpy(a1, x1.data(), sizeof(int16_t) * dsize);
    return PANO_OK;
}

int pano_host_linear_exact_axis(int ssize, int dsize, int *ofs, int *c1)
{
    if (ssize < 1 || dsize < 1 || !ofs || !c1) return PANO_ERR;
    std::vector<int> o, c;
    linearExactAxis(ssize, dsize, o, c);
    std::memcpy(ofs, o.data(), sizeof(int) * dsize);
    std::memcpy(c1, c.data(), sizeof(int) * dsize);
    return PANO_OK;
}

double pano_host_seam_scale(int src_w, int src_h) { return (src_w > 0 && src_h > 0) ? seamWorkAspect(src_w, src_h) : 0.0; }

int pano_host_seam_input(int warp_kind, float warped_image_scale, const float *K, const float *R, const uint8_t *frame, int src_w,
                         int src_h, int stride, int *roi, uint8_t *image_warped, uint8_t *mask_warped)
{
    if (!K || !R || !roi || src_w < 2 || src_h < 2) return PANO_ERR;
    // include/ocvstitcher.hpp:988-1017: resize by seam_work_aspect, K scaled in float, warper at float(scale * aspect)
    const double aspect = seamWorkAspect(src_w, src_h);
    const int sw = (int)std::nearbyint(src_w * aspect), sh = (int)std::nearbyint(src_h * aspect);   // saturate_cast<int>(double) == cvRound
    if (sw < 1 || sh < 1) return PANO_ERR;
    float Ks[9];
    std::memcpy(Ks, K, sizeof Ks);
    const float swa = (float)aspect;
    Ks[0] *= swa; Ks[2] *= swa; Ks[4] *= swa; Ks[5] *= swa;
    RotationWarper w(warp_kind, static_cast<float>(warped_image_scale * aspect));
    w.setCamera(Ks, R);
    const Rect r = w.warpRoi(sw, sh);
    roi[0] = r.x; roi[1] = r.y; roi[2] = r.w; roi[3] = r.h;
    if (!image_warped && !mask_warped) return PANO_OK;
    if (r.w <= 0 || r.h <= 0) return PANO_ERR;
    std::vector<float> xm((size_t)r.w * r.h), ym((size_t)r.w * r.h);
    w.buildMaps(sw, sh, r, xm.data(), ym.data());
    if (mask_warped) warpedFullMask(xm.data(), ym.data(), r.w, r.h, sw, sh, mask_warped, r.w);
    if (image_warped) {
        if (!frame || stride < 3 * src_w) return PANO_ERR;
        std::vector<uint8_t> small((size_t)sw * sh * 3);
        resizeLinearExactU8(frame, src_w, src_h, stride, 3, aspect, aspect, sw, sh, small.data(), 3 * sw);
        remapBilinearReflectU8(small.data(), sw, sh, 3 * sw, 3, xm.data(), ym.data(), r.w, r.h, image_warped, 3 * r.w);
    }
    return PANO_OK;
}

}  // extern "C"
