// ocvstitcher_b200.hpp -- source-level stand-in for the reference's `ocvStitcher`
// (/root/reference/include/ocvstitcher.hpp:254-1306) over the C ABI in panob200.h.
//
// Same method names, argument meaning and RET_OK / RET_ERR behaviour for the per-frame path;
// images cross as raw interleaved BGR pointers (pano::Image) instead of cv::Mat so the header
// needs no OpenCV.  A caller that has cv::Mat passes {mat.data, mat.cols, mat.rows, mat.step}.
//
//   reference                                   here
//   ---------                                   ----
//   init(std::string& cfgPath)   :262           init(const StitcherParams&)   (yaml parsing stays in the app)
//   calibration(vector<Mat>&)    :592           calibration(masks)            fixed-parameter modes 2/3 only
//   process(vector<Mat>&, Mat&)  :1141          process(frames, out)
//   updateMask(vector<Mat>&)     :1218          updateMask(masks)             masks come from the host seam search
#ifndef OCVSTITCHER_B200_HPP
#define OCVSTITCHER_B200_HPP

#include <cmath>
#include <string>
#include <vector>

#include "panob200.h"

namespace pano {

constexpr int RET_OK = 0, RET_ERR = -1;   // include/stitcherglobal.h:13-14

struct Image {           // view of an 8-bit interleaved image
    unsigned char *data;
    int width, height, step;   // step in bytes
};

struct StitcherParams {  // stStitcherCfg (+ the calibration the reference keeps on the object)
    int width = 0, height = 0, num_images = 0;
    float blendStrength = 1.f;                 // stitcherBlenderStrength
    std::vector<float> K, R;                   // num_images x 9 each (camK, cameraR)
    float warped_image_scale = 0.f;
    int cut[4] = {0, 0, 0, 0};                 // m_cutParams; w <= 0 -> whole panorama
    int warp_kind = PANO_WARP_SPHERICAL;
    int blender = -1;                          // -1: reference rule (:1188-1195), else PANO_BLEND_*
    int num_bands = -1;                        // -1: reference rule
    int device = 0, max_batch = 1;
};

class ocvStitcher {
public:
    ocvStitcher() = default;
    ocvStitcher(const ocvStitcher &) = delete;
    ocvStitcher &operator=(const ocvStitcher &) = delete;
    ~ocvStitcher() { pano_destroy(h_); }

    int init(const StitcherParams &p)
    {
        p_ = p;
        if (p.num_images < 1 || (int)p.K.size() != 9 * p.num_images || (int)p.R.size() != 9 * p.num_images) {
            err_ = "bad camera parameters";
            return RET_ERR;
        }
        return RET_OK;
    }

    // initSeam with fixed parameters: geometry + tables on the device.  `masks` (optional) are the
    // m_blenderMask images produced by the host seam search; without them the warped all-255
    // masks are used (NoSeamFinder behaviour, :1034).
    int calibration(const std::vector<Image> *masks = nullptr)
    {
        // dst size decides the blender exactly like :1188-1195
        std::vector<int> corners(2 * p_.num_images), sizes(2 * p_.num_images);
        for (int i = 0; i < p_.num_images; ++i) {
            int roi[4];
            if (pano_host_warp_roi(p_.warp_kind, p_.warped_image_scale, &p_.K[9 * i], &p_.R[9 * i], p_.width, p_.height, roi))
                return RET_ERR;
            corners[2 * i] = roi[0]; corners[2 * i + 1] = roi[1]; sizes[2 * i] = roi[2]; sizes[2 * i + 1] = roi[3];
        }
        int dst[4];
        pano_host_blend_geometry(p_.num_images, corners.data(), sizes.data(), 0, dst, nullptr, nullptr, nullptr, nullptr);
        const float blend_width = std::sqrt(static_cast<float>(dst[2] * dst[3])) * p_.blendStrength / 100.f;
        pano_config c{};
        c.num_images = p_.num_images; c.src_width = p_.width; c.src_height = p_.height;
        c.warp_kind = p_.warp_kind; c.warped_image_scale = p_.warped_image_scale;
        c.K = p_.K.data(); c.R = p_.R.data();
        c.blender = p_.blender >= 0 ? p_.blender : (blend_width < 1.f ? PANO_BLEND_NO : PANO_BLEND_MULTIBAND);
        c.num_bands = p_.num_bands >= 0 ? p_.num_bands
                                        : (blend_width < 1.f ? 0 : static_cast<int>(std::ceil(std::log(blend_width) / std::log(2.)) - 1.));
        c.sharpness = blend_width > 0.f ? 1.f / blend_width : 0.02f;
        for (int k = 0; k < 4; ++k) c.cut[k] = p_.cut[k];
        c.device = p_.device; c.max_batch = p_.max_batch;
        pano_destroy(h_);
        h_ = nullptr;
        if (pano_create(&c, &h_) != PANO_OK) { err_ = pano_last_error(nullptr); return RET_ERR; }
        return masks ? updateMask(*masks) : RET_OK;
    }

    int updateMask(const std::vector<Image> &masks)
    {
        if (!h_ || (int)masks.size() != p_.num_images) return RET_ERR;
        for (int i = 0; i < p_.num_images; ++i)
            if (pano_set_mask(h_, i, masks[i].data, masks[i].width, masks[i].height, masks[i].step) != PANO_OK) {
                err_ = pano_last_error(h_);
                return RET_ERR;
            }
        return RET_OK;
    }

    // process(imgs, ret): `out` must be outWidth() x outHeight() x 3
    int process(const std::vector<Image> &imgs, Image &out)
    {
        if (!h_ || (int)imgs.size() != p_.num_images) return RET_ERR;
        std::vector<const unsigned char *> ptr(imgs.size());
        std::vector<int> step(imgs.size());
        for (size_t i = 0; i < imgs.size(); ++i) { ptr[i] = imgs[i].data; step[i] = imgs[i].step; }
        if (pano_process(h_, ptr.data(), step.data(), out.data, out.step) != PANO_OK) { err_ = pano_last_error(h_); return RET_ERR; }
        return RET_OK;
    }

    int outWidth() const { int wh[2] = {0, 0}; pano_get_geometry(h_, nullptr, nullptr, nullptr, wh); return wh[0]; }
    int outHeight() const { int wh[2] = {0, 0}; pano_get_geometry(h_, nullptr, nullptr, nullptr, wh); return wh[1]; }
    pano_handle handle() const { return h_; }
    const std::string &lastError() const { return err_; }

private:
    StitcherParams p_;
    pano_handle h_ = nullptr;
    std::string err_;
};

}  // namespace pano
#endif
