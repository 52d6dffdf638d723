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
#include <ctime>
#include <fstream>
#include <sstream>
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

    // chain a front end: process() then takes the camera frames themselves (getFrame x N + process, src/master.cpp:300-318)
    int attachFrontEnd(pano_frontend_handle f)
    {
        if (!h_ || pano_attach_frontend(h_, -1, f) != PANO_OK) { err_ = h_ ? pano_last_error(h_) : "calibration() first"; return RET_ERR; }
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

// ---------------------------------------------------------------------------------------------------------
// Calibration-file persistence (include/ocvstitcher.hpp:452-562): cameraparaout_<id>.txt holds blocks of
//   <%F-%H-%M-%S>:            one line containing ':'
//   k00,...,k22,r00,...,r22,  one line of 18 floats per camera
//   <warped_image_scale>
// initCamParams reads the LAST block; saveCameraParams appends one.
inline int loadCameraParams(const std::string &cfgPath, int id, int num_images, StitcherParams &p, std::string *err = nullptr)
{
    const std::string filename = cfgPath + "cameraparaout_" + std::to_string(id) + ".txt";
    std::ifstream fin(filename);
    auto fail = [&](const char *m) { if (err) *err = std::string(m) + " (" + filename + ")"; return RET_ERR; };
    if (!fin.is_open()) return fail("can not open camerapara file, no preset parameters found");
    std::vector<std::string> lines;
    for (std::string l; std::getline(fin, l);) {
        while (!l.empty() && (l.back() == '\r' || l.back() == ' ')) l.pop_back();
        if (!l.empty()) lines.push_back(l);
    }
    int last = -1;
    for (size_t i = 0; i < lines.size(); ++i)
        if (lines[i].find(':') != std::string::npos) last = (int)i;
    if (last < 0) return fail("no preset parameters found");
    if ((int)lines.size() < last + 2 + num_images) return fail("camera preset parameter block incomplete");
    p.K.assign(9 * num_images, 0.f); p.R.assign(9 * num_images, 0.f);
    for (int i = 0; i < num_images; ++i) {
        std::vector<float> v;
        std::stringstream ss(lines[last + 1 + i]);
        for (std::string tok; std::getline(ss, tok, ',');)
            if (!tok.empty()) v.push_back(std::stof(tok));
        if (v.size() != 18) return fail("camera preset parameter incorrect");                     // :497-501
        for (int k = 0; k < 9; ++k) { p.K[9 * i + k] = v[k]; p.R[9 * i + k] = v[9 + k]; }
    }
    p.warped_image_scale = std::stof(lines[last + 1 + num_images]);
    p.num_images = num_images;
    return RET_OK;
}

inline int saveCameraParams(const std::string &cfgPath, int id, const StitcherParams &p)
{
    const std::string filename = cfgPath + "cameraparaout_" + std::to_string(id) + ".txt";
    std::ofstream fout(filename, std::ofstream::out | std::ofstream::app);
    if (!fout.is_open()) return RET_ERR;
    const std::time_t tt = std::time(nullptr);
    char stamp[64];
    std::strftime(stamp, sizeof stamp, "%F-%H-%M-%S", std::localtime(&tt));
    fout << stamp << ":\n";
    for (int i = 0; i < p.num_images; ++i) {
        for (int k = 0; k < 9; ++k) fout << p.K[9 * i + k] << ",";
        for (int k = 0; k < 9; ++k) fout << p.R[9 * i + k] << ",";
        fout << "\n";
    }
    fout << p.warped_image_scale << "\n";
    return RET_OK;
}

// ---------------------------------------------------------------------------------------------------------
// nvCam's pixel pipeline (include/nvcam.hpp:898-929 + getFrame(mat, false) :1092-1094) behind the reference's
// getFrame name; capture itself (V4L2 / NvBuffer) stays with the application, which hands the camera frame in.
struct CamParams {               // stCamCfg (include/stitcherglobal.h:39-55) + the matched cameras.yaml entry
    int camSrcWidth = 1920, camSrcHeight = 1080, undistoredWidth = 1920, undistoredHeight = 1080;
    int outPutWidth = 1920, outPutHeight = 1080;
    bool undistor = true;
    double K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, distorParams[4] = {0, 0, 0, 0};
    double newK[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // getOptimalNewCameraMatrix(K, D, size, 1, size, 0)  (:831, host, OpenCV)
    int rect[4] = {0, 0, 0, 0};
    bool yuyv = false;           // 8UC2 YUYV frames (the YUYVCAM build, :880-886) instead of the VIC's 8UC4
    int device = 0, max_batch = 1;
};

class nvCamFrontEnd {
public:
    nvCamFrontEnd() = default;
    nvCamFrontEnd(const nvCamFrontEnd &) = delete;
    nvCamFrontEnd &operator=(const nvCamFrontEnd &) = delete;
    ~nvCamFrontEnd() { pano_frontend_destroy(h_); }
    int init(const CamParams &c)
    {
        pano_frontend_config f{};
        f.cam_src_width = c.camSrcWidth; f.cam_src_height = c.camSrcHeight;
        f.undist_width = c.undistoredWidth; f.undist_height = c.undistoredHeight;
        f.out_width = c.outPutWidth; f.out_height = c.outPutHeight;
        f.undistort = c.undistor ? 1 : 0;
        for (int k = 0; k < 9; ++k) { f.K[k] = c.K[k]; f.newK[k] = c.newK[k]; }
        for (int k = 0; k < 4; ++k) { f.D[k] = c.distorParams[k]; f.rect[k] = c.rect[k]; }
        f.device = c.device; f.max_batch = c.max_batch;
        f.src_format = c.yuyv ? PANO_SRC_YUYV : PANO_SRC_BGRA;
        pano_frontend_destroy(h_);
        h_ = nullptr;
        if (pano_frontend_create(&f, &h_) != PANO_OK) { err_ = pano_frontend_last_error(nullptr); return RET_ERR; }
        return RET_OK;
    }
    // getFrame(frame, src=false): camera frame in (8UC4, or 8UC2 when yuyv), stitcher input (8UC3) out
    int getFrame(const Image &camFrame, Image &out)
    {
        if (!h_) return RET_ERR;
        if (pano_frontend_process(h_, camFrame.data, camFrame.step, out.data, out.step) != PANO_OK) {
            err_ = pano_frontend_last_error(h_);
            return RET_ERR;
        }
        return RET_OK;
    }
    pano_frontend_handle handle() const { return h_; }
    const std::string &lastError() const { return err_; }

private:
    pano_frontend_handle h_ = nullptr;
    std::string err_;
};

// The two-ring caller step after the two process calls (src/master.cpp:321-326; crop variant src/panocamimpl.cpp:354-360)
class RingComposer {
public:
    RingComposer() = default;
    RingComposer(const RingComposer &) = delete;
    RingComposer &operator=(const RingComposer &) = delete;
    ~RingComposer() { pano_ring_destroy(h_); }
    int init(int upW, int upH, int downW, int downH, int mode = PANO_RING_RESIZE, int finalcut = 0, int bar = -1, int device = 0)
    {
        pano_ring_config c{};
        c.up_width = upW; c.up_height = upH; c.down_width = downW; c.down_height = downH;
        c.mode = mode; c.finalcut = finalcut; c.bar = bar >= 0 ? bar : (mode == PANO_RING_RESIZE ? 10 : 4); c.device = device;
        pano_ring_destroy(h_);
        h_ = nullptr;
        if (pano_ring_create(&c, &h_) != PANO_OK) { err_ = pano_ring_last_error(nullptr); return RET_ERR; }
        return RET_OK;
    }
    int outWidth() const { int wh[2] = {0, 0}; pano_ring_out_size(h_, wh); return wh[0]; }
    int outHeight() const { int wh[2] = {0, 0}; pano_ring_out_size(h_, wh); return wh[1]; }
    int compose(const Image &up, const Image &down, Image &ret)
    {
        if (!h_) return RET_ERR;
        if (pano_ring_compose(h_, up.data, up.step, down.data, down.step, ret.data, ret.step) != PANO_OK) {
            err_ = pano_ring_last_error(h_);
            return RET_ERR;
        }
        return RET_OK;
    }
    const std::string &lastError() const { return err_; }

private:
    pano_ring_handle h_ = nullptr;
    std::string err_;
};

}  // namespace pano
#endif
