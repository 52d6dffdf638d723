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
//   calibration(vector<Mat>&)    :592           calibration(imgs, seamFinder) fixed-parameter modes 2/3 only (initSeam); the
//                                               seam search is the application's callback (it owns OpenCV), everything
//                                               around it -- low-res resize + warps, dilate / upsample / AND, weights --
//                                               is the library's.  calibration(masks) takes finished m_blenderMask images.
//   process(vector<Mat>&, Mat&)  :1141          process(frames, out)
//   updateMask(vector<Mat>&)     :1218          updateMask(imgs, seamFinder) / updateMask(masks)
//   panocam::{init,getCamFrame,getPanoFrame}    pano::panocam (include/panocam.h:8-28, src/panocamimpl.cpp:149-360)
//   nvrenderAlpha::fit2final                    pano::FitCanvas (src/nvrenderAlpha.cpp:153-189)
#ifndef OCVSTITCHER_B200_HPP
#define OCVSTITCHER_B200_HPP

#include <cmath>
#include <ctime>
#include <fstream>
#include <functional>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
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

// What initSeam / updateMask hand to seam_finder->find (include/ocvstitcher.hpp:1031-1035, 1243-1245): the low-resolution
// warped images (8UC3; the reference converts them to CV_32F first), their corners, and the warped all-255 masks, which
// the finder overwrites with the seam masks.  All tight (step = 3 * width / width).
struct SeamInputs {
    std::vector<std::vector<unsigned char>> images, masks;
    std::vector<int> corners;    // x, y per camera
    std::vector<int> sizes;      // w, h per camera
};
// The application's seam search, e.g. [](pano::SeamInputs &s) { cv::detail::GraphCutSeamFinder(COST_COLOR).find(...); }.
// Leaving the masks untouched is cv::detail::NoSeamFinder.
using SeamFinder = std::function<void(SeamInputs &)>;

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

    // initSeam (:975-1139) from the frames themselves: geometry + tables, then the seam finder's inputs are built exactly as
    // the reference builds them (pano_host_seam_input), `finder` searches the seams, and the low-resolution result goes
    // through the device tail (dilate -> INTER_LINEAR_EXACT -> AND, pano_set_seam_mask).  No Python, no OpenCV in here.
    int calibration(const std::vector<Image> &imgs, const SeamFinder &finder)
    {
        if (calibration(static_cast<const std::vector<Image> *>(nullptr)) != RET_OK) return RET_ERR;
        return updateMask(imgs, finder);
    }

    // updateMask (:1218-1261): re-run the seam search on the current frames
    int updateMask(const std::vector<Image> &imgs, const SeamFinder &finder)
    {
        if (!h_ || (int)imgs.size() != p_.num_images) return RET_ERR;
        SeamInputs s;
        const int n = p_.num_images;
        s.images.resize(n); s.masks.resize(n); s.corners.resize(2 * n); s.sizes.resize(2 * n);
        for (int i = 0; i < n; ++i) {
            int roi[4];
            if (imgs[i].width != p_.width || imgs[i].height != p_.height) { err_ = "frame size differs from the configured one"; return RET_ERR; }
            if (pano_host_seam_input(p_.warp_kind, p_.warped_image_scale, &p_.K[9 * i], &p_.R[9 * i], nullptr, p_.width, p_.height, 0, roi,
                                     nullptr, nullptr) != PANO_OK) { err_ = "seam geometry failed"; return RET_ERR; }
            s.corners[2 * i] = roi[0]; s.corners[2 * i + 1] = roi[1]; s.sizes[2 * i] = roi[2]; s.sizes[2 * i + 1] = roi[3];
            s.images[i].resize((size_t)roi[2] * roi[3] * 3);
            s.masks[i].resize((size_t)roi[2] * roi[3]);
            if (pano_host_seam_input(p_.warp_kind, p_.warped_image_scale, &p_.K[9 * i], &p_.R[9 * i], imgs[i].data, p_.width, p_.height,
                                     imgs[i].step, roi, s.images[i].data(), s.masks[i].data()) != PANO_OK) { err_ = "seam input failed"; return RET_ERR; }
        }
        if (finder) finder(s);
        for (int i = 0; i < n; ++i)
            if (pano_set_seam_mask(h_, i, s.masks[i].data(), s.sizes[2 * i], s.sizes[2 * i + 1], s.sizes[2 * i]) != PANO_OK) {
                err_ = pano_last_error(h_);
                return RET_ERR;
            }
        return RET_OK;
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

// nvrenderAlpha::fit2final (src/nvrenderAlpha.cpp:153-189): the stacked frame scaled to the display canvas
class FitCanvas {
public:
    FitCanvas() = default;
    FitCanvas(const FitCanvas &) = delete;
    FitCanvas &operator=(const FitCanvas &) = delete;
    ~FitCanvas() { pano_fit_destroy(h_); }
    int init(int inW, int inH, int canvasW = 1920, int canvasH = 1080, int device = 0)
    {
        pano_fit_config c{};
        c.in_width = inW; c.in_height = inH; c.canvas_width = canvasW; c.canvas_height = canvasH; c.device = device;
        pano_fit_destroy(h_);
        h_ = nullptr;
        if (pano_fit_create(&c, &h_) != PANO_OK) { err_ = pano_fit_last_error(nullptr); return RET_ERR; }
        return RET_OK;
    }
    int fit2final(const Image &input, Image &canvas)
    {
        if (!h_) return RET_ERR;
        if (pano_fit_compose(h_, input.data, input.step, canvas.data, canvas.step) != PANO_OK) { err_ = pano_fit_last_error(h_); return RET_ERR; }
        return RET_OK;
    }
    const std::string &lastError() const { return err_; }

private:
    pano_fit_handle h_ = nullptr;
    std::string err_;
};

// ---------------------------------------------------------------------------------------------------------
// The SDK facade's pixel path (include/panocam.h:8-28; src/panocamimpl.cpp:149-155 construction, :199-257 init,
// :300-360 getPanoFrame): 2 x num_images cameras, an upper- and a lower-ring stitcher, finalcut crop + vconcat + 4-row
// separator.  Capture (V4L2 / UDP ingest of the slave board), detect, imgEnhancement, render, drawCross, saveAndSend and
// the licence / CAN status byte are outside the compose path: the application hands camera frames in with
// setCamFrame(id, frame) where the reference's nvCam threads would have produced them.
struct PanoCamParams {
    int num_images = 4;                 // cameras per ring (USED_CAMERA_NUM = 2 * num_images)
    CamParams cam;                      // one camera model for the rig, like the reference's camcfgs[] (same sizes / intrinsics)
    StitcherParams stitcher[2];         // [0] upper ring (id 1), [1] lower ring (id 2)
    int finalcut = 0;                   // rows dropped above and below each ring (panocamimpl's finalcut)
};

class panocam {
public:
    panocam() = default;
    panocam(const panocam &) = delete;
    panocam &operator=(const panocam &) = delete;

    // panocam::init, first half (src/panocamimpl.cpp:149-155): the camera pipeline; frames can be handed in afterwards
    int init(const PanoCamParams &p)
    {
        p_ = p;
        if (p.num_images < 1) { err_ = "bad camera count"; return RET_ERR; }
        if (cam_.init(p.cam) != RET_OK) { err_ = "front end: " + cam_.lastError(); return RET_ERR; }
        frames_.assign(2 * p.num_images, Image{nullptr, 0, 0, 0});
        return RET_OK;
    }

    // panocam::init, second half (src/panocamimpl.cpp:199-257: getFrame x 8, stitchers[k]->init(imgs)): both stitchers
    // calibrated with fixed parameters on the current camera frames (setCamFrame every camera first) -- the seam search
    // is `finder` (nullptr = NoSeamFinder) -- then the front end is chained in and the ring composer sized.
    int calibration(const SeamFinder &finder = nullptr)
    {
        const int n = p_.num_images, W = p_.cam.outPutWidth, H = p_.cam.outPutHeight;
        std::vector<std::vector<unsigned char>> buf(2 * n, std::vector<unsigned char>((size_t)W * H * 3));
        std::vector<Image> first(2 * n);
        for (int i = 0; i < 2 * n; ++i) {
            first[i] = Image{buf[i].data(), W, H, W * 3};
            if (getCamFrame(i, first[i]) != RET_OK) { err_ = "setCamFrame() every camera before calibration(): " + err_; return RET_ERR; }
        }
        for (int r = 0; r < 2; ++r) {
            if (st_[r].init(p_.stitcher[r]) != RET_OK ||
                st_[r].calibration(std::vector<Image>(first.begin() + r * n, first.begin() + (r + 1) * n), finder) != RET_OK ||
                st_[r].attachFrontEnd(cam_.handle()) != RET_OK) { err_ = st_[r].lastError(); return RET_ERR; }
            pano_[r].assign((size_t)st_[r].outWidth() * st_[r].outHeight() * 3, 0);
        }
        if (rc_.init(st_[0].outWidth(), st_[0].outHeight(), st_[1].outWidth(), st_[1].outHeight(), PANO_RING_CROP, p_.finalcut) != RET_OK) {
            err_ = "ring composer: " + rc_.lastError();
            return RET_ERR;
        }
        return RET_OK;
    }

    // stands in for the capture threads: the latest camera frame of camera `id` (8UC4, or 8UC2 with cam.yuyv); the
    // buffer must stay valid until getPanoFrame returns
    int setCamFrame(int id, const Image &frame)
    {
        if (id < 0 || id >= (int)frames_.size()) return RET_ERR;
        frames_[id] = frame;
        return RET_OK;
    }
    // panocam::getCamFrame (src/panocam.cpp): camera `id`'s stitcher input (undistorted, cropped, resized), outPutWidth x outPutHeight x 3
    int getCamFrame(int id, Image &frame)
    {
        if (id < 0 || id >= (int)frames_.size() || !frames_[id].data) return RET_ERR;
        if (cam_.getFrame(frames_[id], frame) != RET_OK) { err_ = cam_.lastError(); return RET_ERR; }
        return RET_OK;
    }
    int outWidth() const { return rc_.outWidth(); }
    int outHeight() const { return rc_.outHeight(); }
    // panocam::getPanoFrame (src/panocamimpl.cpp:300-360): both rings composed on two threads like the reference's t1 / t2,
    // then cropped by finalcut, stacked and separated by the 4-row bar.  ret: outWidth() x outHeight() x 3
    int getPanoFrame(Image &ret)
    {
        const int n = p_.num_images;
        for (const Image &f : frames_)
            if (!f.data) { err_ = "setCamFrame() every camera first"; return RET_ERR; }
        int rcode[2] = {RET_ERR, RET_ERR};
        Image out[2];
        std::thread t[2];
        for (int r = 0; r < 2; ++r) {
            out[r] = Image{pano_[r].data(), st_[r].outWidth(), st_[r].outHeight(), st_[r].outWidth() * 3};
            t[r] = std::thread([this, r, n, &rcode, &out]() {
                rcode[r] = st_[r].process(std::vector<Image>(frames_.begin() + r * n, frames_.begin() + (r + 1) * n), out[r]);
            });
        }
        t[0].join();
        t[1].join();
        for (int r = 0; r < 2; ++r)
            if (rcode[r] != RET_OK) { err_ = st_[r].lastError(); return RET_ERR; }
        if (rc_.compose(out[0], out[1], ret) != RET_OK) { err_ = rc_.lastError(); return RET_ERR; }
        return RET_OK;
    }
    ocvStitcher &stitcher(int ring) { return st_[ring]; }
    const std::string &lastError() const { return err_; }

private:
    PanoCamParams p_;
    nvCamFrontEnd cam_;
    ocvStitcher st_[2];
    RingComposer rc_;
    std::vector<Image> frames_;
    std::vector<unsigned char> pano_[2];
    std::string err_;
};

}  // namespace pano
#endif
