/*
 * panob200.h -- C ABI of the B200-native per-frame compose path of Img-Stitching.
 *
 * The reference has no FFI: its hot path is a header-only C++ class compiled into every
 * executable (`ocvStitcher`, /root/reference/include/ocvstitcher.hpp:254-1306, and the
 * `nvCam` pixel pipeline, include/nvcam.hpp:898-929,1083-1100).  This header is the
 * boundary a maintainer binds instead; each entry point cites the reference interface it
 * replaces.  Plain pointers and sizes only; no exceptions cross it; every call returns
 * PANO_OK (0) or PANO_ERR (-1), mirroring RET_OK / RET_ERR
 * (include/stitcherglobal.h:13-14).  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with PANO_ERR.
 *
 * Image conventions (same as the reference's cv::Mat usage): 8-bit, interleaved, row-major,
 * BGR order as delivered by nvCam::getFrame; strides in bytes.
 */
#ifndef PANOB200_H
#define PANOB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PANO_OK 0
#define PANO_ERR (-1)

/* warper kinds: cv::SphericalWarper (include/ocvstitcher.hpp:1000) and
 * cv::CylindricalWarper (src/stitching_detailed.cpp:652-653; ocvstitcher.hpp:999 commented) */
#define PANO_WARP_SPHERICAL 0
#define PANO_WARP_CYLINDRICAL 1

/* blender kinds: Blender::NO (include/ocvstitcher.hpp:1190-1191), FeatherBlender
 * (src/stitching_detailed.cpp:865-869), MultiBandBlender (include/ocvstitcher.hpp:1186-1199) */
#define PANO_BLEND_NO 0
#define PANO_BLEND_FEATHER 1
#define PANO_BLEND_MULTIBAND 2

typedef struct pano_ctx *pano_handle;
typedef struct pano_frontend_ctx *pano_frontend_handle;

/* Static configuration = the members ocvStitcher keeps after init
 * (include/ocvstitcher.hpp:1266-1299: camK, cameraR, warped_image_scale, m_cutParams) plus
 * the stStitcherCfg fields that shape the compose (include/stitcherglobal.h:68-81). */
typedef struct pano_config {
    int num_images;            /* stStitcherCfg::num_images */
    int src_width, src_height; /* stStitcherCfg::width/height (outPutWidth/outPutHeight) */
    int warp_kind;             /* PANO_WARP_* */
    float warped_image_scale;  /* ocvStitcher::warped_image_scale */
    const float *K;            /* num_images x 9, row-major float32 (camK[i]) */
    const float *R;            /* num_images x 9, row-major float32 (cameraR[i]) */
    int blender;               /* PANO_BLEND_* */
    int num_bands;             /* MultiBandBlender::setNumBands argument (:1195) */
    float sharpness;           /* FeatherBlender::setSharpness argument */
    int cut[4];                /* m_cutParams x,y,w,h in dst-roi coordinates; w<=0 -> whole dst roi */
    int device;                /* CUDA device ordinal */
    int max_batch;             /* frame-sets resident per launch wave (workspace sizing), >=1 */
} pano_config;

/* Error text of the last failing call on this thread (create) or handle. */
const char *pano_last_error(pano_handle h);

/* ocvStitcher::init + the geometry half of initSeam (:1054-1063, :1110): builds the
 * rotation-warp tables on the host (bit-exact with RotationWarperBase::buildMaps), the dst
 * roi, and the blender geometry; allocates all device memory.  Masks default to the warped
 * all-255 mask (what NoSeamFinder would leave, :1034). */
int pano_create(const pano_config *cfg, pano_handle *out);
int pano_destroy(pano_handle h);

/* m_corners / m_sizes / resultRoi (include/ocvstitcher.hpp:1055-1063,1110).
 * corners, sizes: 2*num_images ints; dst_roi: x,y,w,h; out_wh: size process() writes. */
int pano_get_geometry(pano_handle h, int *corners, int *sizes, int *dst_roi, int *out_wh);
/* effective band count and the padded dst size after MultiBandBlender::prepare */
int pano_get_blend_geometry(pano_handle h, int *num_bands, int *padded_wh, int *feed_rects /*4 per image*/);

/* The remap tables RotationWarperBase::buildMaps would produce for camera `cam` (float32,
 * sizes[cam] large), and their fixed-point form as cv::remap consumes them
 * (ixy: int16 pairs, frac: (fy<<5)|fx) -- the parity tests check these bit-exactly. */
int pano_get_warp_maps(pano_handle h, int cam, float *xmap, float *ymap);
int pano_get_fixed_maps(pano_handle h, int cam, int16_t *ixy, uint16_t *frac);

/* m_blenderMask[cam] (include/ocvstitcher.hpp:1101,1257).  Re-callable at run time: this is
 * how updateMask (:1218-1261) lands.  mask is sizes[cam] large, 8-bit, soft (0..255). */
int pano_set_mask(pano_handle h, int cam, const uint8_t *mask, int width, int height, int stride);
/* The tail of initSeam / updateMask on the device (include/ocvstitcher.hpp:1095-1101, 1251-1257), starting from the
 * seam finder's LOW-RESOLUTION mask of camera `cam` (masks_warped[i] after seam_finder->find, any size):
 *   cv::dilate(3x3) -> cv::resize(INTER_LINEAR_EXACT) to sizes[cam] -> AND with the warped full mask
 *   (m_compensatorMaskWarped[cam]) = m_blenderMask[cam], then everything pano_set_mask does.
 * Bit-exact with the OpenCV calls; uploads a few KB instead of the full-resolution mask. */
int pano_set_seam_mask(pano_handle h, int cam, const uint8_t *seam, int width, int height, int stride);
/* m_blenderMask[cam] as the handle currently holds it (sizes[cam] large) */
int pano_get_mask(pano_handle h, int cam, uint8_t *mask, int stride);
/* Optional: override the float weight pyramid level the library derives from the mask
 * (MultiBandBlender::feed builds it with cv::pyrDown on CV_32F; the library's own builder follows
 * OpenCV 4.x's per-column summation order and is bit-exact with it, so this is only needed for
 * tables that come from elsewhere).  level in [0, num_bands]; size = feed rect >> level. */
int pano_set_weight_level(pano_handle h, int cam, int level, const float *w, int width, int height);
/* The weight level the compose currently uses for camera `cam` (feed rect >> level large, float32): the library's
 * own pyramid (built on the device at every pano_set_mask) or the caller's override.  Multiband: level in
 * [0, num_bands]; feather: level 0 = the feather weight map (sizes[cam] large). */
int pano_get_weight_level(pano_handle h, int cam, int level, float *w);
/* Optional: FeatherBlender weight map (distanceTransform-derived, static); sizes[cam] large.
 * Without it the library derives it from the mask. */
int pano_set_feather_weight(pano_handle h, int cam, const float *w, int width, int height);
/* compensator->apply (src/stitching_detailed.cpp:841; ocvstitcher.hpp:1178 commented):
 * full-resolution float gain map (BlocksGainCompensator) or scalar gain (GainCompensator).
 * gain == NULL switches gain apply off for the camera. */
int pano_set_gain_map(pano_handle h, int cam, const float *gain, int width, int height);
int pano_set_gain_scalar(pano_handle h, int cam, double gain);

/* ocvStitcher::process (include/ocvstitcher.hpp:1141-1216) for one frame-set held in HOST
 * memory: frames[i] -> src_height rows of strides[i] bytes; out receives the cut rectangle,
 * out_stride bytes per row (>= 3*cut_w).  Synchronous. */
int pano_process(pano_handle h, const uint8_t *const *frames, const int *strides,
                 uint8_t *out, int out_stride);
/* Same for `batch` frame-sets already resident in DEVICE memory, packed
 * [batch][num_images][src_height][src_width][3]; out packed [batch][cut_h][cut_w][3].
 * Asynchronous on `stream` (a cudaStream_t; NULL = legacy default stream).
 * Alignment: a 16-byte aligned frames_dev takes the staged 16-byte-vector / TMA kernels; any other address is
 * accepted and runs the generic byte-wise gather (slower, same bytes).
 * Concurrency: a handle owns ONE set of pyramid workspaces -- calls on the same handle must be issued on one
 * stream at a time (or be ordered by the caller); use one handle per concurrent stream, as the reference uses one
 * ocvStitcher per thread (src/master.cpp:314-318). */
int pano_process_device(pano_handle h, const uint8_t *frames_dev, uint8_t *out_dev, int batch, void *stream);
/* Host-memory batch (packed as above): H2D, compose and D2H are pipelined on internal
 * streams.  Pinned host memory gives full PCIe rate.  Synchronous. */
int pano_process_batch(pano_handle h, const uint8_t *frames_host, uint8_t *out_host, int batch);

/* ---------------------------------------------------------------- column-strip split
 * Spatial decomposition of ONE large panorama across GPUs (BASELINE config 4): each
 * rank owns the padded-dst columns [x0, x1) (multiples of 2^num_bands) and runs the compose as
 * `pano_strip_phase_count()` phases; after every phase that reports halo bytes the ranks exchange
 * the pyramid columns next to the strip edges (2 columns of every camera pyramid level, 1 column of
 * every collapsed level) -- e.g. ncclSend/ncclRecv between neighbours -- and unpack them before the
 * next phase.  `margin` = extra level-0 columns computed redundantly on each side (>= 8 with halo
 * exchange; 3*2^num_bands makes the exchange unnecessary: the "redundant halo" variant).
 * The panorama buffer has the full cut size; only the rank's own columns are meaningful.
 * side: 0 = left neighbour, 1 = right neighbour.  One frame-set per call sequence. */
int pano_strip_set_window(pano_handle h, int x0, int x1, int margin);
int pano_strip_phase_count(pano_handle h);
/* needed[i] = 1 when camera i's warped ROI meets the window set by pano_strip_set_window*: the frames of the other cameras
 * are never read by this rank (their part of `frames_dev` may hold anything), so a rank uploads only the cameras it needs. */
int pano_strip_cameras(pano_handle h, int *needed);
int pano_strip_run_phase(pano_handle h, int phase, const uint8_t *frames_dev, uint8_t *pano_dev, void *stream);
size_t pano_strip_halo_bytes(pano_handle h, int phase);
int pano_strip_halo_pack(pano_handle h, int phase, int side, void *buf_dev, void *stream);
int pano_strip_halo_unpack(pano_handle h, int phase, int side, const void *buf_dev, void *stream);

/* ---- hybrid split: thin redundant halos below `split_level`, full width above it, ONE all-gather per panorama.
 * Levels < split_level: the rank computes its strip widened by 3 * 2^split_level level-0 columns (no exchange); levels
 * >= split_level: every rank computes the whole width (4^-split_level of the work), which needs the other ranks' columns of
 * the camera pyramids g[split_level] exactly once.  Per frame-set:
 *   pano_strip_run_phases(h, 0, split_level, ..)                      warp + pyrDown into levels 1 .. split_level
 *   pano_strip_level_pack(h, split_level, x0 >> split_level, n, buf)  own columns -> dense chunk ([cam][plane][row][n] int16)
 *   all-gather of the chunks (e.g. ncclAllGather; chunk_cols = the widest strip)
 *   pano_strip_level_unpack_all(h, split_level, gathered, ..)         every other rank's chunk -> g[split_level]
 *   pano_strip_run_phases(h, split_level, pano_strip_phase_count(h), ..)
 * Bit-identical to the undivided panorama on the rank's own columns. */
int pano_strip_set_window_hybrid(pano_handle h, int x0, int x1, int split_level);
int pano_strip_run_phases(pano_handle h, int first, int last, const uint8_t *frames_dev, uint8_t *pano_dev, void *stream);
size_t pano_strip_level_bytes(pano_handle h, int level, int ncols);
int pano_strip_level_pack(pano_handle h, int level, int col, int ncols, void *buf_dev, void *stream);
int pano_strip_level_unpack_all(pano_handle h, int level, const void *gathered_dev, int chunk_cols, const int *lo, const int *n,
                                int nranks, int self, void *stream);

/* ---- the same exchange over PEER MEMORY (NVLink / NVSwitch) instead of a collective library: every rank owns a
 * mailbox in its HBM; after a phase ONE kernel packs the rank's edge columns and stores them straight into both
 * neighbours' mailboxes (P2P stores), then publishes the frame's sequence number in the neighbour's flag word
 * (system-scope release); before the next phase ONE kernel waits for both neighbours' flags (system-scope acquire) and
 * unpacks.  No host synchronisation, no NCCL call on the data path; bit-identical to the undivided panorama.
 *   create:        allocates the mailbox; ipc_handle64 (64 bytes, may be NULL) receives its cudaIpcMemHandle_t for the
 *                  neighbours' processes (one process per GPU; ship it with any side channel, e.g. torch.distributed)
 *   connect:       maps the neighbour's mailbox on `side` (0 = left, 1 = right); NULL = no neighbour on that side
 *   connect_local: neighbour handle in the same process (tests; single-process multi-GPU)
 *   run_p2p:       one frame-set: begin + every phase with push / wait_unpack in between; asynchronous on `stream`.
 *                  All ranks must call it once per frame-set (a rank waits for its neighbours' columns on the device).
 *                  On a non-default stream the frame's launch sequence is captured once per (frames, panorama)
 *                  buffer pair and replayed as a CUDA graph (the exchange is launch-latency bound).
 *   begin / push / wait_unpack: the same, phase by phase. */
int pano_strip_p2p_create(pano_handle h, void *ipc_handle64, size_t *mailbox_bytes);
int pano_strip_p2p_connect(pano_handle h, int side, const void *ipc_handle64);
int pano_strip_p2p_connect_local(pano_handle h, int side, pano_handle neighbour);
int pano_strip_p2p_begin(pano_handle h, void *stream);
int pano_strip_p2p_push(pano_handle h, int phase, void *stream);
int pano_strip_p2p_wait_unpack(pano_handle h, int phase, void *stream);
int pano_strip_p2p_prepare(pano_handle h, const uint8_t *frames_dev, uint8_t *pano_dev, void *stream);   /* builds the graph only */
int pano_strip_run_p2p(pano_handle h, const uint8_t *frames_dev, uint8_t *pano_dev, void *stream);
/* Every device-side wait of the exchange is bounded (about a second of polling); a neighbour that never publishes
 * raises an error flag instead of hanging the GPU.  Synchronises the handle's device; PANO_ERR if the flag was set
 * (and clears it). */
int pano_strip_p2p_check(pano_handle h);

/* Per-kernel device time of the last pano_process_device call made while profiling was
 * enabled (CUDA events on the launching stream).  names: up to max entries. */
int pano_profile_enable(pano_handle h, int on);
int pano_profile_read(pano_handle h, int max, const char **names, float *ms, int *launches, double *alg_bytes);
/* number of kernel launches issued by the last process call */
int pano_last_launch_count(pano_handle h);

/* ---------------------------------------------------------------- nvCam front end
 * nvCam::read_frame's pixel pipeline (include/nvcam.hpp:898-929): resize(8UC4) -> drop alpha
 * -> remap INTER_CUBIC over initUndistortRectifyMap maps -> crop rect -> resize, followed by
 * getFrame(.., src=false)'s resize to the stitcher input (:1092-1094). */
#define PANO_SRC_BGRA 0   /* 8UC4: the VIC's ABGR32 output (NvBufferTransform, include/nvcam.hpp:889-893) */
#define PANO_SRC_YUYV 1   /* 8UC2 YUYV 4:2:2 as captured: the YUYVCAM build converts it with cv::cvtColor(COLOR_YUV2BGRA_YUYV)
                             (include/nvcam.hpp:880-886); here that conversion runs on the device, so the caller
                             transfers 2 bytes per pixel instead of 4 */
typedef struct pano_frontend_config {
    int cam_src_width, cam_src_height;   /* stCamCfg::camSrcWidth/Height: camera frame (8UC4, or 8UC2 with PANO_SRC_YUYV) */
    int undist_width, undist_height;     /* stCamCfg::undistoredWidth/Height */
    int out_width, out_height;           /* stCamCfg::outPutWidth/Height */
    int undistort;                       /* stCamCfg::undistor */
    double K[9];                         /* m_cameraK */
    double D[4];                         /* m_cameraDistorParams k1,k2,p1,p2 */
    double newK[9];                      /* getOptimalNewCameraMatrix(K,D,size,1,size,0) (:831) */
    int rect[4];                         /* m_rectPara x,y,w,h (:916) */
    const float *mapx, *mapy;            /* optional: caller-supplied m_mapx/m_mapy (undist size) */
    int device;
    int max_batch;
    int src_format;                      /* PANO_SRC_*; frames are [cam_src_height][cam_src_width][4 or 2] */
} pano_frontend_config;

int pano_frontend_create(const pano_frontend_config *cfg, pano_frontend_handle *out);
int pano_frontend_destroy(pano_frontend_handle h);
const char *pano_frontend_last_error(pano_frontend_handle h);
/* m_mapx / m_mapy as prepareUndistorMap builds them (include/nvcam.hpp:823-833) */
int pano_frontend_get_maps(pano_frontend_handle h, float *mapx, float *mapy);
/* frames: [batch][cam_src_height][cam_src_width][4 (2 for YUYV)] -> out [batch][out_height][out_width][3].
 * A 16-byte aligned argb_dev takes the tiled (TMA-staged) kernels; other addresses run the generic ones. */
int pano_frontend_process_device(pano_frontend_handle h, const uint8_t *argb_dev, uint8_t *out_dev,
                                 int batch, void *stream);
int pano_frontend_process(pano_frontend_handle h, const uint8_t *argb_host, int stride,
                          uint8_t *out_host, int out_stride);

/* ---------------------------------------------------------------- host-only table builders
 * The init-time products of ocvStitcher::initSeam / nvCam::prepareUndistorMap, computed on the
 * host exactly as the handle does internally.  None of these needs a CUDA device. */
/* blender_warper->warpRoi (include/ocvstitcher.hpp:1057): roi = x,y,w,h */
int pano_host_warp_roi(int warp_kind, float scale, const float *K, const float *R, int src_w, int src_h, int *roi);
/* RotationWarperBase::buildMaps: xmap/ymap sized roi.h x roi.w */
int pano_host_build_maps(int warp_kind, float scale, const float *K, const float *R, int src_w, int src_h,
                         float *xmap, float *ymap);
/* resultRoi + MultiBandBlender::prepare/feed rectangle arithmetic (reached from :1198,:1202).
 * feed_rects: x,y,w,h in padded-dst coordinates; borders: top,bottom,left,right */
int pano_host_blend_geometry(int n, const int *corners, const int *sizes, int num_bands, int *dst_roi,
                             int *eff_bands, int *padded_wh, int *feed_rects, int *borders);
/* BORDER_REFLECT folded into an in-range sample position 32*i'+f' (map folding) */
unsigned pano_host_fold_reflect(int i, int f, int n);
/* cv::convertMaps view of float maps (what cv::remap uses internally) */
int pano_host_fixed_maps(const float *xmap, const float *ymap, size_t count, int16_t *ixy, uint16_t *frac);
/* cv::pyrDown on CV_32F as the handle evaluates it for the weight pyramids */
int pano_host_pyrdown_f32(const float *src, int w, int h, float *dst);
/* FeatherBlender weight: min(distanceTransform(mask, DIST_L1, 3) * sharpness, 1) */
int pano_host_feather_weight(const uint8_t *mask, int w, int h, int stride, float sharpness, float *out);
/* cv::initUndistortRectifyMap(K, D, I, newK, size, CV_32FC1) (include/nvcam.hpp:832) */
int pano_host_undistort_maps(const double *K, const double *D, const double *newK, int w, int h, float *mapx, float *mapy);
/* cv::remap's INTER_CUBIC 15-bit table (1024 x 16) and cv::resize's INTER_LINEAR axis tables */
int pano_host_cubic_table(int16_t *tab);
int pano_host_resize_axis(int ssize, int dsize, int clamp_frac, int *ofs, int16_t *a0, int16_t *a1);
/* cv::resize's INTER_LINEAR_EXACT axis table (8.8 fixed point; pano_set_seam_mask): weight of src[ofs+1] is c1/256 */
int pano_host_linear_exact_axis(int ssize, int dsize, int *ofs, int *c1);

/* The seam finder's inputs for ONE camera exactly as initSeam / updateMask build them (include/ocvstitcher.hpp:985-1017,
 * 1228-1242): the frame resized by seam_work_aspect = min(1, sqrt(1e5 / (w*h))) with INTER_LINEAR_EXACT, warped at that
 * scale (K scaled in float32, warper scale float(scale * aspect), INTER_LINEAR / BORDER_REFLECT), and the all-255 mask
 * warped with INTER_NEAREST / BORDER_CONSTANT.  roi = x,y,w,h of the low-resolution warp (corner = what the seam finder
 * gets as corners[i]); image_warped: roi.h x roi.w x 3 (tight), mask_warped: roi.h x roi.w; both may be NULL (geometry
 * only).  The seam search itself (cv::detail::GraphCutSeamFinder::find) stays with the application; its result goes
 * to pano_set_seam_mask.  Host only, init-time. */
double pano_host_seam_scale(int src_w, int src_h);
int pano_host_seam_input(int warp_kind, float warped_image_scale, const float *K, const float *R, const uint8_t *frame, int src_w,
                         int src_h, int stride, int *roi, uint8_t *image_warped, uint8_t *mask_warped);

/* Chain a front end in front of a stitcher handle (cam = -1: every camera): from then on
 * pano_process / pano_process_device / pano_process_batch take the 8UC4 camera frames
 * ([batch][num_images][cam_src_height][cam_src_width][4]) and run nvCam's pixel pipeline on the
 * device first -- the loop of src/master.cpp:300-318 (getFrame x N, then process) as one call.
 * f == NULL detaches.  Re-callable at any time: the host staging buffers are re-sized when the caller-side frame
 * format changes; a failing call leaves the handle untouched.  The stitcher handle keeps its own copy of the front
 * end's intermediate buffers, so ONE front-end handle may be attached to several stitcher handles that run on
 * different threads / streams at the same time (the two rings of src/panocamimpl.cpp share the camera model); the
 * front end's own host-buffer calls (pano_frontend_process*) remain one thread at a time. */
int pano_attach_frontend(pano_handle h, int cam, pano_frontend_handle f);

/* ---------------------------------------------------------------- two-ring epilogue (caller step after process)
 * What the callers do with the upper- and lower-ring panoramas right after the two process calls:
 * PANO_RING_RESIZE  src/master.cpp:321-326      cv::resize(up, up, down.size()) [INTER_LINEAR]; cv::vconcat(up, down, ret);
 *                                               cv::rectangle(ret, Rect(0, ret.rows/2 - bar/2, ret.cols, bar), 0, -1)   (bar = 10)
 * PANO_RING_CROP    src/panocamimpl.cpp:354-360 both cropped to Rect(0, finalcut, min width, min height - 2*finalcut);
 *                                               cv::vconcat; cv::rectangle(ret, Rect(0, height - bar/2, width, bar), 0, -1) (bar = 4)
 * One kernel writes the stacked frame (resize, copy and bar fused); bit-exact with the OpenCV calls. */
#define PANO_RING_RESIZE 0
#define PANO_RING_CROP 1
typedef struct pano_ring_ctx *pano_ring_handle;
typedef struct pano_ring_config {
    int up_width, up_height;       /* stitcherOut[0] / `up` (8UC3) */
    int down_width, down_height;   /* stitcherOut[1] / `down` */
    int mode;                      /* PANO_RING_* */
    int finalcut;                  /* PANO_RING_CROP only (panocamimpl's finalcut) */
    int bar;                       /* separator rows */
    int device;
} pano_ring_config;
int pano_ring_create(const pano_ring_config *cfg, pano_ring_handle *out);
int pano_ring_destroy(pano_ring_handle h);
const char *pano_ring_last_error(pano_ring_handle h);
int pano_ring_out_size(pano_ring_handle h, int *wh);
/* batch pairs in DEVICE memory: images `stride` bytes per row, consecutive images stride*height bytes apart; asynchronous */
int pano_ring_compose_device(pano_ring_handle h, const uint8_t *up_dev, int up_stride, const uint8_t *down_dev, int down_stride,
                             uint8_t *out_dev, int out_stride, int batch, void *stream);
/* one pair in HOST memory; synchronous */
int pano_ring_compose(pano_ring_handle h, const uint8_t *up_host, int up_stride, const uint8_t *down_host, int down_stride,
                      uint8_t *out_host, int out_stride);

/* ---------------------------------------------------------------- fit to the display canvas (caller step after the epilogue)
 * nvrenderAlpha::fit2final (src/nvrenderAlpha.cpp:153-189): a frame that already has the canvas size is copied;
 * otherwise fitscale = in_width > canvas_width ? canvas_width * 1.0 / in_width : 1, the frame is scaled with
 * cv::resize(input, tmp, Size(), fitscale, fitscale) [INTER_LINEAR, dsize = cvRound(size * fitscale), coordinates mapped
 * with 1 / fitscale] and pasted at ((canvas_w - w) / 2, (canvas_h - h) / 2) of the canvas, which is black elsewhere
 * (canvas.setTo(0), :11).  One kernel writes the whole canvas; bit-exact with the OpenCV calls.  The reference's canvas
 * is 1920 x 1080 (nvbufferWidth x nvbufferHeight). */
typedef struct pano_fit_ctx *pano_fit_handle;
typedef struct pano_fit_config {
    int in_width, in_height;           /* the stacked frame (8UC3) */
    int canvas_width, canvas_height;   /* nvbufferWidth, nvbufferHeight */
    int device;
} pano_fit_config;
int pano_fit_create(const pano_fit_config *cfg, pano_fit_handle *out);
int pano_fit_destroy(pano_fit_handle h);
const char *pano_fit_last_error(pano_fit_handle h);
/* rect: offsetX, offsetY, w, h of the pasted frame (:176-181) */
int pano_fit_geometry(pano_fit_handle h, int *rect, double *fitscale);
int pano_fit_compose_device(pano_fit_handle h, const uint8_t *in_dev, int in_stride, uint8_t *canvas_dev, int canvas_stride,
                            int batch, void *stream);
int pano_fit_compose(pano_fit_handle h, const uint8_t *in_host, int in_stride, uint8_t *canvas_host, int canvas_stride);

/* How an attached front end is evaluated.
 * PANO_FRONTEND_SEQUENTIAL (default, the parity mode): the reference's own order -- undistort (INTER_CUBIC) ->
 *   crop -> resize -> rotation warp, every stage rounding to 8 bits like nvCam::read_frame + ocvStitcher::process
 *   (include/nvcam.hpp:898-929, include/ocvstitcher.hpp:1171); bit-exact with the OpenCV CPU path.
 * PANO_FRONTEND_FUSED: the coordinate maps of all stages are composed on the host into ONE remap table per camera and
 *   the warp gathers bilinearly straight from the 8UC4 camera frame -- a single-gather variant that is NOT bit-exact
 *   (one interpolation instead of three); report it separately with its own PSNR.  pano_get_warp_maps /
 *   pano_get_fixed_maps then return the composed maps (into the camera frame).  Re-callable; needs a front end. */
#define PANO_FRONTEND_SEQUENTIAL 0
#define PANO_FRONTEND_FUSED 1
int pano_set_frontend_mode(pano_handle h, int mode);

/* library / build identification: returns e.g. "panob200 sm_100a" */
const char *pano_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PANOB200_H */
