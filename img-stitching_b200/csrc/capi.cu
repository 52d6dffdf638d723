// C ABI implementation (include/panob200.h): context, init-time table build + upload,
// wave scheduling of the kernels, host<->device pipelining.  No CPU compute fallback exists:
// every process entry point launches CUDA kernels or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/panob200.h"
#include "geometry.hpp"
#include "host_pool.hpp"
#include "pano_dev.h"

using namespace pano;

// internal front-end entry points (frontend.cu)
struct pano_front_scratch;
int pano_frontend_run(pano_frontend_handle h, const uint8_t *argb, size_t in_img, uint8_t *out, size_t o_img, int batch,
                      cudaStream_t st, int out_px, pano_front_scratch *scratch);
pano_front_scratch *pano_frontend_scratch_create(pano_frontend_handle h);
void pano_frontend_scratch_destroy(pano_front_scratch *s);
bool pano_frontend_can_words(pano_frontend_handle h);
int pano_frontend_max_batch(pano_frontend_handle h);
int pano_frontend_launches(pano_frontend_handle h);
void pano_frontend_sizes(pano_frontend_handle h, int *in_wh, int *out_wh);
bool pano_frontend_set_prof(pano_frontend_handle h, cudaEvent_t *ev, double *cubic_bytes, double *resize_bytes, int out_px);
void pano_frontend_backmap(pano_frontend_handle h, double *xs, double *ys, size_t count);
int pano_frontend_in_px(pano_frontend_handle h);
int pano_frontend_convert(pano_frontend_handle h, const uint8_t *yuyv, size_t in_img, uint8_t *dst, int count, cudaStream_t st);

namespace {

thread_local std::string g_create_error;

struct ProfEntry {
    const char *name;
    cudaEvent_t e0, e1;
    double bytes;
    bool shared_e0 = false;     // e0 belongs to the previous entry (do not destroy twice)
};

constexpr int kPipeDepth = 4;

}  // namespace

constexpr int kBounceBands = 4;      // row bands of the panorama / chunks per camera frame on the pageable-buffer path

struct pano_ctx {
    pano_config cfg{};
    std::vector<float> K, R;
    int device = 0;
    int n = 0;
    int blender = PANO_BLEND_MULTIBAND;
    std::string err;

    // geometry
    std::vector<Rect> rois;            // warped roi per camera (corner + size)
    Rect dst_roi;
    int nb = 0, pad_w = 0, pad_h = 0;
    std::vector<FeedRect> feed;
    std::vector<std::vector<float>> xmap, ymap;   // float maps (host, kept for inspection)
    std::vector<std::vector<uint8_t>> mask;       // current blend masks (sizes[cam])
    std::vector<uint8_t *> cam_full;              // device: warped all-255 mask per camera (m_compensatorMaskWarped), tight
    uint8_t *seam_tmp = nullptr;                  // device scratch of pano_set_seam_mask: low-res mask + result + axis tables
    size_t seam_tmp_bytes = 0;
    bool map64 = false;

    // device tables
    PanoTables host{};
    PanoTables *dev = nullptr;
    KernelChoice kc;
    bool tables_dirty = true;
    // optional nvCam front end per camera (pano_attach_frontend): inputs become 8UC4 camera frames
    pano_frontend_handle front[kMaxCams] = {};
    // this handle's own set of the front ends' intermediate buffers (cameras that share a front end share the set: they
    // run one after the other on one stream): a front-end handle can then serve several stitchers at the same time
    pano_front_scratch *front_scratch[kMaxCams] = {};
    bool has_front = false;
    // fused front-end mode (pano_set_frontend_mode): the whole nvCam pipeline is folded into the warp's remap table
    // and the gather reads the 8UC4 camera frames directly -- NOT bit-exact with the sequential path
    bool fused = false;
    std::vector<std::vector<float>> fxmap, fymap;  // composed float maps (rois[cam] large) while fused
    size_t in_frame_bytes = 0;                    // bytes of one input frame as the caller passes it
    size_t in_frame_bytes4 = 0;                   // the same frame as 8UC4 (what the fused gather reads)
    uint8_t *fused_in = nullptr;                  // fused mode + YUYV ingest: converted 8UC4 frames [max_batch][n]
    uint8_t *front_out = nullptr;                 // [max_batch][n][H][W][3 or 4] stitcher inputs produced by the front end
    bool front_px4 = false;                       // the front ends hand over one word per pixel (warp_tile_kernel<.., kSrc4> + TMA)
    // pano_strip_run_phases: captured launch sequences of phase ranges, keyed on (range, buffers); dropped with the tables
    struct PhaseGraph { int first, last; const void *frames; void *pano; cudaGraphExec_t exec; int launches; };
    std::vector<PhaseGraph> phase_graphs;
    cudaGraphExec_t graph1 = nullptr;             // pano_process: the kernel chain of ONE frame-set (stage_in[0] -> stage_out[0])
    int graph1_launches = 0;
    SideStream side{};                              // second stream for the seam-tile kernel of a collapse level (launch_collapse)
    // pano_process, overlapped form: camera i's chain (front end, warp, pyrDown levels) is replayed as soon as ITS frame has
    // landed (copies on s_h2d, one event per camera), the tail (coarsest + collapse) after the last one
    cudaEvent_t ev_cam_in[kMaxCams] = {};
    cudaGraphExec_t graph_cam[kMaxCams] = {}, graph_tail = nullptr;
    int graph_split_launches = 0;
    // pano_process with PAGEABLE host buffers: pinned bounce buffers filled / drained by worker threads (host_pool.hpp)
    HostPool *pool = nullptr;
    uint8_t *pin_in = nullptr, *pin_out = nullptr;
    size_t pin_in_bytes = 0, pin_out_bytes = 0;
    cudaEvent_t ev_band[kBounceBands] = {};
    int strip_x0 = 0, strip_x1 = 0;               // own dst columns (level 0, padded coords); full width = no split
    // walker tiles (kWalkTileW x kWalkTileH): [level][cam][tile] -> any non-zero weight / count of weights == 1
    std::vector<std::vector<std::vector<uint8_t>>> walk_nz;
    std::vector<std::vector<std::vector<int>>> walk_ones;
    std::vector<uint32_t *> d_walk_list, d_gen_list;
    uint8_t *d_stat_nz = nullptr;                 // walker-tile statistics of one camera, all levels (device build of the weights)
    int *d_stat_ones = nullptr;
    std::vector<void *> owned;                    // device allocations to free
    std::vector<void *> cam_mask0, cam_gain;      // per camera, re-uploadable
    std::vector<std::vector<void *>> cam_wt;      // per camera per level

    // peer-memory halo exchange (pano_strip_p2p_*): own mailbox (flags + halo slots), the neighbours' mapped mailboxes
    uint8_t *mailbox = nullptr;
    size_t mailbox_bytes = 0;
    std::vector<size_t> mail_off;                 // [phase][from-side][parity] -> byte offset of the slot
    uint8_t *peer_mail[2] = {nullptr, nullptr};   // left / right neighbour's mailbox as addressable from this device
    bool peer_ipc[2] = {false, false};            // mapped with cudaIpcOpenMemHandle (to be closed)
    unsigned *p2p_counters = nullptr;             // [0], [1] block-completion counters of the push kernels, [2] error flag (a spin ran out)
    int p2p_resident = 0;                         // most halo_exchange_kernel blocks that are certainly co-resident (fused exchange)
    uint32_t *p2p_seq = nullptr;                  // device: frame sequence number (bumped by the first kernel of a frame)
    // one frame's launch sequence captured once and replayed (the exchange is launch-latency bound)
    cudaGraphExec_t p2p_graph = nullptr;
    const void *p2p_graph_frames = nullptr, *p2p_graph_pano = nullptr;

    // staging for host entry points
    uint8_t *stage_in[kPipeDepth] = {};
    uint8_t *stage_out[kPipeDepth] = {};
    cudaStream_t s_h2d = nullptr, s_compute = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_in[kPipeDepth]{}, ev_done[kPipeDepth]{}, ev_out[kPipeDepth]{};

    // profiling
    bool profiling = false;
    std::vector<ProfEntry> prof;
    int last_launches = 0;

    size_t frame_bytes() const { return (size_t)cfg.src_width * cfg.src_height * 3; }
    size_t front_out_bytes() const { return (size_t)cfg.src_width * cfg.src_height * 4; }   // per frame, sized for the word hand-over
    size_t front_frame_bytes() const { return (size_t)cfg.src_width * cfg.src_height * (front_px4 ? 4 : 3); }
    size_t gather_frame_bytes() const { return fused ? in_frame_bytes4 : (has_front ? front_frame_bytes() : frame_bytes()); }   // frame the warp gathers from
    size_t set_bytes() const { return (has_front ? in_frame_bytes : frame_bytes()) * n; }   // caller-side frame-set
    size_t out_bytes() const { return (size_t)host.cut_w * host.cut_h * 3; }
};

namespace {

int fail(pano_ctx *h, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return PANO_ERR;
}

#define CK(h, call)                                                                        \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline int roundUp(int v, int m) { return (v + m - 1) / m * m; }

int reflectIdx(int p, int n)  // BORDER_REFLECT
{
    if (n == 1) return 0;
    while ((unsigned)p >= (unsigned)n) p = p < 0 ? -p - 1 : 2 * n - 1 - p;
    return p;
}

template <typename T>
int devAlloc(pano_ctx *h, T **p, size_t count, bool zero = true)
{
    CK(h, cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)));
    if (zero) CK(h, cudaMemset(*p, 0, std::max<size_t>(count, 1) * sizeof(T)));
    h->owned.push_back(*p);
    return PANO_OK;
}

// upload a w x h host array into a pitched device array
template <typename T>
int upload2d(pano_ctx *h, T *dst, int dpitch, const T *src, int spitch, int w, int hh)
{
    CK(h, cudaMemcpy2D(dst, (size_t)dpitch * sizeof(T), src, (size_t)spitch * sizeof(T), (size_t)w * sizeof(T), hh,
                       cudaMemcpyHostToDevice));
    return PANO_OK;
}

inline int walkTilesX(const pano_ctx *h, int l) { return ((h->pad_w >> l) + kWalkTileW - 1) / kWalkTileW; }
inline int walkTilesY(const pano_ctx *h, int l) { return ((h->pad_h >> l) + kWalkTileH - 1) / kWalkTileH; }

// record which 256x8 dst tiles of level l see a non-zero weight of camera `cam`
// per walker tile of level l: does camera `cam` have any non-zero weight there, and how many of its weights
// are exactly one (`one` = the value that means weight 1.0f: 255 for the 8-bit level-0 mask)
template <typename T>
void markTiles(pano_ctx *h, int cam, int l, const T *data, int w, int hh, int pitch, T one)
{
    if (h->walk_nz.empty()) return;
    const CamTables &C = h->host.cam[cam];
    const int ox = C.rx >> l, oy = C.ry >> l, wtx = walkTilesX(h, l);
    std::vector<uint8_t> &nz = h->walk_nz[l][cam];
    std::vector<int> &ones = h->walk_ones[l][cam];
    std::fill(nz.begin(), nz.end(), 0);
    std::fill(ones.begin(), ones.end(), 0);
    for (int y = 0; y < hh; ++y) {
        const T *row = data + (size_t)y * pitch;
        const int wty = (oy + y) / kWalkTileH;
        for (int x = 0; x < w; ++x) {
            if (row[x] == (T)0) continue;
            const size_t wt = (size_t)wty * wtx + (ox + x) / kWalkTileW;
            nz[wt] = 1;
            if (row[x] == one) ++ones[wt];
        }
    }
}

// Collapse work lists (PanoTables::walk_list / gen_list) from the per-camera tile statistics.
int uploadTileLists(pano_ctx *h)
{
    if (h->walk_nz.empty()) return PANO_OK;
    static const bool no_walk = getenv("PANO_NO_WALK") != nullptr;      // A/B switch: everything through collapse8_kernel
    PanoTables &T = h->host;
    for (int l = 0; l < h->nb; ++l) {
        T.walk_list[l] = nullptr; T.gen_list[l] = nullptr; T.walk_n[l] = 0; T.gen_n[l] = 0;
        if (!h->kc.collapse8[l]) continue;
        const bool walk_ok = (T.unit_norm_exact & 1) && !no_walk;
        const int wf = h->pad_w >> l, hf = h->pad_h >> l, wtx = walkTilesX(h, l), wty = walkTilesY(h, l);
        std::vector<uint32_t> walk, gen;
        size_t n_unit = 0, n_empty = 0, n_gen_by_cams[4] = {0, 0, 0, 0};
        for (int ty = 0; ty < wty; ++ty) {
            if (l == 0 && (ty * kWalkTileH >= T.cut_y + T.cut_h || (ty + 1) * kWalkTileH <= T.cut_y)) continue;
            for (int tx = 0; tx < wtx; ++tx) {
                const size_t t = (size_t)ty * wtx + tx;
                const int npx = (std::min(wf, (tx + 1) * kWalkTileW) - tx * kWalkTileW) *
                                (std::min(hf, (ty + 1) * kWalkTileH) - ty * kWalkTileH);
                uint32_t mask = 0;
                int ncam = 0, cam = -1;
                for (int c = 0; c < h->n; ++c)
                    if (h->walk_nz[l][c][t]) { ++ncam; cam = c; mask |= 1u << c; }
                const uint32_t pos = (uint32_t)tx | ((uint32_t)ty << 12);
                if (walk_ok && ncam == 0) { walk.push_back(pos | ((uint32_t)kWalkEmpty << 24)); ++n_empty; }
                else if (walk_ok && ncam == 1 && h->walk_ones[l][cam][t] == npx &&
                         (l > 0 || T.cam[cam].use_wt0 || (T.unit_norm_exact & 2))) { walk.push_back(pos | ((uint32_t)(1 + cam) << 24)); ++n_unit; }
                else { gen.push_back(pos | (mask << 24)); ++n_gen_by_cams[std::min(ncam, 3)]; }
            }
        }
        if (!walk.empty()) CK(h, cudaMemcpy(h->d_walk_list[l], walk.data(), walk.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        if (!gen.empty()) CK(h, cudaMemcpy(h->d_gen_list[l], gen.data(), gen.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        T.walk_list[l] = h->d_walk_list[l]; T.walk_n[l] = (int)walk.size();
        T.gen_list[l] = h->d_gen_list[l]; T.gen_n[l] = (int)gen.size();
        if (getenv("PANO_DEBUG"))
            fprintf(stderr, "[panob200] level %d: %dx%d tiles of %dx%d: unit-weight %zu, empty %zu, generic %zu (cameras with weight: "
                            "0: %zu, 1: %zu, 2: %zu, 3+: %zu)\n", l, wtx, wty, kWalkTileW, kWalkTileH, n_unit, n_empty, gen.size(),
                    n_gen_by_cams[0], n_gen_by_cams[1], n_gen_by_cams[2], n_gen_by_cams[3]);
    }
    return PANO_OK;
}

template <typename T>
void devFree(pano_ctx *h, T *p)
{
    if (!p) return;
    auto it = std::find(h->owned.begin(), h->owned.end(), (void *)p);
    if (it != h->owned.end()) h->owned.erase(it);
    cudaFree((void *)p);
}

// Device remap table of camera `cam` from float backward maps xm/ym (rois[cam] large) into a W x H source: the
// folded fixed-point table over the feed rect and the per-tile source footprints of the staged gather.  Re-callable
// (the fused front-end mode swaps the maps); the previous tables are released.
int buildCamMap(pano_ctx *h, int cam, const float *xm, const float *ym, int W, int H)
{
    CamTables &C = h->host.cam[cam];
    const FeedRect &fr = h->feed[cam];
    const Rect &img = h->rois[cam];
    const bool map64 = (32 * (W - 1) + 31 > 65535) || (32 * (H - 1) + 31 > 65535);
    C.rx = fr.rect.x; C.ry = fr.rect.y; C.rw = fr.rect.w; C.rh = fr.rect.h;
    C.map_pitch = roundUp(C.rw, 64);
    // folded map over the feed rect: copyMakeBorder(BORDER_REFLECT) of the warped image is a
    // re-read of the warp at the mirrored coordinate
    std::vector<uint32_t> m32;
    std::vector<uint2> m64;
    if (map64) m64.assign((size_t)C.map_pitch * C.rh, make_uint2(0, 0));
    else m32.assign((size_t)C.map_pitch * C.rh, 0u);
    for (int Y = 0; Y < C.rh; ++Y) {
        const int y = reflectIdx(Y - fr.top, img.h);
        for (int X = 0; X < C.rw; ++X) {
            const int x = reflectIdx(X - fr.left, img.w);
            const FixedCoord fc = toFixed(xm[(size_t)y * img.w + x], ym[(size_t)y * img.w + x]);
            const uint32_t sx = foldReflect(fc.ix, fc.fx, W), sy = foldReflect(fc.iy, fc.fy, H);
            if (map64) m64[(size_t)Y * C.map_pitch + X] = make_uint2(sx, sy);
            else m32[(size_t)Y * C.map_pitch + X] = sx | (sy << 16);
        }
    }
    devFree(h, C.map32); devFree(h, C.map64); devFree(h, C.tiles);
    C.map32 = nullptr; C.map64 = nullptr; C.tiles = nullptr;
    if (map64) {
        uint2 *d = nullptr;
        if (devAlloc(h, &d, m64.size(), false)) return PANO_ERR;
        if (cudaMemcpy(d, m64.data(), m64.size() * sizeof(uint2), cudaMemcpyHostToDevice) != cudaSuccess) return fail(h, "map upload failed");
        C.map64 = d;
    } else {
        uint32_t *d = nullptr;
        if (devAlloc(h, &d, m32.size(), false)) return PANO_ERR;
        if (cudaMemcpy(d, m32.data(), m32.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) return fail(h, "map upload failed");
        C.map32 = d;
    }
    // source footprint of every 128x16 output tile (staged-gather warp kernel)
    C.tiles_x = (C.rw + kWarpTileW - 1) / kWarpTileW;
    C.tiles_y = (C.rh + kWarpTileH - 1) / kWarpTileH;
    std::vector<int4> tl((size_t)C.tiles_x * C.tiles_y);
    for (int ty = 0; ty < C.tiles_y; ++ty)
        for (int tx = 0; tx < C.tiles_x; ++tx) {
            int x0 = INT32_MAX, x1 = -1, y0 = INT32_MAX, y1 = -1;
            const int xe = std::min(C.rw, (tx + 1) * kWarpTileW);
            for (int Y = ty * kWarpTileH; Y < std::min(C.rh, (ty + 1) * kWarpTileH); ++Y)
                for (int X = tx * kWarpTileW; X < xe; ++X) {
                    uint32_t sx, sy;
                    if (map64) { sx = m64[(size_t)Y * C.map_pitch + X].x; sy = m64[(size_t)Y * C.map_pitch + X].y; }
                    else { sx = m32[(size_t)Y * C.map_pitch + X] & 0xffffu; sy = m32[(size_t)Y * C.map_pitch + X] >> 16; }
                    const int ix = sx >> 5, iy = sy >> 5;
                    // taps (ix+1, iy+1) may lie one past the frame (weight 0 there); the kernel's
                    // staging loop clamps the source address, the box keeps the unclamped extent
                    x0 = std::min(x0, ix); x1 = std::max(x1, ix + 1);
                    y0 = std::min(y0, iy); y1 = std::max(y1, iy + 1);
                }
            const int px0 = x0 / 16 * 16, groups = (x1 - px0) / 16 + 1;
            const int rows = y1 - y0 + 1;
            int4 d = make_int4(0, 0, 0, 0);
            // capacity as the TMA staging needs it (rows in boxes of 4, pitch >= 128 words); the LDG/STS loop needs less
            if (W % 16 == 0 && groups <= 16 && ((rows + 3) & ~3) * std::max(128, (groups * 16 + 31) & ~31) <= kWarpSmemWords)
                d = make_int4(px0, y0, rows, groups);
            tl[(size_t)ty * C.tiles_x + tx] = d;
        }
    int4 *dt = nullptr;
    if (devAlloc(h, &dt, tl.size(), false)) return PANO_ERR;
    if (cudaMemcpy(dt, tl.data(), tl.size() * sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess) return fail(h, "tile table upload failed");
    C.tiles = dt;
    return PANO_OK;
}

int buildWeights(pano_ctx *h, int cam)
{
    CamTables &C = h->host.cam[cam];
    const Rect &img = h->rois[cam];
    const FeedRect &fr = h->feed[cam];
    const std::vector<uint8_t> &m = h->mask[cam];
    static const bool host_weights = getenv("PANO_HOST_WEIGHTS") != nullptr;    // A/B switch: the host builders
    if (h->blender == PANO_BLEND_FEATHER) {
        if (!host_weights) {
            // device build: the mask goes up once (1 byte per pixel instead of a 4-byte weight map), the separable exact
            // L1 distance transform + sharpness + clamp run as two kernels
            uint8_t *m0d = (uint8_t *)h->cam_mask0[cam];
            int *tmp = nullptr;
            CK(h, cudaMalloc((void **)&tmp, (size_t)img.w * img.h * sizeof(int)));
            cudaError_t e = cudaMemcpy2DAsync(m0d, C.mask_pitch, m.data(), img.w, img.w, img.h, cudaMemcpyHostToDevice, nullptr);
            if (e == cudaSuccess) {
                launch_feather_weight(m0d, C.mask_pitch, img.w, img.h, h->cfg.sharpness, tmp, (float *)h->cam_wt[cam][0], C.wt_pitch[0], nullptr);
                e = cudaStreamSynchronize(nullptr);
            }
            if (e == cudaSuccess) e = cudaGetLastError();
            cudaFree(tmp);
            if (e != cudaSuccess) return fail(h, "feather weight build failed: %s", cudaGetErrorString(e));
            C.use_wt0 = 1;
            return PANO_OK;
        }
        std::vector<float> w((size_t)img.w * img.h);
        featherWeight(m.data(), img.w, img.h, img.w, h->cfg.sharpness, w.data());
        if (upload2d(h, (float *)h->cam_wt[cam][0], C.wt_pitch[0], w.data(), img.w, img.w, img.h)) return PANO_ERR;
        C.use_wt0 = 1;
        return PANO_OK;
    }
    const int W = fr.rect.w, H = fr.rect.h;
    if (!host_weights) {
        // device build: upload the mask into the zeroed feed rect (copyMakeBorder CONSTANT), then the float pyrDown
        // chain and the walker-tile statistics run as kernels; only the few-KB statistics come back
        uint8_t *m0d = (uint8_t *)h->cam_mask0[cam];
        CK(h, cudaMemsetAsync(m0d, 0, (size_t)C.mask_pitch * H, nullptr));
        CK(h, cudaMemcpy2DAsync(m0d + (size_t)fr.top * C.mask_pitch + fr.left, C.mask_pitch, m.data(), img.w, img.w, img.h,
                                cudaMemcpyHostToDevice, nullptr));
        C.use_wt0 = 0;
        const bool mb = h->blender == PANO_BLEND_MULTIBAND;
        const int levels = mb ? h->nb : 0;
        size_t off = 0;
        std::vector<size_t> offs(levels + 1);
        for (int l = 0; l <= levels && mb; ++l) {
            const int lw = W >> l, lh = H >> l;
            if (l > 0)
                launch_weight_pyrdown(l == 1 ? (const void *)m0d : (const void *)h->cam_wt[cam][l - 1], l == 1,
                                      l == 1 ? C.mask_pitch : C.wt_pitch[l - 1], W >> (l - 1), H >> (l - 1),
                                      (float *)h->cam_wt[cam][l], C.wt_pitch[l], nullptr);
            offs[l] = off;
            const int wtx = walkTilesX(h, l), wty = walkTilesY(h, l);
            launch_tile_stats(l == 0 ? (const void *)m0d : (const void *)h->cam_wt[cam][l], l == 0,
                              l == 0 ? C.mask_pitch : C.wt_pitch[l], lw, lh, C.rx >> l, C.ry >> l, wtx, wty,
                              h->d_stat_nz + off, h->d_stat_ones + off, nullptr);
            off += (size_t)wtx * wty;
        }
        if (mb) {
            std::vector<uint8_t> nz(off);
            std::vector<int> ones(off);
            CK(h, cudaMemcpyAsync(nz.data(), h->d_stat_nz, off, cudaMemcpyDeviceToHost, nullptr));
            CK(h, cudaMemcpyAsync(ones.data(), h->d_stat_ones, off * sizeof(int), cudaMemcpyDeviceToHost, nullptr));
            CK(h, cudaStreamSynchronize(nullptr));
            for (int l = 0; l <= levels; ++l) {
                const size_t cnt = h->walk_nz[l][cam].size();
                std::copy(nz.begin() + offs[l], nz.begin() + offs[l] + cnt, h->walk_nz[l][cam].begin());
                std::copy(ones.begin() + offs[l], ones.begin() + offs[l] + cnt, h->walk_ones[l][cam].begin());
            }
        } else {
            CK(h, cudaStreamSynchronize(nullptr));
        }
        CK(h, cudaGetLastError());
        return PANO_OK;
    }
    // level 0: padded 8-bit mask, zero outside the image (copyMakeBorder CONSTANT)
    std::vector<uint8_t> m0((size_t)W * H, 0);
    for (int y = 0; y < img.h; ++y)
        std::memcpy(&m0[(size_t)(y + fr.top) * W + fr.left], &m[(size_t)y * img.w], img.w);
    if (upload2d(h, (uint8_t *)h->cam_mask0[cam], C.mask_pitch, m0.data(), W, W, H)) return PANO_ERR;
    C.use_wt0 = 0;
    markTiles(h, cam, 0, m0.data(), W, H, W, (uint8_t)255);
    if (h->blender != PANO_BLEND_MULTIBAND) return PANO_OK;
    // float weight pyramid (MultiBandBlender::feed: convertTo(CV_32F, 1/255) + pyrDown chain)
    std::vector<float> cur((size_t)W * H);
    for (size_t i = 0; i < cur.size(); ++i) cur[i] = (float)m0[i] * (1.f / 255.f);
    int cw = W, ch = H;
    for (int l = 1; l <= h->nb; ++l) {
        std::vector<float> nxt((size_t)((cw + 1) / 2) * ((ch + 1) / 2));
        pyrDownF32(cur.data(), cw, ch, nxt.data());
        cw = (cw + 1) / 2; ch = (ch + 1) / 2;
        if (upload2d(h, (float *)h->cam_wt[cam][l], C.wt_pitch[l], nxt.data(), cw, cw, ch)) return PANO_ERR;
        markTiles(h, cam, l, nxt.data(), cw, ch, cw, 1.0f);
        cur.swap(nxt);
    }
    return PANO_OK;
}

// the CUDA graphs of pano_process bake in staging pointers, table pointers and launch shapes: they die with any of those
void dropProcessGraphs(pano_ctx *h)
{
    if (h->graph1) { cudaGraphExecDestroy(h->graph1); h->graph1 = nullptr; }
    for (auto &g : h->graph_cam)
        if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    if (h->graph_tail) { cudaGraphExecDestroy(h->graph_tail); h->graph_tail = nullptr; }
}

int syncTables(pano_ctx *h)
{
    if (!h->tables_dirty) return PANO_OK;
    // captured launch sequences bake table pointers, work-list sizes and grid shapes in: they die with the tables
    dropProcessGraphs(h);
    for (auto &g : h->phase_graphs) cudaGraphExecDestroy(g.exec);
    h->phase_graphs.clear();
    if (uploadTileLists(h)) return PANO_ERR;
    CK(h, cudaMemcpy(h->dev, &h->host, sizeof(PanoTables), cudaMemcpyHostToDevice));
    h->tables_dirty = false;
    return PANO_OK;
}

struct Launch {
    pano_ctx *h;
    cudaStream_t st;
    int idx = 0;
    void begin(const char *name, double bytes)
    {
        if (!h->profiling) return;
        ProfEntry pe{};
        pe.name = name; pe.bytes = bytes;
        cudaEventCreate(&pe.e0); cudaEventCreate(&pe.e1);
        cudaEventRecord(pe.e0, st);
        h->prof.push_back(pe);
    }
    void end()
    {
        ++h->last_launches;
        if (!h->profiling) return;
        cudaEventRecord(h->prof.back().e1, st);
    }
};

void clearProf(pano_ctx *h)
{
    for (auto &p : h->prof) { if (!p.shared_e0) cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
    h->prof.clear();
}

// algorithmic bytes per kernel (DESIGN.md "roofline accounting")
double warpBytes(const pano_ctx *h, int slots)
{
    double b = 0;
    for (int i = 0; i < h->n; ++i) {
        const CamTables &C = h->host.cam[i];
        b += (double)C.rw * C.rh * ((h->map64 ? 8 : 4) + 3 + (C.gain_mode == 1 ? 4 : 0)) + (double)h->gather_frame_bytes();
    }
    return b * slots;
}
double pyrdownBytes(const pano_ctx *h, int l, int slots)
{
    double b = 0;
    for (int i = 0; i < h->n; ++i) {
        const CamTables &C = h->host.cam[i];
        b += 3.0 * (C.rw >> l) * (C.rh >> l) + 3.0 * (C.rw >> (l + 1)) * (C.rh >> (l + 1));
    }
    return b * slots;
}
double collapseBytes(const pano_ctx *h, int l, int slots)
{
    double b = 0;
    const int nb = h->nb;
    for (int i = 0; i < h->n; ++i) {
        const CamTables &C = h->host.cam[i];
        const double px = (double)(C.rw >> l) * (C.rh >> l);
        b += px * (3 + ((l == 0 && !C.use_wt0) ? 1 : 4));
        if (l < nb) b += 3.0 * (C.rw >> (l + 1)) * (C.rh >> (l + 1));
    }
    if (l < nb) b += 6.0 * (h->pad_w >> (l + 1)) * (h->pad_h >> (l + 1));
    if (l == 0) b += (double)h->out_bytes();
    else b += 6.0 * (h->pad_w >> l) * (h->pad_h >> l);
    return b * slots;
}
double directBytes(const pano_ctx *h, int slots)
{
    double b = (double)h->out_bytes() + (double)h->set_bytes();
    for (int i = 0; i < h->n; ++i) {
        const CamTables &C = h->host.cam[i];
        b += (double)C.rw * C.rh * ((h->map64 ? 8 : 4) + (h->blender == PANO_BLEND_FEATHER ? 4 : 1) + (C.gain_mode == 1 ? 4 : 0));
    }
    return b * slots;
}

double blendG0Bytes(const pano_ctx *h, int slots)
{
    double b = (double)h->out_bytes();
    for (int i = 0; i < h->n; ++i) {
        const CamTables &C = h->host.cam[i];
        b += (double)C.rw * C.rh * (3 + (h->blender == PANO_BLEND_FEATHER ? 4 : 1));
    }
    return b * slots;
}

// The kernel chain is a sequence of PHASES (one wave = all phases for up to max_batch frame-sets):
//   multiband:  p in [0, nb): (p == 0: warp) + pyrDown level p     -> produces g[p+1]
//               p == nb: coarsest level                            -> produces out[nb]
//               p in (nb, 2nb]: collapse level 2nb - p             -> produces out[L] / the panorama
//   feather / no-blend: a single phase.
// The strip split (pano_strip_*) runs them one at a time with halo exchanges in between.
int phaseCount(const pano_ctx *h) { return h->blender == PANO_BLEND_MULTIBAND ? 2 * h->nb + 1 : 1; }

int runPhase(pano_ctx *h, int p, const uint8_t *frames_dev, uint8_t *out_dev, int slots, cudaStream_t st)
{
    Launch L{h, st};
    static const char *kDown[] = {"pyrdown_l0", "pyrdown_l1", "pyrdown_l2", "pyrdown_l3", "pyrdown_l4",
                                  "pyrdown_l5", "pyrdown_l6", "pyrdown_l7", "pyrdown_l8"};
    static const char *kCol[] = {"collapse_l0", "collapse_l1", "collapse_l2", "collapse_l3", "collapse_l4",
                                 "collapse_l5", "collapse_l6", "collapse_l7", "collapse_l8"};
    const int nb = h->nb;
    if (h->blender != PANO_BLEND_MULTIBAND) {
        static const bool force_direct = getenv("PANO_DIRECT_BLEND") != nullptr;
        const bool windowed = h->strip_x0 != 0 || h->strip_x1 != h->pad_w;
        if (h->kc.warp_tiled && !windowed && !force_direct) {
            // two passes: staged tile gather into the warped images, then a streaming weight / accumulate / normalise pass
            L.begin("warp", warpBytes(h, slots));
            launch_warp(h->dev, h->host, h->kc, frames_dev, slots, st);
            L.end();
            L.begin("blend_warped", blendG0Bytes(h, slots));
            launch_blend_g0(h->dev, h->host, h->blender, out_dev, slots, st);
            L.end();
        } else {
            L.begin("direct_blend", directBytes(h, slots));
            launch_direct_blend(h->dev, h->host, h->blender, frames_dev, out_dev, slots, st);
            L.end();
        }
    } else if (p < nb) {
        if (p == 0) {
            L.begin("warp", warpBytes(h, slots));
            launch_warp(h->dev, h->host, h->kc, frames_dev, slots, st);
            L.end();
        }
        L.begin(kDown[p], pyrdownBytes(h, p, slots));
        launch_pyrdown(h->dev, h->host, h->kc, p, slots, st);
        L.end();
    } else if (p == nb) {
        if (nb == 0) {
            L.begin("warp", warpBytes(h, slots));
            launch_warp(h->dev, h->host, h->kc, frames_dev, slots, st);
            L.end();
        }
        L.begin("coarsest", collapseBytes(h, nb, slots));
        launch_coarsest(h->dev, h->host, out_dev, slots, st);
        L.end();
    } else {
        const int l = 2 * nb - p;
        L.begin(kCol[l], collapseBytes(h, l, slots));
        h->last_launches += launch_collapse(h->dev, h->host, h->kc, l, out_dev, slots, st, h->side.st ? &h->side : nullptr) - 1;
        L.end();
    }
    return PANO_OK;
}

int runFrontEnds(pano_ctx *h, const uint8_t *&frames_dev, int slots, cudaStream_t st)
{
    if (!h->has_front) return PANO_OK;
    Launch L{h, st};
    if (h->fused) {
        // single-gather variant: the warp reads the camera frames itself; only a YUYV ingest still needs its conversion
        if (h->in_frame_bytes4 == h->in_frame_bytes) return PANO_OK;
        L.begin("fe_yuyv_to_bgra", (double)slots * h->n * (h->in_frame_bytes + h->in_frame_bytes4));
        if (pano_frontend_convert(h->front[0], frames_dev, h->in_frame_bytes, h->fused_in, slots * h->n, st))
            return fail(h, "front end: %s", pano_frontend_last_error(h->front[0]));
        L.end();
        frames_dev = h->fused_in;
        return PANO_OK;
    }
    bool same = true;
    for (int i = 1; i < h->n; ++i) same = same && h->front[i] == h->front[0];
    const int opx = h->front_px4 ? 4 : 3;
    const size_t ofb = h->front_frame_bytes();
    // per-kernel timing only when the front end runs the batch as ONE chunk (its events are re-recorded per chunk)
    if (same && h->profiling && slots * h->n <= pano_frontend_max_batch(h->front[0])) {
        // per-kernel timing of the fast path (cubic undistort / bilinear resize)
        cudaEvent_t ev[4];
        double cb = 0, rb = 0;
        for (auto &e : ev) cudaEventCreate(&e);
        const bool yuyv = h->in_frame_bytes4 != h->in_frame_bytes;
        if (pano_frontend_set_prof(h->front[0], ev, &cb, &rb, opx)) {
            const int rc = pano_frontend_run(h->front[0], frames_dev, h->in_frame_bytes, h->front_out, ofb, slots * h->n, st, opx, h->front_scratch[0]);
            pano_frontend_set_prof(h->front[0], nullptr, nullptr, nullptr, opx);
            if (rc) return fail(h, "front end: %s", pano_frontend_last_error(h->front[0]));
            if (yuyv) {
                ProfEntry y{"fe_yuyv_to_bgra", ev[3], ev[0], (double)(h->in_frame_bytes + h->in_frame_bytes4) * slots * h->n};
                h->prof.push_back(y);
            } else {
                cudaEventDestroy(ev[3]);
            }
            {
                ProfEntry a{"fe_cubic_undistort", ev[0], ev[1], cb * slots * h->n, yuyv};
                h->prof.push_back(a);
                ProfEntry b{"fe_resize", ev[1], ev[2], rb * slots * h->n, true};   // shares the middle event
                h->prof.push_back(b);
            }
            h->last_launches += pano_frontend_launches(h->front[0]);
            frames_dev = h->front_out;
            return PANO_OK;
        }
        pano_frontend_set_prof(h->front[0], nullptr, nullptr, nullptr, opx);
        for (auto &e : ev) cudaEventDestroy(e);
    }
    L.begin("front_end", (double)slots * h->n * (h->in_frame_bytes + ofb));
    if (same) {
        if (pano_frontend_run(h->front[0], frames_dev, h->in_frame_bytes, h->front_out, ofb, slots * h->n, st, opx, h->front_scratch[0]))
            return fail(h, "front end: %s", pano_frontend_last_error(h->front[0]));
        h->last_launches += pano_frontend_launches(h->front[0]) - 1;
    } else {
        for (int i = 0; i < h->n; ++i) {
            if (pano_frontend_run(h->front[i], frames_dev + i * h->in_frame_bytes, h->in_frame_bytes * h->n,
                                  h->front_out + i * ofb, ofb * h->n, slots, st, opx, h->front_scratch[i]))
                return fail(h, "front end: %s", pano_frontend_last_error(h->front[i]));
            h->last_launches += pano_frontend_launches(h->front[i]);
        }
        h->last_launches -= 1;
    }
    L.end();
    frames_dev = h->front_out;
    return PANO_OK;
}

// Phases [first, last) of one wave.  (An experiment ran the top of the pyramid -- pyrDown l >= 3, coarsest, collapse l >= 3 -- as
// ONE thread-block-cluster kernel per frame-set slot with cluster barriers between the phases: bit-exact, 5 launches fewer,
// but SLOWER -- 0.372 ms against 0.207 ms per 64-frame-set wave and 0.376 ms against 0.240 ms device time for a single
// frame-set: the separate launches spread every tiny level over all 148 SMs, a cluster of 8 CTAs cannot.  Removed.)
int runPhases(pano_ctx *h, int first, int last, const uint8_t *frames_dev, uint8_t *out_dev, int slots, cudaStream_t st)
{
    for (int p = first; p < last; ++p)
        if (runPhase(h, p, frames_dev, out_dev, slots, st)) return PANO_ERR;
    return PANO_OK;
}

int runWave(pano_ctx *h, const uint8_t *frames_dev, uint8_t *out_dev, int slots, cudaStream_t st)
{
    if (runFrontEnds(h, frames_dev, slots, st)) return PANO_ERR;
    if (runPhases(h, 0, phaseCount(h), frames_dev, out_dev, slots, st)) return PANO_ERR;
    CK(h, cudaGetLastError());
    return PANO_OK;
}

// ---- pano_process, overlapped form (one frame-set, host frames): everything one camera needs before the cameras meet
// in the blend -- its front end, its warp, its Gaussian pyramid -- only reads that camera's frame, so it can run while the
// NEXT camera's frame is still crossing PCIe.  All chains run on ONE stream, in camera order (a camera's chain is ~50 us,
// its copy ~150 us: the stream is idle most of the time anyway, and a front-end handle shared by the cameras is used
// by one of them at a time); the tail (coarsest + collapse levels) follows the last chain.
bool canOverlapCameras(const pano_ctx *h)
{
    static const bool off = getenv("PANO_NO_OVERLAP") != nullptr;       // A/B switch: copy everything, then compute
    const bool full = h->strip_x0 == 0 && h->strip_x1 == h->pad_w;
    return !off && h->blender == PANO_BLEND_MULTIBAND && h->nb >= 1 && h->n >= 2 && !h->fused && full && !h->profiling;
}

int runCameraChain(pano_ctx *h, int cam, const uint8_t *stage, cudaStream_t st)
{
    const uint8_t *fr = stage;
    if (h->has_front) {
        const int opx = h->front_px4 ? 4 : 3;
        const size_t ofb = h->front_frame_bytes();
        if (pano_frontend_run(h->front[cam], stage + (size_t)cam * h->in_frame_bytes, h->in_frame_bytes * h->n,
                              h->front_out + (size_t)cam * ofb, ofb * h->n, 1, st, opx, h->front_scratch[cam]))
            return fail(h, "front end: %s", pano_frontend_last_error(h->front[cam]));
        h->last_launches += pano_frontend_launches(h->front[cam]);
        fr = h->front_out;
    }
    launch_warp(h->dev, h->host, h->kc, fr, 1, st, CamRange{cam, 1});
    for (int l = 0; l < h->nb; ++l) launch_pyrdown(h->dev, h->host, h->kc, l, 1, st, CamRange{cam, 1});
    h->last_launches += 1 + h->nb;
    CK(h, cudaGetLastError());
    return PANO_OK;
}

int runTail(pano_ctx *h, const uint8_t *stage, uint8_t *out_dev, cudaStream_t st)
{
    const uint8_t *fr = h->has_front ? h->front_out : stage;
    if (runPhases(h, h->nb, phaseCount(h), fr, out_dev, 1, st)) return PANO_ERR;
    CK(h, cudaGetLastError());
    return PANO_OK;
}

// capture what `body` launches on `st` into an executable graph
template <typename F>
int captureGraph(pano_ctx *h, cudaStream_t st, cudaGraphExec_t *exec, F body)
{
    cudaGraph_t g = nullptr;
    CK(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = body();
    const cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        (void)cudaGetLastError();
        return rc ? PANO_ERR : fail(h, "pano_process: graph capture failed: %s", cudaGetErrorString(e));
    }
    const cudaError_t ei = cudaGraphInstantiate(exec, g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) { *exec = nullptr; return fail(h, "pano_process: graph instantiation failed: %s", cudaGetErrorString(ei)); }
    return PANO_OK;
}

// halo exchanged after phase p: kind (0 = camera pyramids g, 1 = collapsed pyramid out), level, columns
bool phaseHalo(const pano_ctx *h, int p, int &kind, int &level, int &ncols)
{
    if (h->blender != PANO_BLEND_MULTIBAND) return false;
    const int nb = h->nb;
    if (p < nb) { kind = 0; level = p + 1; ncols = 2; return true; }
    if (p == nb) { kind = 1; level = nb; ncols = 1; return nb > 0; }
    const int l = 2 * nb - p;
    kind = 1; level = l; ncols = 1;
    return l >= 1;
}

int ensureStaging(pano_ctx *h)
{
    if (!h->s_h2d) {
        for (int i = 0; i < kPipeDepth; ++i) {
            CK(h, cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
            CK(h, cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
            CK(h, cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
        }
        for (auto &e : h->ev_cam_in) CK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(h, cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
        CK(h, cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking));
        CK(h, cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    }
    if (h->stage_in[0]) return PANO_OK;
    const size_t S = h->cfg.max_batch;
    for (int i = 0; i < kPipeDepth; ++i) {
        if (devAlloc(h, &h->stage_in[i], h->set_bytes() * S, false)) return PANO_ERR;
        if (devAlloc(h, &h->stage_out[i], h->out_bytes() * S, false)) return PANO_ERR;
    }
    return PANO_OK;
}

// One private set of front-end intermediates per DISTINCT front end of this handle (cameras sharing a front end run one
// after the other on one stream and share the set).  The device must be idle.
void freeFrontScratch(pano_ctx *h)
{
    for (int i = 0; i < kMaxCams; ++i) {
        pano_front_scratch *s = h->front_scratch[i];
        if (!s) continue;
        for (int j = i; j < kMaxCams; ++j)
            if (h->front_scratch[j] == s) h->front_scratch[j] = nullptr;
        pano_frontend_scratch_destroy(s);
    }
}

// The staging buffers are sized from set_bytes(), which changes when a front end is attached or detached (8UC4 / YUYV
// camera frames vs BGR stitcher inputs): release them so that the next host-side call re-creates them at the new size.
// The device must be idle (callers synchronise first).
void dropStaging(pano_ctx *h)
{
    for (int i = 0; i < kPipeDepth; ++i) {
        devFree(h, h->stage_in[i]); devFree(h, h->stage_out[i]);
        h->stage_in[i] = nullptr; h->stage_out[i] = nullptr;
    }
    dropProcessGraphs(h);
}

}  // namespace

// =========================================================================== C ABI

extern "C" {

const char *pano_version(void) { return "panob200 0.1 sm_100a"; }

const char *pano_last_error(pano_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int pano_create(const pano_config *cfg, pano_handle *out)
{
    if (!cfg || !out) return fail(nullptr, "pano_create: null argument");
    *out = nullptr;
    if (cfg->num_images < 1 || cfg->num_images > kMaxCams) return fail(nullptr, "num_images must be 1..%d", kMaxCams);
    if (cfg->src_width < 2 || cfg->src_height < 2) return fail(nullptr, "bad source size");
    if (!cfg->K || !cfg->R) return fail(nullptr, "K/R missing");
    if (cfg->blender < PANO_BLEND_NO || cfg->blender > PANO_BLEND_MULTIBAND) return fail(nullptr, "bad blender kind");
    if (cfg->warp_kind != PANO_WARP_SPHERICAL && cfg->warp_kind != PANO_WARP_CYLINDRICAL) return fail(nullptr, "bad warp kind");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, "no CUDA device: this library has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, "bad device ordinal %d", cfg->device);

    pano_ctx *h = new pano_ctx();
    auto bail = [&](int) { g_create_error = h->err; pano_destroy(h); return PANO_ERR; };
    h->cfg = *cfg;
    h->cfg.max_batch = std::max(1, cfg->max_batch);
    h->n = cfg->num_images;
    h->blender = cfg->blender;
    h->device = cfg->device;
    h->K.assign(cfg->K, cfg->K + 9 * h->n);
    h->R.assign(cfg->R, cfg->R + 9 * h->n);
    h->cfg.K = h->K.data();
    h->cfg.R = h->R.data();
    if (cudaSetDevice(h->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return bail(0); }

    const int W = cfg->src_width, H = cfg->src_height, n = h->n;
    h->map64 = (32 * (W - 1) + 31 > 65535) || (32 * (H - 1) + 31 > 65535);

    // ---- warp tables (RotationWarperBase::warpRoi / buildMaps) ----
    h->rois.resize(n); h->xmap.resize(n); h->ymap.resize(n); h->mask.resize(n);
    for (int i = 0; i < n; ++i) {
        RotationWarper wp(cfg->warp_kind, cfg->warped_image_scale);
        wp.setCamera(&h->K[9 * i], &h->R[9 * i]);
        Rect r = wp.warpRoi(W, H);
        if (r.w <= 0 || r.h <= 0 || (double)r.w * r.h > 4e8) { h->err = "degenerate warp roi (check K/R/scale)"; return bail(0); }
        h->rois[i] = r;
        h->xmap[i].resize((size_t)r.w * r.h);
        h->ymap[i].resize((size_t)r.w * r.h);
        wp.buildMaps(W, H, r, h->xmap[i].data(), h->ymap[i].data());
        // default mask = warp of an all-255 mask with INTER_NEAREST / BORDER_CONSTANT (:1085)
        h->mask[i].resize((size_t)r.w * r.h);
        for (size_t p = 0; p < h->mask[i].size(); ++p) {
            const int ix = std::max(-32768, std::min(32767, (int)lrintf(h->xmap[i][p])));
            const int iy = std::max(-32768, std::min(32767, (int)lrintf(h->ymap[i][p])));
            h->mask[i][p] = ((unsigned)ix < (unsigned)W && (unsigned)iy < (unsigned)H) ? 255 : 0;
        }
    }
    h->dst_roi = resultRoi(h->rois);
    h->cam_full.assign(n, nullptr);
    for (int i = 0; i < n; ++i) {
        if (devAlloc(h, &h->cam_full[i], h->mask[i].size(), false)) return bail(0);
        if (cudaMemcpy(h->cam_full[i], h->mask[i].data(), h->mask[i].size(), cudaMemcpyHostToDevice) != cudaSuccess) { h->err = "mask upload failed"; return bail(0); }
    }

    // ---- blender geometry ----
    PanoTables &T = h->host;
    T.num_cams = n; T.src_w = W; T.src_h = H; T.src_px = 3;
    T.roi_w = h->dst_roi.w; T.roi_h = h->dst_roi.h;
    h->feed.resize(n);
    if (h->blender == PANO_BLEND_MULTIBAND) {
        h->nb = multibandPrepare(h->dst_roi, std::max(0, cfg->num_bands), h->pad_w, h->pad_h);
        if (h->nb + 1 > kMaxLevels) { h->err = "too many bands"; return bail(0); }
        for (int i = 0; i < n; ++i) h->feed[i] = multibandFeedRect(h->dst_roi, h->pad_w, h->pad_h, h->nb, h->rois[i]);
    } else {
        h->nb = 0; h->pad_w = h->dst_roi.w; h->pad_h = h->dst_roi.h;
        for (int i = 0; i < n; ++i) {
            FeedRect fr{};
            fr.rect.x = h->rois[i].x - h->dst_roi.x; fr.rect.y = h->rois[i].y - h->dst_roi.y;
            fr.rect.w = h->rois[i].w; fr.rect.h = h->rois[i].h;
            h->feed[i] = fr;
        }
    }
    T.nb = h->nb; T.pad_w = h->pad_w; T.pad_h = h->pad_h;
    for (int l = 0; l < kMaxLevels; ++l) { T.win_lo[l] = 0; T.win_hi[l] = INT_MAX; }
    h->strip_x0 = 0; h->strip_x1 = h->pad_w;
    if (cfg->cut[2] > 0 && cfg->cut[3] > 0) {
        T.cut_x = cfg->cut[0]; T.cut_y = cfg->cut[1]; T.cut_w = cfg->cut[2]; T.cut_h = cfg->cut[3];
    } else {
        T.cut_x = 0; T.cut_y = 0; T.cut_w = T.roi_w; T.cut_h = T.roi_h;
    }
    if (T.cut_x < 0 || T.cut_y < 0 || T.cut_x + T.cut_w > T.roi_w || T.cut_y + T.cut_h > T.roi_h) {
        h->err = "cut rectangle lies outside the dst roi";  // cv::Mat::operator()(Rect) would assert (:1210)
        return bail(0);
    }

    // ---- per-camera device tables ----
    const int S = h->cfg.max_batch;
    h->cam_mask0.assign(n, nullptr); h->cam_gain.assign(n, nullptr);
    h->cam_wt.assign(n, std::vector<void *>(kMaxLevels, nullptr));
    for (int i = 0; i < n; ++i) {
        CamTables &C = T.cam[i];
        const FeedRect &fr = h->feed[i];
        const Rect &img = h->rois[i];
        C.rx = fr.rect.x; C.ry = fr.rect.y; C.rw = fr.rect.w; C.rh = fr.rect.h;
        if (buildCamMap(h, i, h->xmap[i].data(), h->ymap[i].data(), W, H)) return bail(0);
        C.gain_mode = 0; C.gain_map = nullptr; C.gain_scalar = 1.0;
        // weights
        C.mask_pitch = roundUp(C.rw, 64);
        uint8_t *m0 = nullptr;
        if (devAlloc(h, &m0, (size_t)C.mask_pitch * C.rh + 16)) return bail(0);     // + slack: 4-byte funnel loads may touch the next word
        h->cam_mask0[i] = m0; C.mask0 = m0;
        const int wl0 = (h->blender == PANO_BLEND_MULTIBAND) ? 0 : 0;
        for (int l = wl0; l <= h->nb; ++l) {
            C.wt_pitch[l] = roundUp(std::max(1, C.rw >> l), 32);
            float *w = nullptr;
            if (devAlloc(h, &w, (size_t)C.wt_pitch[l] * std::max(1, C.rh >> l))) return bail(0);
            h->cam_wt[i][l] = w; C.wt[l] = w;
        }
        // pyramid workspace (feather / no-blend: level 0 only = the warped images of the two-pass blend)
        {
            for (int l = 0; l <= h->nb; ++l) {
                const int lw = C.rw >> l, lh = C.rh >> l;
                C.g_pitch[l] = roundUp(lw + 24, 128);  // 8-bit samples, 128-byte aligned rows; packed kernels over-read up to 20 bytes past a row
                C.g_plane[l] = (size_t)C.g_pitch[l] * lh;
                C.g_slot[l] = 3 * C.g_plane[l];
                uint8_t *g = nullptr;
                if (devAlloc(h, &g, C.g_slot[l] * S)) return bail(0);
                C.g[l] = g;
            }
        }
    }
    if (h->blender == PANO_BLEND_MULTIBAND) {
        for (int l = 1; l <= h->nb; ++l) {
            const int lw = h->pad_w >> l, lh = h->pad_h >> l;
            T.out_pitch[l] = roundUp(lw + 8, 64);
            T.out_plane[l] = (size_t)T.out_pitch[l] * lh;
            T.out_slot[l] = 3 * T.out_plane[l];
            int16_t *o = nullptr;
            if (devAlloc(h, &o, T.out_slot[l] * S)) return bail(0);
            T.outp[l] = o;
        }
    }
    // which fast kernels this geometry admits (the generic ones cover everything else)
    h->kc.warp_tiled = (W % 16 == 0);
    if (h->blender == PANO_BLEND_MULTIBAND) {
        for (int l = 0; l < h->nb; ++l) {
            bool ok = true;
            for (int i = 0; i < n; ++i) ok = ok && ((T.cam[i].rw >> l) % 4 == 0) && ((T.cam[i].rh >> l) % 2 == 0) && ((T.cam[i].rw >> l) >= 16);
            h->kc.pyrdown8[l] = ok;
            h->kc.collapse8[l] = (h->nb - l >= 3);
        }
    }
    if (h->blender == PANO_BLEND_MULTIBAND) {
        h->walk_nz.assign(h->nb + 1, std::vector<std::vector<uint8_t>>(n));
        h->walk_ones.assign(h->nb + 1, std::vector<std::vector<int>>(n));
        h->d_walk_list.assign(h->nb + 1, nullptr);
        h->d_gen_list.assign(h->nb + 1, nullptr);
        for (int l = 0; l <= h->nb; ++l) {
            const size_t wcnt = (size_t)walkTilesX(h, l) * walkTilesY(h, l);
            for (int c = 0; c < n; ++c) { h->walk_nz[l][c].assign(wcnt, 0); h->walk_ones[l][c].assign(wcnt, 0); }
            if (devAlloc(h, &h->d_walk_list[l], wcnt) || devAlloc(h, &h->d_gen_list[l], wcnt)) return bail(0);
        }
        size_t total = 0;
        for (int l = 0; l <= h->nb; ++l) total += (size_t)walkTilesX(h, l) * walkTilesY(h, l);
        if (devAlloc(h, &h->d_stat_nz, total) || devAlloc(h, &h->d_stat_ones, total)) return bail(0);
    }
    {
        // exactness of the two kernel shortcuts under this host's IEEE arithmetic (always true on
        // conforming hardware; the kernels fall back to the general path otherwise)
        volatile float one = 1.0f, eps = 1e-5f, m255 = 255.f, inv = 1.f / 255.f;
        const float den = one + eps;
        bool ok = true;
        for (int a = -32768; a <= 32767 && ok; ++a) {
            if (a == 0) continue;
            volatile float q = (float)a / den;
            ok = ((int)(short)(int)q) == a - (a > 0 ? 1 : -1);
        }
        volatile float p = m255 * inv;
        T.unit_norm_exact = (ok ? 1 : 0) | (p == 1.0f ? 2 : 0);
    }
    if (devAlloc(h, &h->dev, 1)) return bail(0);
    if (cudaStreamCreateWithFlags(&h->side.st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->side.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->side.join, cudaEventDisableTiming) != cudaSuccess) return bail(0);
    for (int i = 0; i < n; ++i)
        if (buildWeights(h, i)) return bail(0);
    h->tables_dirty = true;
    *out = h;
    return PANO_OK;
}

int pano_destroy(pano_handle h)
{
    if (!h) return PANO_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    clearProf(h);
    if (h->p2p_graph) cudaGraphExecDestroy(h->p2p_graph);
    dropProcessGraphs(h);
    freeFrontScratch(h);
    if (h->side.st) cudaStreamDestroy(h->side.st);
    if (h->side.fork) cudaEventDestroy(h->side.fork);
    if (h->side.join) cudaEventDestroy(h->side.join);
    for (auto &e : h->ev_cam_in)
        if (e) cudaEventDestroy(e);
    for (auto &g : h->phase_graphs) cudaGraphExecDestroy(g.exec);
    if (h->pool) host_pool_destroy(h->pool);
    if (h->pin_in) cudaFreeHost(h->pin_in);
    if (h->pin_out) cudaFreeHost(h->pin_out);
    for (auto &e : h->ev_band)
        if (e) cudaEventDestroy(e);
    for (int s = 0; s < 2; ++s)
        if (h->peer_mail[s] && h->peer_ipc[s]) cudaIpcCloseMemHandle(h->peer_mail[s]);
    for (void *p : h->owned) cudaFree(p);
    for (int i = 0; i < kPipeDepth; ++i) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_compute) cudaStreamDestroy(h->s_compute);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    delete h;
    return PANO_OK;
}

int pano_get_geometry(pano_handle h, int *corners, int *sizes, int *dst_roi, int *out_wh)
{
    if (!h) return PANO_ERR;
    for (int i = 0; i < h->n; ++i) {
        if (corners) { corners[2 * i] = h->rois[i].x; corners[2 * i + 1] = h->rois[i].y; }
        if (sizes) { sizes[2 * i] = h->rois[i].w; sizes[2 * i + 1] = h->rois[i].h; }
    }
    if (dst_roi) { dst_roi[0] = h->dst_roi.x; dst_roi[1] = h->dst_roi.y; dst_roi[2] = h->dst_roi.w; dst_roi[3] = h->dst_roi.h; }
    if (out_wh) { out_wh[0] = h->host.cut_w; out_wh[1] = h->host.cut_h; }
    return PANO_OK;
}

int pano_get_blend_geometry(pano_handle h, int *num_bands, int *padded_wh, int *feed_rects)
{
    if (!h) return PANO_ERR;
    if (num_bands) *num_bands = h->nb;
    if (padded_wh) { padded_wh[0] = h->pad_w; padded_wh[1] = h->pad_h; }
    if (feed_rects)
        for (int i = 0; i < h->n; ++i) {
            feed_rects[4 * i] = h->feed[i].rect.x; feed_rects[4 * i + 1] = h->feed[i].rect.y;
            feed_rects[4 * i + 2] = h->feed[i].rect.w; feed_rects[4 * i + 3] = h->feed[i].rect.h;
        }
    return PANO_OK;
}

int pano_get_warp_maps(pano_handle h, int cam, float *xmap, float *ymap)
{
    if (!h || cam < 0 || cam >= h->n) return fail(h, "bad camera index");
    const std::vector<float> &xm = h->fused ? h->fxmap[cam] : h->xmap[cam], &ym = h->fused ? h->fymap[cam] : h->ymap[cam];
    if (xmap) std::memcpy(xmap, xm.data(), xm.size() * sizeof(float));
    if (ymap) std::memcpy(ymap, ym.data(), ym.size() * sizeof(float));
    return PANO_OK;
}

int pano_get_fixed_maps(pano_handle h, int cam, int16_t *ixy, uint16_t *frac)
{
    if (!h || cam < 0 || cam >= h->n) return fail(h, "bad camera index");
    const std::vector<float> &xm = h->fused ? h->fxmap[cam] : h->xmap[cam], &ym = h->fused ? h->fymap[cam] : h->ymap[cam];
    const size_t cnt = xm.size();
    for (size_t p = 0; p < cnt; ++p) {
        const FixedCoord fc = toFixed(xm[p], ym[p]);
        if (ixy) { ixy[2 * p] = (int16_t)fc.ix; ixy[2 * p + 1] = (int16_t)fc.iy; }
        if (frac) frac[p] = (uint16_t)(fc.fy * 32 + fc.fx);
    }
    return PANO_OK;
}

int pano_set_mask(pano_handle h, int cam, const uint8_t *mask, int width, int height, int stride)
{
    if (!h || cam < 0 || cam >= h->n || !mask) return fail(h, "pano_set_mask: bad argument");
    const Rect &img = h->rois[cam];
    if (width != img.w || height != img.h || stride < width)
        return fail(h, "pano_set_mask: mask must be %dx%d (got %dx%d)", img.w, img.h, width, height);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());  // tables may be in use by an in-flight wave
    for (int y = 0; y < height; ++y) std::memcpy(&h->mask[cam][(size_t)y * width], mask + (size_t)y * stride, width);
    if (buildWeights(h, cam)) return PANO_ERR;
    h->tables_dirty = true;
    return PANO_OK;
}

int pano_set_seam_mask(pano_handle h, int cam, const uint8_t *seam, int width, int height, int stride)
{
    if (!h || cam < 0 || cam >= h->n || !seam || width < 1 || height < 1 || stride < width)
        return fail(h, "pano_set_seam_mask: bad argument");
    const Rect &img = h->rois[cam];
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());  // tables may be in use by an in-flight wave
    std::vector<int> xo, xc, yo, yc;
    linearExactAxis(width, img.w, xo, xc);
    linearExactAxis(height, img.h, yo, yc);
    // scratch layout: [xo | xc | yo | yc] ints, low-res mask (tight), result (tight)
    const size_t tab_bytes = (size_t)(2 * img.w + 2 * img.h) * sizeof(int);
    const size_t low_bytes = ((size_t)width * height + 15) & ~(size_t)15, res_bytes = (size_t)img.w * img.h;
    const size_t need = tab_bytes + low_bytes + res_bytes;
    if (need > h->seam_tmp_bytes) {
        devFree(h, h->seam_tmp);
        h->seam_tmp = nullptr; h->seam_tmp_bytes = 0;
        if (devAlloc(h, &h->seam_tmp, need, false)) return PANO_ERR;
        h->seam_tmp_bytes = need;
    }
    int *d_xo = reinterpret_cast<int *>(h->seam_tmp), *d_xc = d_xo + img.w, *d_yo = d_xc + img.w, *d_yc = d_yo + img.h;
    uint8_t *d_low = h->seam_tmp + tab_bytes, *d_res = d_low + low_bytes;
    CK(h, cudaMemcpyAsync(d_xo, xo.data(), img.w * sizeof(int), cudaMemcpyHostToDevice, nullptr));
    CK(h, cudaMemcpyAsync(d_xc, xc.data(), img.w * sizeof(int), cudaMemcpyHostToDevice, nullptr));
    CK(h, cudaMemcpyAsync(d_yo, yo.data(), img.h * sizeof(int), cudaMemcpyHostToDevice, nullptr));
    CK(h, cudaMemcpyAsync(d_yc, yc.data(), img.h * sizeof(int), cudaMemcpyHostToDevice, nullptr));
    CK(h, cudaMemcpy2DAsync(d_low, width, seam, stride, width, height, cudaMemcpyHostToDevice, nullptr));
    launch_seam_mask(d_low, width, height, width, d_xo, d_xc, d_yo, d_yc, h->cam_full[cam], d_res, img.w, img.h, nullptr);
    CK(h, cudaMemcpyAsync(h->mask[cam].data(), d_res, res_bytes, cudaMemcpyDeviceToHost, nullptr));
    CK(h, cudaStreamSynchronize(nullptr));
    CK(h, cudaGetLastError());
    if (buildWeights(h, cam)) return PANO_ERR;
    h->tables_dirty = true;
    return PANO_OK;
}

int pano_get_mask(pano_handle h, int cam, uint8_t *mask, int stride)
{
    if (!h || cam < 0 || cam >= h->n || !mask) return fail(h, "pano_get_mask: bad argument");
    const Rect &img = h->rois[cam];
    if (stride < img.w) return fail(h, "pano_get_mask: stride too small");
    for (int y = 0; y < img.h; ++y) std::memcpy(mask + (size_t)y * stride, &h->mask[cam][(size_t)y * img.w], img.w);
    return PANO_OK;
}

int pano_set_weight_level(pano_handle h, int cam, int level, const float *w, int width, int height)
{
    if (!h || cam < 0 || cam >= h->n || !w) return fail(h, "pano_set_weight_level: bad argument");
    if (h->blender != PANO_BLEND_MULTIBAND || level < 0 || level > h->nb) return fail(h, "pano_set_weight_level: bad level");
    CamTables &C = h->host.cam[cam];
    if (width != (C.rw >> level) || height != (C.rh >> level))
        return fail(h, "pano_set_weight_level: level %d must be %dx%d", level, C.rw >> level, C.rh >> level);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    if (upload2d(h, (float *)h->cam_wt[cam][level], C.wt_pitch[level], w, width, width, height)) return PANO_ERR;
    if (level == 0) C.use_wt0 = 1;
    markTiles(h, cam, level, w, width, height, width, 1.0f);
    h->tables_dirty = true;
    return PANO_OK;
}

int pano_get_weight_level(pano_handle h, int cam, int level, float *w)
{
    if (!h || cam < 0 || cam >= h->n || !w) return fail(h, "pano_get_weight_level: bad argument");
    const int top = h->blender == PANO_BLEND_MULTIBAND ? h->nb : 0;
    if (level < 0 || level > top) return fail(h, "pano_get_weight_level: bad level");
    const CamTables &C = h->host.cam[cam];
    const int lw = C.rw >> level, lh = C.rh >> level;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    if (level == 0 && !C.use_wt0) {
        std::vector<uint8_t> m((size_t)lw * lh);
        CK(h, cudaMemcpy2D(m.data(), lw, C.mask0, C.mask_pitch, lw, lh, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < m.size(); ++i) w[i] = (float)m[i] * (1.f / 255.f);
        return PANO_OK;
    }
    CK(h, cudaMemcpy2D(w, (size_t)lw * sizeof(float), C.wt[level], (size_t)C.wt_pitch[level] * sizeof(float),
                       (size_t)lw * sizeof(float), lh, cudaMemcpyDeviceToHost));
    return PANO_OK;
}

int pano_set_feather_weight(pano_handle h, int cam, const float *w, int width, int height)
{
    if (!h || cam < 0 || cam >= h->n || !w) return fail(h, "pano_set_feather_weight: bad argument");
    if (h->blender != PANO_BLEND_FEATHER) return fail(h, "pano_set_feather_weight: blender is not feather");
    CamTables &C = h->host.cam[cam];
    if (width != C.rw || height != C.rh) return fail(h, "pano_set_feather_weight: must be %dx%d", C.rw, C.rh);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    if (upload2d(h, (float *)h->cam_wt[cam][0], C.wt_pitch[0], w, width, width, height)) return PANO_ERR;
    C.use_wt0 = 1;
    h->tables_dirty = true;
    return PANO_OK;
}

int pano_set_gain_map(pano_handle h, int cam, const float *gain, int width, int height)
{
    if (!h || cam < 0 || cam >= h->n) return fail(h, "pano_set_gain_map: bad argument");
    CamTables &C = h->host.cam[cam];
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    h->tables_dirty = true;
    if (!gain) { C.gain_mode = 0; return PANO_OK; }
    const Rect &img = h->rois[cam];
    const FeedRect &fr = h->feed[cam];
    if (width != img.w || height != img.h) return fail(h, "pano_set_gain_map: must be %dx%d", img.w, img.h);
    std::vector<float> g((size_t)C.map_pitch * C.rh, 1.f);
    for (int Y = 0; Y < C.rh; ++Y) {
        const int y = reflectIdx(Y - fr.top, img.h);
        for (int X = 0; X < C.rw; ++X) g[(size_t)Y * C.map_pitch + X] = gain[(size_t)y * img.w + reflectIdx(X - fr.left, img.w)];
    }
    if (!h->cam_gain[cam]) {
        float *d = nullptr;
        if (devAlloc(h, &d, g.size(), false)) return PANO_ERR;
        h->cam_gain[cam] = d;
    }
    CK(h, cudaMemcpy(h->cam_gain[cam], g.data(), g.size() * sizeof(float), cudaMemcpyHostToDevice));
    C.gain_map = (const float *)h->cam_gain[cam];
    C.gain_mode = 1;
    return PANO_OK;
}

int pano_set_gain_scalar(pano_handle h, int cam, double gain)
{
    if (!h || cam < 0 || cam >= h->n) return fail(h, "pano_set_gain_scalar: bad argument");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    h->host.cam[cam].gain_mode = 2;
    h->host.cam[cam].gain_scalar = gain;
    h->tables_dirty = true;
    return PANO_OK;
}

int pano_attach_frontend(pano_handle h, int cam, pano_frontend_handle f)
{
    if (!h || cam < -1 || cam >= h->n) return fail(h, "pano_attach_frontend: bad argument");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    if (!f) {
        if (h->fused && pano_set_frontend_mode(h, PANO_FRONTEND_SEQUENTIAL)) return PANO_ERR;
        for (int i = 0; i < h->n; ++i) h->front[i] = nullptr;
        freeFrontScratch(h);
        if (h->has_front) dropStaging(h);         // the caller-side frame-set shrinks or grows: re-size on next use
        h->has_front = false;
        h->front_px4 = false;
        h->host.src_px = 3;
        h->tables_dirty = true;
        return PANO_OK;
    }
    if (h->fused) return fail(h, "pano_attach_frontend: switch back to PANO_FRONTEND_SEQUENTIAL before re-attaching");
    // every check comes before the handle is touched: a failing call leaves it exactly as it was
    int in_wh[2], out_wh[2];
    pano_frontend_sizes(f, in_wh, out_wh);
    if (out_wh[0] != h->cfg.src_width || out_wh[1] != h->cfg.src_height)
        return fail(h, "pano_attach_frontend: front end delivers %dx%d, stitcher expects %dx%d", out_wh[0], out_wh[1],
                    h->cfg.src_width, h->cfg.src_height);
    const size_t in_bytes = (size_t)in_wh[0] * in_wh[1] * pano_frontend_in_px(f);
    if (h->has_front && in_bytes != h->in_frame_bytes) {
        bool replaces_all = cam < 0;
        if (!replaces_all) return fail(h, "pano_attach_frontend: all cameras must share one frame size and format");
    }
    if (!h->front_out && devAlloc(h, &h->front_out, h->front_out_bytes() * h->n * h->cfg.max_batch, false)) return PANO_ERR;
    // the new assignment and its intermediate-buffer sets (one per distinct front end; a set that is still needed is kept)
    // are built on the side: running out of device memory here leaves the handle as it was
    pano_frontend_handle nf[kMaxCams] = {};
    pano_front_scratch *ns[kMaxCams] = {};
    for (int i = 0; i < h->n; ++i) nf[i] = (cam < 0 || cam == i || !h->front[i]) ? f : h->front[i];   // every camera needs one once the input format changes
    for (int i = 0; i < h->n; ++i) {
        for (int j = 0; j < i && !ns[i]; ++j)
            if (nf[j] == nf[i]) ns[i] = ns[j];
        for (int j = 0; j < h->n && !ns[i]; ++j)
            if (h->front[j] == nf[i] && h->front_scratch[j]) ns[i] = h->front_scratch[j];
        if (!ns[i]) ns[i] = pano_frontend_scratch_create(nf[i]);
        if (!ns[i]) {
            for (int k = 0; k < i; ++k) {
                bool fresh = ns[k] != nullptr;
                for (int j = 0; j < h->n; ++j) fresh = fresh && ns[k] != h->front_scratch[j];
                for (int j = 0; j < k; ++j) fresh = fresh && ns[k] != ns[j];
                if (fresh) pano_frontend_scratch_destroy(ns[k]);
            }
            return fail(h, "pano_attach_frontend: out of device memory for the front end's intermediate buffers");
        }
    }
    const size_t old_set = h->set_bytes();
    for (int i = 0; i < kMaxCams; ++i) {          // release the sets nobody uses any more, then install
        pano_front_scratch *o = h->front_scratch[i];
        if (!o) continue;
        bool kept = false;
        for (int j = 0; j < h->n; ++j) kept = kept || ns[j] == o;
        for (int j = i; j < kMaxCams; ++j)
            if (h->front_scratch[j] == o) h->front_scratch[j] = nullptr;
        if (!kept) pano_frontend_scratch_destroy(o);
    }
    for (int i = 0; i < h->n; ++i) { h->front[i] = nf[i]; h->front_scratch[i] = ns[i]; }
    h->in_frame_bytes = in_bytes;
    h->in_frame_bytes4 = (size_t)in_wh[0] * in_wh[1] * 4;
    h->has_front = true;
    // hand-over layout: one word per pixel when every camera's front end can produce it (the warp then stages its
    // source tiles by TMA), else packed BGR
    h->front_px4 = h->kc.warp_tiled;
    for (int i = 0; i < h->n; ++i) h->front_px4 = h->front_px4 && pano_frontend_can_words(h->front[i]);
    h->host.src_px = h->front_px4 ? 4 : 3;
    h->tables_dirty = true;
    dropProcessGraphs(h);
    if (h->set_bytes() != old_set) dropStaging(h);
    return PANO_OK;
}

int pano_set_frontend_mode(pano_handle h, int mode)
{
    if (!h || (mode != PANO_FRONTEND_SEQUENTIAL && mode != PANO_FRONTEND_FUSED)) return fail(h, "pano_set_frontend_mode: bad argument");
    if (mode == PANO_FRONTEND_FUSED && !h->has_front) return fail(h, "pano_set_frontend_mode: attach a front end first");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    if ((mode == PANO_FRONTEND_FUSED) == h->fused) return PANO_OK;
    const int n = h->n;
    int W = h->cfg.src_width, H = h->cfg.src_height;
    if (mode == PANO_FRONTEND_FUSED) {
        int in_wh[2], out_wh[2];
        pano_frontend_sizes(h->front[0], in_wh, out_wh);
        for (int i = 1; i < n; ++i) {
            int iw[2], ow[2];
            pano_frontend_sizes(h->front[i], iw, ow);
            if (iw[0] != in_wh[0] || iw[1] != in_wh[1]) return fail(h, "pano_set_frontend_mode: all cameras must share one frame size");
        }
        const int sw = out_wh[0], sh = out_wh[1];       // stitcher input = front-end output
        W = in_wh[0]; H = in_wh[1];
        if (h->in_frame_bytes4 != h->in_frame_bytes && !h->fused_in &&
            devAlloc(h, &h->fused_in, h->in_frame_bytes4 * n * h->cfg.max_batch, false))
            return PANO_ERR;
        h->fxmap.assign(n, {}); h->fymap.assign(n, {});
        // BORDER_REFLECT of the rotation warp, continuous form: mirror about -0.5 and n - 0.5 (pixel-centre
        // coordinates), then clamp to the pixel centres
        auto foldc = [](double c, int nn) {
            if (!(std::fabs(c) < 1e9)) return 0.0;
            double t = std::fmod(c + 0.5, 2.0 * nn);
            if (t < 0) t += 2.0 * nn;
            if (t >= nn) t = 2.0 * nn - t;
            return std::min((double)(nn - 1), std::max(0.0, t - 0.5));
        };
        for (int i = 0; i < n; ++i) {
            const size_t cnt = h->xmap[i].size();
            std::vector<double> xs(cnt), ys(cnt);
            for (size_t p = 0; p < cnt; ++p) { xs[p] = foldc(h->xmap[i][p], sw); ys[p] = foldc(h->ymap[i][p], sh); }
            pano_frontend_backmap(h->front[i], xs.data(), ys.data(), cnt);
            h->fxmap[i].resize(cnt); h->fymap[i].resize(cnt);
            for (size_t p = 0; p < cnt; ++p) {
                // the sequential path reads 0 outside the camera frame (cv::remap BORDER_CONSTANT); here the edge is
                // replicated -- only reachable when the crop rect keeps pixels whose undistort map leaves the frame
                h->fxmap[i][p] = (float)std::min((double)(W - 1), std::max(0.0, xs[p]));
                h->fymap[i][p] = (float)std::min((double)(H - 1), std::max(0.0, ys[p]));
            }
        }
    }
    h->fused = (mode == PANO_FRONTEND_FUSED);
    for (int i = 0; i < n; ++i) {
        const float *xm = h->fused ? h->fxmap[i].data() : h->xmap[i].data(), *ym = h->fused ? h->fymap[i].data() : h->ymap[i].data();
        if (buildCamMap(h, i, xm, ym, W, H)) return PANO_ERR;
    }
    if (!h->fused) { h->fxmap.clear(); h->fymap.clear(); }
    h->host.src_w = W; h->host.src_h = H; h->host.src_px = (h->fused || h->front_px4) ? 4 : 3;
    h->map64 = (32 * (W - 1) + 31 > 65535) || (32 * (H - 1) + 31 > 65535);
    h->kc.warp_tiled = (W % 16 == 0);
    h->tables_dirty = true;
    return PANO_OK;
}

int pano_process_device(pano_handle h, const uint8_t *frames_dev, uint8_t *out_dev, int batch, void *stream)
{
    if (!h || !frames_dev || !out_dev || batch < 1) return fail(h, "pano_process_device: bad argument");
    CK(h, cudaSetDevice(h->device));
    if (syncTables(h)) return PANO_ERR;
    cudaStream_t st = (cudaStream_t)stream;
    h->last_launches = 0;
    if (h->profiling) clearProf(h);
    const int S = h->cfg.max_batch;
    for (int b0 = 0; b0 < batch; b0 += S) {
        const int slots = std::min(S, batch - b0);
        if (runWave(h, frames_dev + (size_t)b0 * h->set_bytes(), out_dev + (size_t)b0 * h->out_bytes(), slots, st)) return PANO_ERR;
    }
    return PANO_OK;
}

namespace {

// host memory the CUDA driver can DMA from directly (cudaHostAlloc / cudaHostRegister / managed); anything else is pageable
bool isPinned(const void *p)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type != cudaMemoryTypeUnregistered;
}

int ensureBounce(pano_ctx *h, size_t in_bytes, size_t out_bytes)
{
    if (!h->pool) {
        static const int nthreads = getenv("PANO_HOST_THREADS") ? std::max(1, atoi(getenv("PANO_HOST_THREADS"))) : 4;
        h->pool = host_pool_create(nthreads);
    }
    if (in_bytes > h->pin_in_bytes) {
        if (h->pin_in) cudaFreeHost(h->pin_in);
        h->pin_in = nullptr; h->pin_in_bytes = 0;
        CK(h, cudaHostAlloc((void **)&h->pin_in, in_bytes, cudaHostAllocDefault));
        h->pin_in_bytes = in_bytes;
    }
    if (out_bytes > h->pin_out_bytes) {
        if (h->pin_out) cudaFreeHost(h->pin_out);
        h->pin_out = nullptr; h->pin_out_bytes = 0;
        CK(h, cudaHostAlloc((void **)&h->pin_out, out_bytes, cudaHostAllocDefault));
        h->pin_out_bytes = out_bytes;
    }
    for (auto &e : h->ev_band)
        if (!e) CK(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return PANO_OK;
}

}  // namespace

int pano_process(pano_handle h, const uint8_t *const *frames, const int *strides, uint8_t *out, int out_stride)
{
    if (!h || !frames || !out) return fail(h, "pano_process: bad argument");
    CK(h, cudaSetDevice(h->device));
    if (ensureStaging(h) || syncTables(h)) return PANO_ERR;
    int W3 = h->cfg.src_width * 3, H = h->cfg.src_height;
    size_t fbytes = h->frame_bytes();
    if (h->has_front) {
        int in_wh[2], out_wh[2];
        pano_frontend_sizes(h->front[0], in_wh, out_wh);
        W3 = in_wh[0] * pano_frontend_in_px(h->front[0]); H = in_wh[1]; fbytes = h->in_frame_bytes;
    }
    if (out_stride < h->host.cut_w * 3) return fail(h, "pano_process: out_stride too small");
    for (int i = 0; i < h->n; ++i)
        if (!frames[i] || (strides ? strides[i] : W3) < W3) return fail(h, "pano_process: bad frame %d", i);
    if (h->profiling) clearProf(h);
    // The drop-in call (one frame-set per process(), src/replay.cpp:284-292): the ~20 launches of the kernel chain are
    // captured ONCE (they only touch the handle's own staging buffers and tables) and replayed as CUDA graphs -- one per
    // camera chain plus the tail when the cameras' chains overlap the copies (canOverlapCameras), one for everything otherwise.
    static const bool no_graph = getenv("PANO_NO_GRAPH") != nullptr;
    const bool overlap = canOverlapCameras(h), graphs = !no_graph && !h->profiling;
    if (overlap && graphs && !h->graph_tail) {
        h->last_launches = 0;
        for (int i = 0; i < h->n; ++i)
            if (captureGraph(h, h->s_compute, &h->graph_cam[i], [&] { return runCameraChain(h, i, h->stage_in[0], h->s_compute); })) {
                dropProcessGraphs(h);
                return PANO_ERR;
            }
        if (captureGraph(h, h->s_compute, &h->graph_tail, [&] { return runTail(h, h->stage_in[0], h->stage_out[0], h->s_compute); })) {
            dropProcessGraphs(h);
            return PANO_ERR;
        }
        h->graph_split_launches = h->last_launches;
    }
    if (overlap) h->last_launches = graphs ? h->graph_split_launches : 0;
    cudaStream_t cs = overlap ? h->s_h2d : h->s_compute;          // the stream the input copies go to
    // camera i's copies are all enqueued on `cs`: its chain may start as soon as they have landed
    auto camera_ready = [&](int i) -> int {
        if (!overlap) return PANO_OK;
        CK(h, cudaEventRecord(h->ev_cam_in[i], cs));
        CK(h, cudaStreamWaitEvent(h->s_compute, h->ev_cam_in[i], 0));
        if (graphs) CK(h, cudaGraphLaunch(h->graph_cam[i], h->s_compute));
        else if (runCameraChain(h, i, h->stage_in[0], h->s_compute)) return PANO_ERR;
        return PANO_OK;
    };
    // Pageable buffers (what the reference's cv::Mat frames are) must not reach the driver: its internal staging copy is
    // single-threaded (~12 GB/s).  Worker threads move them through pinned bounce buffers chunk by chunk instead, each
    // chunk's DMA starting as soon as it has landed (PANO_NO_HOST_STAGING=1: the driver's path, for A/B).
    static const bool no_bounce = getenv("PANO_NO_HOST_STAGING") != nullptr;
    bool pageable_in = false;
    for (int i = 0; i < h->n && !no_bounce; ++i) pageable_in = pageable_in || !isPinned(frames[i]);
    const bool pageable_out = !no_bounce && !isPinned(out);
    const size_t out_row = (size_t)h->host.cut_w * 3;
    if ((pageable_in || pageable_out) && ensureBounce(h, fbytes * h->n, out_row * h->host.cut_h)) return PANO_ERR;
    if (pageable_in) {
        static const bool stream_in = getenv("PANO_HOST_NO_STREAM") == nullptr;      // A/B: plain memcpy into the bounce buffer
        int ticket[kMaxCams][kBounceBands];
        const int band = (H + kBounceBands - 1) / kBounceBands;
        for (int i = 0; i < h->n; ++i)
            for (int k = 0; k < kBounceBands; ++k) {
                const int r0 = std::min(H, k * band), nr = std::min(H, r0 + band) - r0;
                const int st = strides ? strides[i] : W3;
                ticket[i][k] = host_pool_copy2d(h->pool, h->pin_in + (size_t)i * fbytes + (size_t)r0 * W3, W3,
                                                frames[i] + (size_t)r0 * st, st, W3, nr, stream_in);
            }
        cudaError_t e = cudaSuccess;
        int rc = PANO_OK;
        for (int i = 0; i < h->n; ++i) {
            for (int k = 0; k < kBounceBands; ++k) {
                host_pool_wait(h->pool, ticket[i][k]);
                const int r0 = std::min(H, k * band), nr = std::min(H, r0 + band) - r0;
                const size_t off = (size_t)i * fbytes + (size_t)r0 * W3;
                if (e == cudaSuccess && rc == PANO_OK && nr > 0)
                    e = cudaMemcpyAsync(h->stage_in[0] + off, h->pin_in + off, (size_t)nr * W3, cudaMemcpyHostToDevice, cs);
            }
            if (e == cudaSuccess && rc == PANO_OK) rc = camera_ready(i);
        }
        host_pool_wait_all(h->pool);      // the workers read the caller's frames: never return while one is still running
        CK(h, e);
        if (rc) return PANO_ERR;
    } else {
        for (int i = 0; i < h->n; ++i) {
            CK(h, cudaMemcpy2DAsync(h->stage_in[0] + (size_t)i * fbytes, W3, frames[i], strides ? strides[i] : W3, W3, H,
                                    cudaMemcpyHostToDevice, cs));
            if (camera_ready(i)) return PANO_ERR;
        }
    }
    if (overlap) {
        if (graphs) CK(h, cudaGraphLaunch(h->graph_tail, h->s_compute));
        else if (runTail(h, h->stage_in[0], h->stage_out[0], h->s_compute)) return PANO_ERR;
    } else if (!graphs) {
        h->last_launches = 0;
        if (runWave(h, h->stage_in[0], h->stage_out[0], 1, h->s_compute)) return PANO_ERR;
    } else {
        if (!h->graph1) {
            h->last_launches = 0;
            if (captureGraph(h, h->s_compute, &h->graph1, [&] { return runWave(h, h->stage_in[0], h->stage_out[0], 1, h->s_compute); }))
                return PANO_ERR;
            h->graph1_launches = h->last_launches;
        }
        CK(h, cudaGraphLaunch(h->graph1, h->s_compute));
        h->last_launches = h->graph1_launches;
    }
    if (pageable_out) {
        // the panorama comes back in row bands: band k is copied out of the pinned buffer by a worker while band k + 1 is in flight
        static const bool stream_out = getenv("PANO_HOST_STREAM_OUT") != nullptr;     // A/B: streaming stores into the caller's panorama
        const int rows = h->host.cut_h, band = (rows + kBounceBands - 1) / kBounceBands;
        cudaError_t e = cudaSuccess;
        for (int k = 0; k < kBounceBands && e == cudaSuccess; ++k) {
            const int r0 = std::min(rows, k * band), nr = std::min(rows, r0 + band) - r0;
            if (nr > 0) e = cudaMemcpyAsync(h->pin_out + (size_t)r0 * out_row, h->stage_out[0] + (size_t)r0 * out_row, (size_t)nr * out_row,
                                            cudaMemcpyDeviceToHost, h->s_compute);
            if (e == cudaSuccess) e = cudaEventRecord(h->ev_band[k], h->s_compute);
        }
        for (int k = 0; k < kBounceBands && e == cudaSuccess; ++k) {
            const int r0 = std::min(rows, k * band), nr = std::min(rows, r0 + band) - r0;
            e = cudaEventSynchronize(h->ev_band[k]);
            if (e == cudaSuccess && nr > 0)
                host_pool_copy2d(h->pool, out + (size_t)r0 * out_stride, (size_t)out_stride, h->pin_out + (size_t)r0 * out_row, out_row, out_row, nr,
                                 stream_out);
        }
        host_pool_wait_all(h->pool);
        CK(h, e);
        CK(h, cudaStreamSynchronize(h->s_compute));
        return PANO_OK;
    }
    CK(h, cudaMemcpy2DAsync(out, out_stride, h->stage_out[0], out_row, out_row, h->host.cut_h, cudaMemcpyDeviceToHost, h->s_compute));
    CK(h, cudaStreamSynchronize(h->s_compute));
    return PANO_OK;
}

int pano_process_batch(pano_handle h, const uint8_t *frames_host, uint8_t *out_host, int batch)
{
    if (!h || !frames_host || !out_host || batch < 1) return fail(h, "pano_process_batch: bad argument");
    CK(h, cudaSetDevice(h->device));
    if (ensureStaging(h) || syncTables(h)) return PANO_ERR;
    // small chunks keep the H2D / compute / D2H pipeline full (fill + drain cost one chunk each);
    // the kernels have ample headroom over PCIe, so short waves do not hurt here
    static const int chunk = getenv("PANO_HOST_CHUNK") ? std::max(1, atoi(getenv("PANO_HOST_CHUNK"))) : 2;   // tuning knob
    const int S = std::min(h->cfg.max_batch, chunk);
    h->last_launches = 0;
    const bool prof = h->profiling;
    h->profiling = false;
    int wave = 0;
    for (int b0 = 0; b0 < batch; b0 += S, ++wave) {
        const int slots = std::min(S, batch - b0);
        const int q = wave % kPipeDepth;
        // staging slot q is free once its previous D2H has finished
        if (wave >= kPipeDepth) CK(h, cudaStreamWaitEvent(h->s_h2d, h->ev_out[q], 0));
        CK(h, cudaMemcpyAsync(h->stage_in[q], frames_host + (size_t)b0 * h->set_bytes(), h->set_bytes() * slots,
                              cudaMemcpyHostToDevice, h->s_h2d));
        CK(h, cudaEventRecord(h->ev_in[q], h->s_h2d));
        CK(h, cudaStreamWaitEvent(h->s_compute, h->ev_in[q], 0));
        if (wave >= kPipeDepth) CK(h, cudaStreamWaitEvent(h->s_compute, h->ev_out[q], 0));
        if (runWave(h, h->stage_in[q], h->stage_out[q], slots, h->s_compute)) { h->profiling = prof; return PANO_ERR; }
        CK(h, cudaEventRecord(h->ev_done[q], h->s_compute));
        CK(h, cudaStreamWaitEvent(h->s_d2h, h->ev_done[q], 0));
        CK(h, cudaMemcpyAsync(out_host + (size_t)b0 * h->out_bytes(), h->stage_out[q], h->out_bytes() * slots,
                              cudaMemcpyDeviceToHost, h->s_d2h));
        CK(h, cudaEventRecord(h->ev_out[q], h->s_d2h));
    }
    h->profiling = prof;
    CK(h, cudaStreamSynchronize(h->s_d2h));
    CK(h, cudaStreamSynchronize(h->s_compute));
    CK(h, cudaStreamSynchronize(h->s_h2d));
    return PANO_OK;
}

int pano_strip_set_window(pano_handle h, int x0, int x1, int margin)
{
    if (!h) return PANO_ERR;
    const int unit = 1 << h->nb;
    if (x0 < 0 || x1 > h->pad_w || x0 >= x1 || x0 % unit || x1 % unit || margin < 0)
        return fail(h, "pano_strip_set_window: [%d,%d) must be a non-empty multiple-of-%d range inside [0,%d)", x0, x1, unit, h->pad_w);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    h->strip_x0 = x0; h->strip_x1 = x1;
    const bool full = (x0 == 0 && x1 == h->pad_w);
    for (int l = 0; l <= h->nb; ++l) {
        h->host.win_lo[l] = full ? 0 : std::max(0, x0 - margin) >> l;
        h->host.win_hi[l] = full ? INT_MAX : ((std::min(h->pad_w, x1 + margin) + (1 << l) - 1) >> l);
    }
    h->tables_dirty = true;
    return PANO_OK;
}

// Hybrid decomposition: below `split_level` every rank computes its strip widened by 3 * 2^split_level level-0 columns (the
// redundant-halo rule of a split_level-band pyramid); from split_level up every rank computes the FULL width, which costs
// 4^-split_level of the work and needs the neighbours' data exactly once: the all-gather of g[split_level].
int pano_strip_set_window_hybrid(pano_handle h, int x0, int x1, int split_level)
{
    if (!h) return PANO_ERR;
    if (h->blender != PANO_BLEND_MULTIBAND || split_level < 1 || split_level > h->nb)
        return fail(h, "pano_strip_set_window_hybrid: split level must lie in [1, %d]", h->nb);
    if (pano_strip_set_window(h, x0, x1, 3 * (1 << split_level))) return PANO_ERR;
    for (int l = split_level; l <= h->nb; ++l) { h->host.win_lo[l] = 0; h->host.win_hi[l] = INT_MAX; }
    h->tables_dirty = true;
    return PANO_OK;
}

// A camera is read by a strip only where its warped ROI meets the level-0 window: the warp skips every other tile, the
// pyramid kernels every other column, and above a hybrid split level the data comes from the all-gather.
int pano_strip_cameras(pano_handle h, int *needed)
{
    if (!h || !needed) return fail(h, "pano_strip_cameras: bad argument");
    for (int i = 0; i < h->host.num_cams; ++i) {
        const auto &C = h->host.cam[i];
        needed[i] = !(C.rx + C.rw <= h->host.win_lo[0] || C.rx >= h->host.win_hi[0]);
    }
    return PANO_OK;
}

int pano_strip_phase_count(pano_handle h) { return h ? phaseCount(h) : 0; }

int pano_strip_run_phases(pano_handle h, int first, int last, const uint8_t *frames_dev, uint8_t *pano_dev, void *stream)
{
    if (!h || first < 0 || last > phaseCount(h) || first > last || !frames_dev || !pano_dev) return fail(h, "pano_strip_run_phases: bad argument");
    CK(h, cudaSetDevice(h->device));
    if (syncTables(h)) return PANO_ERR;
    cudaStream_t st = (cudaStream_t)stream;
    static const bool no_graph = getenv("PANO_NO_GRAPH") != nullptr;
    if (no_graph || h->profiling || st == nullptr || st == cudaStreamLegacy || last - first < 2) {    // the legacy stream cannot be captured
        if (first == 0) h->last_launches = 0;
        if (runPhases(h, first, last, frames_dev, pano_dev, 1, st)) return PANO_ERR;
        CK(h, cudaGetLastError());
        return PANO_OK;
    }
    // a strip's phases are dozens of small launches (a rank owns 1/N of the panorama): captured once per
    // (range, buffers), replayed as a CUDA graph
    for (auto &g : h->phase_graphs)
        if (g.first == first && g.last == last && g.frames == frames_dev && g.pano == pano_dev) {
            CK(h, cudaGraphLaunch(g.exec, st));
            h->last_launches = (first == 0 ? 0 : h->last_launches) + g.launches;
            return PANO_OK;
        }
    cudaGraph_t graph = nullptr;
    const int before = h->last_launches;
    h->last_launches = 0;
    CK(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = runPhases(h, first, last, frames_dev, pano_dev, 1, st);
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    const int captured = h->last_launches;
    h->last_launches = first == 0 ? 0 : before;
    if (rc || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        (void)cudaGetLastError();
        return rc ? PANO_ERR : fail(h, "pano_strip_run_phases: graph capture failed: %s", cudaGetErrorString(e));
    }
    pano_ctx::PhaseGraph pg{first, last, frames_dev, pano_dev, nullptr, captured};
    const cudaError_t ei = cudaGraphInstantiate(&pg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) return fail(h, "pano_strip_run_phases: graph instantiation failed: %s", cudaGetErrorString(ei));
    h->phase_graphs.push_back(pg);
    CK(h, cudaGraphLaunch(pg.exec, st));
    h->last_launches += captured;
    return PANO_OK;
}

size_t pano_strip_level_bytes(pano_handle h, int level, int ncols)
{
    if (!h || level < 0 || level > h->nb || ncols < 1) return 0;
    return halo_elems(h->host, 0, level, ncols) * sizeof(int16_t);
}

int pano_strip_level_pack(pano_handle h, int level, int col, int ncols, void *buf_dev, void *stream)
{
    if (!h || !buf_dev || level < 0 || level > h->nb || ncols < 1) return fail(h, "pano_strip_level_pack: bad argument");
    CK(h, cudaSetDevice(h->device));
    if (syncTables(h)) return PANO_ERR;
    launch_halo_copy(h->dev, h->host, 0, level, col, ncols, (int16_t *)buf_dev, false, 0, (cudaStream_t)stream);
    ++h->last_launches;
    CK(h, cudaGetLastError());
    return PANO_OK;
}

int pano_strip_level_unpack_all(pano_handle h, int level, const void *gathered_dev, int chunk_cols, const int *lo, const int *n,
                                int nranks, int self, void *stream)
{
    if (!h || !gathered_dev || !lo || !n || level < 0 || level > h->nb || nranks < 1 || nranks > 16 || self < 0 || self >= nranks)
        return fail(h, "pano_strip_level_unpack_all: bad argument");
    CK(h, cudaSetDevice(h->device));
    if (syncTables(h)) return PANO_ERR;
    launch_level_unpack_all(h->dev, h->host, level, (const int16_t *)gathered_dev, chunk_cols, lo, n, nranks, self, (cudaStream_t)stream);
    ++h->last_launches;
    CK(h, cudaGetLastError());
    return PANO_OK;
}

int pano_strip_run_phase(pano_handle h, int phase, const uint8_t *frames_dev, uint8_t *pano_dev, void *stream)
{
    if (!h || phase < 0 || phase >= phaseCount(h) || !frames_dev || !pano_dev) return fail(h, "pano_strip_run_phase: bad argument");
    CK(h, cudaSetDevice(h->device));
    if (syncTables(h)) return PANO_ERR;
    if (phase == 0) h->last_launches = 0;
    if (runPhase(h, phase, frames_dev, pano_dev, 1, (cudaStream_t)stream)) return PANO_ERR;
    CK(h, cudaGetLastError());
    return PANO_OK;
}

size_t pano_strip_halo_bytes(pano_handle h, int phase)
{
    int kind, level, ncols;
    if (!h || !phaseHalo(h, phase, kind, level, ncols)) return 0;
    return halo_elems(h->host, kind, level, ncols) * sizeof(int16_t);
}

static int haloCopy(pano_handle h, int phase, int side, void *buf, void *stream, bool unpack)
{
    int kind, level, ncols;
    if (!h || !buf || (side != 0 && side != 1) || !phaseHalo(h, phase, kind, level, ncols))
        return fail(h, "pano_strip_halo: bad argument / no halo after phase %d", phase);
    CK(h, cudaSetDevice(h->device));
    if (syncTables(h)) return PANO_ERR;
    const int lo = h->strip_x0 >> level, hi = h->strip_x1 >> level;
    int col;
    if (!unpack) col = side == 0 ? lo : hi - ncols;          // my own edge columns, for that neighbour
    else col = side == 0 ? lo - ncols : hi;                  // the neighbour's edge columns, into my halo
    launch_halo_copy(h->dev, h->host, kind, level, col, ncols, (int16_t *)buf, unpack, 0, (cudaStream_t)stream);
    ++h->last_launches;
    CK(h, cudaGetLastError());
    return PANO_OK;
}

int pano_strip_halo_pack(pano_handle h, int phase, int side, void *buf_dev, void *stream)
{
    return haloCopy(h, phase, side, buf_dev, stream, false);
}

int pano_strip_halo_unpack(pano_handle h, int phase, int side, const void *buf_dev, void *stream)
{
    return haloCopy(h, phase, side, const_cast<void *>(buf_dev), stream, true);
}

// ---- peer-memory halo exchange ----
namespace {
constexpr size_t kMailFlagBytes = 1024;          // flag words: [phase][from-side] uint32, phases <= 2 * kMaxLevels + 1

int mailLayout(pano_ctx *h)
{
    if (!h->mail_off.empty()) return PANO_OK;
    const int np = phaseCount(h);
    h->mail_off.assign((size_t)np * 4, 0);
    size_t off = kMailFlagBytes;
    for (int p = 0; p < np; ++p) {
        int kind, level, ncols;
        if (!phaseHalo(h, p, kind, level, ncols)) continue;
        const size_t bytes = (halo_elems(h->host, kind, level, ncols) * sizeof(int16_t) + 255) & ~(size_t)255;
        for (int k = 0; k < 4; ++k) { h->mail_off[(size_t)p * 4 + k] = off; off += bytes; }
    }
    h->mailbox_bytes = off;
    return PANO_OK;
}
inline uint8_t *mailSlot(pano_ctx *h, uint8_t *base, int phase, int from_side, int parity)
{
    return base + h->mail_off[(size_t)phase * 4 + from_side * 2 + parity];
}
inline uint32_t *mailFlag(uint8_t *base, int phase, int from_side) { return reinterpret_cast<uint32_t *>(base) + phase * 2 + from_side; }
}  // namespace

int pano_strip_p2p_create(pano_handle h, void *ipc_handle64, size_t *mailbox_bytes)
{
    if (!h) return PANO_ERR;
    if (h->blender != PANO_BLEND_MULTIBAND) return fail(h, "pano_strip_p2p_create: only the multiband path exchanges halos");
    CK(h, cudaSetDevice(h->device));
    if (mailLayout(h)) return PANO_ERR;
    if (!h->mailbox) {
        if (devAlloc(h, &h->mailbox, h->mailbox_bytes, true)) return PANO_ERR;
        if (devAlloc(h, &h->p2p_counters, 4, true) || devAlloc(h, &h->p2p_seq, 1, true)) return PANO_ERR;
        h->p2p_resident = halo_exchange_resident_limit(h->device);
    }
    if (ipc_handle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t hd;
        CK(h, cudaIpcGetMemHandle(&hd, h->mailbox));
        std::memcpy(ipc_handle64, &hd, sizeof hd);
    }
    if (mailbox_bytes) *mailbox_bytes = h->mailbox_bytes;
    return PANO_OK;
}

static int p2pDisconnect(pano_ctx *h, int side)
{
    // the captured frame graph bakes the peer mailbox pointers in: it dies with the mapping
    if (h->p2p_graph) { cudaGraphExecDestroy(h->p2p_graph); h->p2p_graph = nullptr; }
    if (h->peer_mail[side] && h->peer_ipc[side]) cudaIpcCloseMemHandle(h->peer_mail[side]);
    h->peer_mail[side] = nullptr; h->peer_ipc[side] = false;
    return PANO_OK;
}

int pano_strip_p2p_connect(pano_handle h, int side, const void *ipc_handle64)
{
    if (!h || (side != 0 && side != 1)) return fail(h, "pano_strip_p2p_connect: bad argument");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    p2pDisconnect(h, side);
    if (!ipc_handle64) return PANO_OK;
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, ipc_handle64, sizeof hd);
    void *p = nullptr;
    CK(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->peer_mail[side] = (uint8_t *)p; h->peer_ipc[side] = true;
    return PANO_OK;
}

int pano_strip_p2p_connect_local(pano_handle h, int side, pano_handle neighbour)
{
    if (!h || (side != 0 && side != 1)) return fail(h, "pano_strip_p2p_connect_local: bad argument");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    p2pDisconnect(h, side);
    if (!neighbour) return PANO_OK;
    if (!neighbour->mailbox) return fail(h, "pano_strip_p2p_connect_local: the neighbour has no mailbox (pano_strip_p2p_create)");
    if (neighbour->device != h->device) {
        int can = 0;
        CK(h, cudaDeviceCanAccessPeer(&can, h->device, neighbour->device));
        if (!can) return fail(h, "pano_strip_p2p_connect_local: device %d cannot address device %d", h->device, neighbour->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(neighbour->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(h, e);
        (void)cudaGetLastError();
    }
    h->peer_mail[side] = neighbour->mailbox;
    return PANO_OK;
}

int pano_strip_p2p_begin(pano_handle h, void *stream)
{
    if (!h || !h->mailbox) return fail(h, "pano_strip_p2p_begin: no mailbox (pano_strip_p2p_create)");
    CK(h, cudaSetDevice(h->device));
    launch_p2p_begin(h->p2p_seq, (cudaStream_t)stream);
    CK(h, cudaGetLastError());
    return PANO_OK;
}

// side s of this rank <-> the neighbour sees this rank on ITS side 1 - s
static void p2pSides(pano_ctx *h, int phase, int level, int ncols, HaloSide push[2], HaloSide recv[2])
{
    const int lo = h->strip_x0 >> level, hi = h->strip_x1 >> level;
    for (int side = 0; side < 2; ++side) {
        uint8_t *pm = h->peer_mail[side];
        push[side].col = side == 0 ? lo : hi - ncols;                // my own edge columns, for that neighbour
        recv[side].col = side == 0 ? lo - ncols : hi;                // the neighbour's edge columns, into my halo
        for (int par = 0; par < 2; ++par) {
            push[side].buf[par] = pm ? reinterpret_cast<int16_t *>(mailSlot(h, pm, phase, 1 - side, par)) : nullptr;
            recv[side].buf[par] = pm ? reinterpret_cast<int16_t *>(mailSlot(h, h->mailbox, phase, side, par)) : nullptr;
        }
        push[side].flag = pm ? mailFlag(pm, phase, 1 - side) : nullptr;
        recv[side].flag = pm ? mailFlag(h->mailbox, phase, side) : nullptr;
    }
}

int pano_strip_p2p_push(pano_handle h, int phase, void *stream)
{
    int kind, level, ncols;
    if (!h || !h->mailbox || !phaseHalo(h, phase, kind, level, ncols)) return fail(h, "pano_strip_p2p_push: bad argument / no halo after phase %d", phase);
    if (!h->peer_mail[0] && !h->peer_mail[1]) return PANO_OK;
    CK(h, cudaSetDevice(h->device));
    if (syncTables(h)) return PANO_ERR;
    HaloSide push[2], recv[2];
    p2pSides(h, phase, level, ncols, push, recv);
    launch_halo_push(h->dev, h->host, kind, level, ncols, push[0], push[1], h->p2p_seq, h->p2p_counters, (cudaStream_t)stream);
    ++h->last_launches;
    CK(h, cudaGetLastError());
    return PANO_OK;
}

int pano_strip_p2p_wait_unpack(pano_handle h, int phase, void *stream)
{
    int kind, level, ncols;
    if (!h || !h->mailbox || !phaseHalo(h, phase, kind, level, ncols)) return fail(h, "pano_strip_p2p_wait_unpack: bad argument / no halo after phase %d", phase);
    if (!h->peer_mail[0] && !h->peer_mail[1]) return PANO_OK;
    CK(h, cudaSetDevice(h->device));
    HaloSide push[2], recv[2];
    p2pSides(h, phase, level, ncols, push, recv);
    launch_halo_wait_unpack(h->dev, h->host, kind, level, ncols, recv[0], recv[1], h->p2p_seq, h->p2p_counters, (cudaStream_t)stream);
    ++h->last_launches;
    CK(h, cudaGetLastError());
    return PANO_OK;
}

static int p2pFrame(pano_handle h, const uint8_t *frames_dev, uint8_t *pano_dev, void *stream)
{
    if (pano_strip_p2p_begin(h, stream)) return PANO_ERR;
    h->last_launches = 1;
    const int np = phaseCount(h);
    for (int p = 0; p < np; ++p) {
        if (runPhase(h, p, frames_dev, pano_dev, 1, (cudaStream_t)stream)) return PANO_ERR;
        int kind, level, ncols;
        if (!phaseHalo(h, p, kind, level, ncols)) continue;
        if (!h->peer_mail[0] && !h->peer_mail[1]) continue;
        // one launch per exchange while all its blocks are certainly co-resident (the runtime's occupancy figure for the
        // kernel x SMs, minus a quarter as margin for other work on the device; every spin is bounded and raises the
        // handle's error flag -- pano_strip_p2p_check), otherwise push and wait/unpack as two launches
        static const bool split = getenv("PANO_P2P_SPLIT") != nullptr;
        HaloSide push[2], recv[2];
        p2pSides(h, p, level, ncols, push, recv);
        if (!split && launch_halo_exchange(h->dev, h->host, kind, level, ncols, push, recv, h->p2p_seq, h->p2p_counters,
                                           h->p2p_resident, (cudaStream_t)stream)) {
            ++h->last_launches;
            continue;
        }
        if (pano_strip_p2p_push(h, p, stream) || pano_strip_p2p_wait_unpack(h, p, stream)) return PANO_ERR;
    }
    return PANO_OK;
}

static bool p2pUseGraph(pano_handle h, cudaStream_t st)
{
    static const bool no_graph = getenv("PANO_P2P_NO_GRAPH") != nullptr;
    return !(no_graph || h->profiling || st == nullptr || st == cudaStreamLegacy);     // the legacy stream cannot be captured
}

// The frame's ~60 small launches are identical from frame to frame (the sequence number lives in device memory):
// they are captured once per (frames, panorama) buffer pair and replayed as a graph.
int pano_strip_p2p_prepare(pano_handle h, const uint8_t *frames_dev, uint8_t *pano_dev, void *stream)
{
    if (!h || !frames_dev || !pano_dev || !h->mailbox) return fail(h, "pano_strip_p2p_prepare: bad argument / no mailbox");
    CK(h, cudaSetDevice(h->device));
    if (h->tables_dirty && h->p2p_graph) { cudaGraphExecDestroy(h->p2p_graph); h->p2p_graph = nullptr; }
    if (syncTables(h)) return PANO_ERR;
    cudaStream_t st = (cudaStream_t)stream;
    if (!p2pUseGraph(h, st)) return PANO_OK;
    if (h->p2p_graph && (h->p2p_graph_frames != frames_dev || h->p2p_graph_pano != pano_dev)) {
        cudaGraphExecDestroy(h->p2p_graph);
        h->p2p_graph = nullptr;
    }
    if (!h->p2p_graph) {
        cudaGraph_t g = nullptr;
        CK(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = p2pFrame(h, frames_dev, pano_dev, stream);
        const cudaError_t e = cudaStreamEndCapture(st, &g);
        if (rc || e != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            (void)cudaGetLastError();
            return rc ? PANO_ERR : fail(h, "pano_strip_run_p2p: graph capture failed: %s", cudaGetErrorString(e));
        }
        const cudaError_t ei = cudaGraphInstantiate(&h->p2p_graph, g, 0);
        cudaGraphDestroy(g);
        if (ei != cudaSuccess) { h->p2p_graph = nullptr; return fail(h, "pano_strip_run_p2p: graph instantiation failed: %s", cudaGetErrorString(ei)); }
        h->p2p_graph_frames = frames_dev; h->p2p_graph_pano = pano_dev;
        CK(h, cudaGraphUpload(h->p2p_graph, st));
    }
    return PANO_OK;
}

int pano_strip_run_p2p(pano_handle h, const uint8_t *frames_dev, uint8_t *pano_dev, void *stream)
{
    if (pano_strip_p2p_prepare(h, frames_dev, pano_dev, stream)) return PANO_ERR;
    cudaStream_t st = (cudaStream_t)stream;
    if (!p2pUseGraph(h, st)) {
        if (p2pFrame(h, frames_dev, pano_dev, stream)) return PANO_ERR;
        CK(h, cudaGetLastError());
        return PANO_OK;
    }
    CK(h, cudaGraphLaunch(h->p2p_graph, st));
    return PANO_OK;
}

int pano_strip_p2p_check(pano_handle h)
{
    if (!h || !h->p2p_counters) return fail(h, "pano_strip_p2p_check: no mailbox (pano_strip_p2p_create)");
    CK(h, cudaSetDevice(h->device));
    unsigned flag = 0;
    CK(h, cudaMemcpy(&flag, h->p2p_counters + 2, sizeof flag, cudaMemcpyDeviceToHost));
    if (flag) {
        CK(h, cudaMemset(h->p2p_counters, 0, 4 * sizeof(unsigned)));
        return fail(h, "pano_strip_p2p: a halo wait ran out of patience (a neighbour never published its columns)");
    }
    return PANO_OK;
}

int pano_profile_enable(pano_handle h, int on)
{
    if (!h) return PANO_ERR;
    h->profiling = on != 0;
    if (!on) clearProf(h);
    return PANO_OK;
}

int pano_profile_read(pano_handle h, int max, const char **names, float *ms, int *launches, double *alg_bytes)
{
    if (!h) return PANO_ERR;
    // aggregate by name, preserving first-seen order
    std::vector<const char *> order;
    std::vector<float> tms;
    std::vector<int> cnt;
    std::vector<double> bytes;
    for (auto &p : h->prof) {
        if (cudaEventSynchronize(p.e1) != cudaSuccess) return fail(h, "profile event sync failed");
        float t = 0.f;
        cudaEventElapsedTime(&t, p.e0, p.e1);
        size_t k = 0;
        for (; k < order.size(); ++k)
            if (!std::strcmp(order[k], p.name)) break;
        if (k == order.size()) { order.push_back(p.name); tms.push_back(0.f); cnt.push_back(0); bytes.push_back(0.0); }
        tms[k] += t; cnt[k] += 1; bytes[k] += p.bytes;
    }
    const int nout = std::min<int>(max, (int)order.size());
    for (int k = 0; k < nout; ++k) {
        if (names) names[k] = order[k];
        if (ms) ms[k] = tms[k];
        if (launches) launches[k] = cnt[k];
        if (alg_bytes) alg_bytes[k] = bytes[k];
    }
    return nout;
}

int pano_last_launch_count(pano_handle h) { return h ? h->last_launches : 0; }

}  // extern "C"
