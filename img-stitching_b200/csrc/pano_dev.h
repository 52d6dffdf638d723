// Device-side table layout shared by the kernels (kernels.cu) and the C ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace pano {

constexpr int kMaxCams = 8;
constexpr int kMaxLevels = 10;  // num_bands + 1 <= kMaxLevels

// One camera's static tables and per-wave pyramid workspace.  All "pitch" values are in
// ELEMENTS of the array they describe.  Gaussian pyramid planes are channel-planar UINT8
// (every Gaussian level of an 8-bit frame is provably in [0, 255]: level 0 is the warped image after
// the saturating gain, cv::pyrDown is a convex combination with (S + 128) >> 8), rows 128-byte aligned:
// g[l] -> [slot][3 planes][h_l rows][g_pitch[l]].  Only the collapsed dst pyramid out[l] needs 16 bits.
struct CamTables {
    // feed rect of MultiBandBlender::feed at level 0, in padded-dst coordinates
    // (for feather / no-blend: the image rect itself, borders = 0)
    int rx, ry, rw, rh;
    // folded fixed-point remap table over the feed rect: (sy' << 16) | sx'   (map32)
    // or {sx', sy'} pairs (map64) when 32*src_dim does not fit 16 bits
    const uint32_t *map32;
    const uint2 *map64;
    int map_pitch;
    // per 128x16 output tile: source bounding box of the gather {first column (multiple of 16),
    // first row, rows, 16-pixel groups per row}; rows == 0 -> tile gathers straight from global
    const int4 *tiles;
    int tiles_x, tiles_y;
    // exposure gain: mode 0 off, 1 per-pixel float map (feed-rect layout, pitch = map_pitch),
    // 2 scalar (double, GainCompensator)
    int gain_mode;
    const float *gain_map;
    double gain_scalar;
    // weights: level 0 = 8-bit mask (weight = mask * (1/255.f)), levels >= 1 float
    const uint8_t *mask0;
    int mask_pitch;
    const float *wt[kMaxLevels];
    int wt_pitch[kMaxLevels];
    int use_wt0;  // 1: level-0 weights come from wt[0] (caller override / feather) instead of mask0
    // pyramid workspace
    uint8_t *g[kMaxLevels];
    int g_pitch[kMaxLevels];
    size_t g_plane[kMaxLevels];  // elements per plane
    size_t g_slot[kMaxLevels];   // elements per frame-set slot (= 3 planes)
};

struct PanoTables {
    int num_cams;
    int src_w, src_h;
    int src_px;              // bytes per source pixel: 3 = interleaved BGR; 4 = 8UC4 camera frames (fused front end)
    int nb;                  // effective band count (levels 0..nb)
    int pad_w, pad_h;        // padded dst size (level 0)
    int roi_w, roi_h;        // unpadded dst roi size
    int cut_x, cut_y, cut_w, cut_h;
    // column window (spatial strip split): blocks whose dst columns at level l fall entirely
    // outside [win_lo[l], win_hi[l]) are skipped.  Full range = no split.
    int win_lo[kMaxLevels], win_hi[kMaxLevels];
    // Collapse work lists, per level, over kWalkTileW x kWalkTileH tiles of the padded dst (static; rebuilt whenever
    // masks / weights change).  Entry = tx | ty << 12 | info << 24.
    //   walk_list: tiles whose blend needs no weights -- info = 1 + cam (exactly that camera has weight there and
    //              all of its weights are exactly 1.0f) or kWalkEmpty (no camera has weight) -> collapse_walk_kernel
    //   gen_list:  every other tile that is needed -- info = bitmask of the cameras with any non-zero weight inside
    //              the tile (the others contribute exactly nothing and are never touched) -> collapse8_kernel
    // Level-0 tiles outside the cut rows are in neither list.
    const uint32_t *walk_list[kMaxLevels], *gen_list[kMaxLevels];
    int walk_n[kMaxLevels], gen_n[kMaxLevels];
    int unit_norm_exact;     // host verified (short)(a / (1.0f + 1e-5f)) == a - sign(a) for all int16 a
    // collapsed dst pyramid, levels 1..nb: [slot][3][h_l][out_pitch[l]]
    int16_t *outp[kMaxLevels];
    int out_pitch[kMaxLevels];
    size_t out_plane[kMaxLevels];
    size_t out_slot[kMaxLevels];
    CamTables cam[kMaxCams];
};

// launchers (kernels.cu) -- all asynchronous on `stream`
constexpr int kWarpTileW = 128, kWarpTileH = 16;      // output pixels per warp-kernel block
constexpr int kWarpSmemWords = 7168;                  // staged source footprint, one 32-bit word per pixel (28 KB)

// walker tile: kWalkLanesX lanes of 8 pixels across, 32 / kWalkLanesX bands of 2R fine rows down (one warp per plane)
constexpr int kWalkTileW = 64, kWalkR = 4, kWalkLanesX = kWalkTileW / 8, kWalkTileH = (32 / kWalkLanesX) * 2 * kWalkR;
constexpr int kWalkEmpty = 0xff;

struct KernelChoice {
    bool warp_tiled = false;                 // staged-gather warp kernel usable (row bytes % 16 == 0)
    bool pyrdown8[kMaxLevels] = {};          // packed 8-wide pyrDown usable at this source level
    bool collapse8[kMaxLevels] = {};         // packed 8x2 collapse usable at this level
};

// cameras [first, first + count) of every slot; count == 0: all of them.  The per-camera stages (warp, pyrDown levels) can be
// launched camera by camera, so that a camera's chain starts as soon as ITS frame is in device memory (pano_process)
struct CamRange { int first = 0, count = 0; };
void launch_warp(const PanoTables *dev, const PanoTables &host, const KernelChoice &kc, const uint8_t *frames, int nslots,
                 cudaStream_t stream, CamRange cams = CamRange());
void launch_pyrdown(const PanoTables *dev, const PanoTables &host, const KernelChoice &kc, int level, int nslots,
                    cudaStream_t stream, CamRange cams = CamRange());
void launch_coarsest(const PanoTables *dev, const PanoTables &host, uint8_t *pano, int nslots, cudaStream_t stream);
// returns the number of kernels launched
// side: a second stream + two events of the caller.  With one or two frame-set slots in flight the two kernels of a level
// (walker tiles, seam tiles -- disjoint outputs, same inputs) are a few microseconds each and mostly launch + drain: they
// then run side by side (fork / join through the events; capturable).  nullptr, or more slots: one after the other.
struct SideStream { cudaStream_t st; cudaEvent_t fork, join; };
int launch_collapse(const PanoTables *dev, const PanoTables &host, const KernelChoice &kc, int level, uint8_t *pano,
                    int nslots, cudaStream_t stream, const SideStream *side = nullptr);
// feather / no-blend as two passes: launch_warp (staged gather -> g[0]) + this streaming blend over the warped images
void launch_blend_g0(const PanoTables *dev, const PanoTables &host, int blender, uint8_t *pano, int nslots, cudaStream_t stream);
void launch_direct_blend(const PanoTables *dev, const PanoTables &host, int blender, const uint8_t *frames,
                         uint8_t *pano, int nslots, cudaStream_t stream);

// init-time tables built on the device (mask refresh): one level of the float weight pyramid (cv::pyrDown on CV_32F,
// same evaluation order as pano::pyrDownF32; from_mask: src is the 8-bit level-0 mask, weight = mask * (1/255.f)),
// and the per-walker-tile statistics of one camera's weight level (any non-zero / count of exact ones)
void launch_weight_pyrdown(const void *src, bool from_mask, int spitch, int sw, int sh, float *dst, int dpitch,
                           cudaStream_t stream);
// m_blenderMask from the seam finder's low-resolution mask (dilate -> INTER_LINEAR_EXACT -> AND full mask), tight w x h
void launch_seam_mask(const uint8_t *seam, int sw, int sh, int spitch, const int *xo, const int *xc, const int *yo, const int *yc,
                      const uint8_t *full, uint8_t *dst, int w, int h, cudaStream_t stream);
// FeatherBlender weight map of one camera: min(distanceTransform(mask, DIST_L1, 3) * sharpness, 1); tmp = w * h ints
void launch_feather_weight(const uint8_t *mask, int mpitch, int w, int h, float sharpness, int *tmp, float *out, int opitch,
                           cudaStream_t stream);
void launch_tile_stats(const void *data, bool is_mask, int pitch, int w, int h, int ox, int oy, int tiles_x, int tiles_y,
                       uint8_t *nz, int *ones, cudaStream_t stream);

// strip-split halo columns: pack / unpack `ncols` dst columns starting at dst column `col`
// (level `level`) of either every camera's g[level] (kind 0) or out[level] (kind 1) into / from a
// dense buffer laid out [cam][plane][row][ncols] int16 (8-bit Gaussian samples are widened; cameras that do not
// cover the column contribute zeros and ignore the incoming data).
void launch_halo_copy(const PanoTables *dev, const PanoTables &host, int kind, int level, int col, int ncols,
                      int16_t *buf, bool unpack, int slot, cudaStream_t stream);
size_t halo_elems(const PanoTables &host, int kind, int level, int ncols);
// hybrid strip split: scatter the all-gathered level-`level` chunks of the other ranks (<= 16) into g[level]
void launch_level_unpack_all(const PanoTables *dev, const PanoTables &host, int level, const int16_t *buf, int chunk_cols,
                             const int *lo, const int *n, int nranks, int self, cudaStream_t stream);

// Peer-memory halo exchange (no collective library on the data path): one side of one exchange.
//   push:        col = first of this rank's own edge columns, buf = slot in the NEIGHBOUR's mailbox, flag = word in the
//                neighbour's flag array;  buf == nullptr: no neighbour on that side
//   wait/unpack: col = first halo column to fill, buf / flag = this rank's own mailbox slot / flag word
// The frame's sequence number lives in DEVICE memory (*seq, bumped by launch_p2p_begin) and selects the slot parity
// (buf[seq & 1]), so that a whole frame's launch sequence is identical from frame to frame and can be replayed as a CUDA graph.
struct HaloSide { int col; int16_t *buf[2]; uint32_t *flag; };
void launch_p2p_begin(uint32_t *seq, cudaStream_t stream);
void launch_halo_push(const PanoTables *dev, const PanoTables &host, int kind, int level, int ncols, const HaloSide &left,
                      const HaloSide &right, const uint32_t *seq, unsigned *counters, cudaStream_t stream);
// both halves in one launch; false (nothing launched) when the grid would exceed max_resident_blocks
// (halo_exchange_resident_limit: the runtime's occupancy for the kernel x SMs, minus a margin)
int halo_exchange_resident_limit(int device);
bool launch_halo_exchange(const PanoTables *dev, const PanoTables &host, int kind, int level, int ncols, const HaloSide push[2],
                          const HaloSide recv[2], const uint32_t *seq, unsigned *counters, int max_resident_blocks, cudaStream_t stream);
// counters: [0], [1] block-completion counters of the two sides, [2] error flag raised by a spin that ran out
void launch_halo_wait_unpack(const PanoTables *dev, const PanoTables &host, int kind, int level, int ncols, const HaloSide &left,
                             const HaloSide &right, const uint32_t *seq, unsigned *counters, cudaStream_t stream);

}  // namespace pano
