// nvCam front end (include/nvcam.hpp:898-929 + getFrame :1092-1094) on sm_100a:
//   resize(8UC4, bilinear) -> drop alpha -> remap INTER_CUBIC (border 0) -> crop rect
//   -> resize(bilinear) -> resize to the stitcher input size.
// Identity resizes are skipped (cv::resize with equal sizes is a copy); dropping the alpha
// channel is folded into whichever kernel reads the camera frame.  Arithmetic follows
// SURVEY.md A4 (15-bit bicubic table with OpenCV's tap-adjust rule) and A5 (11-bit bilinear
// resize coefficients, rows clipped / columns clamped) exactly.
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
#include "pdl.h"
#include "tma.h"

using namespace pano;

namespace {

struct ResizeTab {
    int sw = 0, sh = 0, dw = 0, dh = 0;
    int *xofs = nullptr, *yofs = nullptr;
    short2 *xa = nullptr, *ya = nullptr;   // (a0, a1)
    int *ybeg = nullptr;                   // [sh + 2]: first output row y with yofs[y] >= v, for v = -1 .. sh
    bool walk = false;                     // rows are up-scaled with non-negative coefficients: resize4_walk_kernel applies
};

__device__ __forceinline__ int sat_u8(int v) { return max(0, min(255, v)); }

// d = c + a.lo16 * b.byte0 + a.hi16 * b.byte1  (signed 16-bit weights x unsigned 8-bit pixels)
__device__ __forceinline__ int dp2a_su(int a, unsigned b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// d = c + a.lo16 * b.byte2 + a.hi16 * b.byte3: one PRMT (b0 b1 g0 g1) feeds the B and the G accumulator
__device__ __forceinline__ int dp2a_su_hi(int a, unsigned b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// cv::resize INTER_LINEAR u8: one thread = one output pixel (3 channels out, CIN in)
template <int CIN>
__global__ void __launch_bounds__(256) resize_kernel(const uint8_t *__restrict__ src, size_t src_img, int sstride,
                                                     uint8_t *__restrict__ dst, size_t dst_img, int dstride,
                                                     ResizeTab t)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= t.dw || y >= t.dh) return;
    const uint8_t *s = src + (size_t)blockIdx.z * src_img;
    uint8_t *d = dst + (size_t)blockIdx.z * dst_img + (size_t)y * dstride + (size_t)x * 3;
    const int sx0 = t.xofs[x], sx1 = min(sx0 + 1, t.sw - 1);
    const int yo = t.yofs[y];
    const int sy0 = min(max(yo, 0), t.sh - 1), sy1 = min(max(yo + 1, 0), t.sh - 1);
    const short2 ax = t.xa[x], ay = t.ya[y];
    const uint8_t *r0 = s + (size_t)sy0 * sstride, *r1 = s + (size_t)sy1 * sstride;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int t0 = __ldg(r0 + sx0 * CIN + c) * ax.x + __ldg(r0 + sx1 * CIN + c) * ax.y;
        const int t1 = __ldg(r1 + sx0 * CIN + c) * ax.x + __ldg(r1 + sx1 * CIN + c) * ax.y;
        d[c] = (uint8_t)sat_u8((((ay.x * (t0 >> 4)) >> 16) + ((ay.y * (t1 >> 4)) >> 16) + 2) >> 2);
    }
}

// drop alpha only (camera size == undistort size and no undistort): 4 -> 3 channels
__global__ void __launch_bounds__(256) drop_alpha_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                         size_t npx)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npx) return;
    const uchar4 p = __ldg(reinterpret_cast<const uchar4 *>(src) + i);
    dst[3 * i] = p.x; dst[3 * i + 1] = p.y; dst[3 * i + 2] = p.z;
}

// cv::cvtColor(COLOR_YUV2BGRA_YUYV) (the YUYVCAM ingest, include/nvcam.hpp:880-886): ITU-R BT.601 in 20-bit fixed
// point, alpha 255.  Streaming kernel: one thread = 16 bytes in (8 pixels: Y0 U Y1 V ...) -> two 16-byte stores.
__device__ __forceinline__ uint32_t yuv_px(int y, int ruv, int guv, int buv)
{
    const int yy = max(0, y - 16) * 1220542;
    uint32_t hi, px;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(255), "r"((yy + ruv) >> 20));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(px) : "r"((yy + guv) >> 20), "r"((yy + buv) >> 20), "r"(hi));
    return px;
}
__device__ __forceinline__ uint2 yuyv_pair(uint32_t w)      // bytes Y0 U Y1 V -> two BGRA words
{
    const int u = (int)((w >> 8) & 255u) - 128, v = (int)(w >> 24) - 128;
    const int ruv = (1 << 19) + 1673527 * v, guv = (1 << 19) - 852492 * v - 409993 * u, buv = (1 << 19) + 2116026 * u;
    return make_uint2(yuv_px((int)(w & 255u), ruv, guv, buv), yuv_px((int)((w >> 16) & 255u), ruv, guv, buv));
}
__global__ void __launch_bounds__(256) yuyv_to_bgra_kernel(const uint4 *__restrict__ src, size_t src_img16, uint4 *__restrict__ dst,
                                                           size_t dst_img16, size_t n16)
{
    pdl_enter();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    const uint4 s = __ldg(src + (size_t)blockIdx.y * src_img16 + i);
    const uint2 a = yuyv_pair(s.x), b = yuyv_pair(s.y), c = yuyv_pair(s.z), d = yuyv_pair(s.w);
    uint4 *o = dst + (size_t)blockIdx.y * dst_img16 + 2 * i;
    o[0] = make_uint4(a.x, a.y, b.x, b.y);
    o[1] = make_uint4(c.x, c.y, d.x, d.y);
}
// any size / alignment: one thread = one pixel pair
__global__ void __launch_bounds__(256) yuyv_to_bgra_pair_kernel(const uint8_t *__restrict__ src, size_t src_img, uint8_t *__restrict__ dst,
                                                                size_t dst_img, size_t npairs)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const uint8_t *s = src + (size_t)blockIdx.y * src_img + 4 * i;
    const uint2 p = yuyv_pair((uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24));
    uint8_t *o = dst + (size_t)blockIdx.y * dst_img + 8 * i;
    for (int k = 0; k < 4; ++k) { o[k] = (uint8_t)(p.x >> (8 * k)); o[4 + k] = (uint8_t)(p.y >> (8 * k)); }
}

// cv::remap INTER_CUBIC, BORDER_CONSTANT(0), restricted to the crop rect.
// map: per undistorted pixel {sx, sy} = cvRound(32*map) with the integer part saturated to int16.
template <int CIN>
__global__ void __launch_bounds__(256) cubic_kernel(const uint8_t *__restrict__ src, size_t src_img, int sw, int sh,
                                                    int sstride, const int2 *__restrict__ map, int map_w,
                                                    const short *__restrict__ tab, int rx, int ry, int rw, int rh,
                                                    uint8_t *__restrict__ dst, size_t dst_img)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= rw || y >= rh) return;
    const int2 m = __ldg(map + (size_t)(y + ry) * map_w + (x + rx));
    const int ix = (m.x >> 5) - 1, iy = (m.y >> 5) - 1;
    const short *w = tab + (((m.y & 31) << 5) | (m.x & 31)) * 16;
    const uint8_t *s = src + (size_t)blockIdx.z * src_img;
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int py = iy + r;
        if ((unsigned)py >= (unsigned)sh) continue;
        const uint8_t *row = s + (size_t)py * sstride;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int px = ix + q;
            if ((unsigned)px >= (unsigned)sw) continue;
            const int wt = __ldg(w + r * 4 + q);
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] += wt * __ldg(row + px * CIN + c);
        }
    }
    uint8_t *d = dst + (size_t)blockIdx.z * dst_img + ((size_t)y * rw + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = (uint8_t)sat_u8((acc[c] + 16384) >> 15);
}


// ---------------------------------------------------------------- fast path (8UC4 camera frames)
// The camera frame already has one 32-bit word per pixel, so every tap is one aligned word load
// and a warp's 32 consecutive output pixels read a compact, coalesced window.  The cropped
// undistorted image is kept in the same word-per-pixel form (an internal buffer), which makes the
// bilinear resize that follows a 4-word gather as well.  Map entries / table offsets of all the
// pixels a thread owns are requested before any of them is used (the kernels are otherwise bound
// by one dependent load per pixel).

// cv::remap INTER_CUBIC over the crop rect; src: 8UC4 words, dst: words (B | G<<8 | R<<16).
// Map entry, packed form (sources up to 2043 pixels per side): (tx+4) | (ty+4) << 11 | fx << 22 | fy << 27
// where (tx, ty) is the first of the 4x4 taps, clamped to [-4, size] (beyond that every tap is
// outside the image and contributes 0 either way).
template <bool kPacked, bool kTexW>
__global__ void __launch_bounds__(256) cubic4_kernel(const uint32_t *__restrict__ src, size_t src_img_words, int sw, int sh,
                                                     const void *__restrict__ map_, int map_w, const uint4 *__restrict__ tab,
                                                     cudaTextureObject_t wtex, int rx, int ry, int rw, int rh,
                                                     uint32_t *__restrict__ dst, size_t dst_img_words)
{
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (y >= rh) return;
    const int xb = blockIdx.x * 128 + threadIdx.x;
    const uint32_t *s = src + (size_t)blockIdx.z * src_img_words;
    uint32_t *d = dst + (size_t)blockIdx.z * dst_img_words + (size_t)y * rw;
    int2 m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = xb + 32 * j;
        m[j] = make_int2(0, 0);
        if (x < rw) {
            const size_t mi = (size_t)(y + ry) * map_w + (x + rx);
            if (kPacked) m[j].x = (int)__ldg(static_cast<const uint32_t *>(map_) + mi);
            else m[j] = __ldg(static_cast<const int2 *>(map_) + mi);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = xb + 32 * j;
        if (x >= rw) continue;
        int ix, iy, fidx;
        if (kPacked) {
            const uint32_t e = (uint32_t)m[j].x;
            ix = (int)(e & 2047u) - 4; iy = (int)((e >> 11) & 2047u) - 4;
            fidx = (int)(((e >> 27) << 5) | ((e >> 22) & 31u));
        } else {
            ix = (m[j].x >> 5) - 1; iy = (m[j].y >> 5) - 1;
            fidx = ((m[j].y & 31) << 5) | (m[j].x & 31);
        }
        uint4 wa, wb;                            // 16 shorts = 2 x uint4
        if (kTexW) {
            // the weights travel through the texture pipe of L1TEX, the taps through the LSU pipe
            wa = tex1Dfetch<uint4>(wtex, fidx << 1); wb = tex1Dfetch<uint4>(wtex, (fidx << 1) + 1);
        } else {
            const uint4 *wp = tab + (fidx << 1);
            wa = __ldg(wp); wb = __ldg(wp + 1);
        }
        const uint32_t wpk[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        uint32_t t[16];
        if (ix >= 0 && iy >= 0 && ix + 3 < sw && iy + 3 < sh) {
            const uint32_t *p = s + (size_t)iy * sw + ix;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) t[4 * r + q] = __ldg(p + r * sw + q);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int px = ix + q, py = iy + r;
                    t[4 * r + q] = ((unsigned)px < (unsigned)sw && (unsigned)py < (unsigned)sh) ? __ldg(s + (size_t)py * sw + px) : 0u;
                }
        }
        // two taps per IDP.2A: the table already stores neighbouring weights as packed int16 pairs;
        // PRMT pairs up the same channel of two neighbouring pixels
        int a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t bg = __byte_perm(t[2 * k], t[2 * k + 1], 0x5140);
            a0 = dp2a_su((int)wpk[k], bg, a0);
            a1 = dp2a_su_hi((int)wpk[k], bg, a1);
            a2 = dp2a_su((int)wpk[k], __byte_perm(t[2 * k], t[2 * k + 1], 0x0062), a2);
        }
        d[x] = (uint32_t)sat_u8((a0 + 16384) >> 15) | ((uint32_t)sat_u8((a1 + 16384) >> 15) << 8) |
               ((uint32_t)sat_u8((a2 + 16384) >> 15) << 16);
    }
}

// Tiled form of the same remap: a block owns a 64 x 16 tile of the crop rect (thread = 2 x 2 pixels, 32 columns /
// 8 rows apart).
//  * The tile's source footprint (static bounding box of all its 4x4 tap windows, from the init-time tile
//    table) is staged in shared memory with coalesced 16-byte loads; positions outside the camera frame are
//    staged as 0 (BORDER_CONSTANT), so the gather has no border logic.  The staged row pitch is a multiple
//    of 32 words: a tap's bank depends on its column only, and the 32 consecutive pixels of a warp read
//    nearly consecutive columns.  Measured (ncu): 1.8 wavefronts per LDS.32 -- 32 output pixels cover 34-37 source
//    columns and bend across source rows, so two lanes always meet in a bank -- still cheaper than the global-load
//    form, whose taps straddle cache lines.
//  * The 32-byte weight entry of every pixel comes through the TEXTURE pipe of L1TEX (two 16-byte fetches of
//    a linear texture over the 1024-entry table), which runs beside the LSU pipe that serves the taps.
//  * Rounding is folded into the accumulator start value; saturate + pack is two cvt.pack instructions.
//  * On the TMA path (default) the footprint is not staged by the threads at all: one elected thread issues a single
//    cp.async.bulk.tensor box load (kCubBoxW x 24 or 32 words at the tile's source origin) that the copy engine lands
//    in shared memory while the block fetches its map entries; out-of-bounds parts of the box are ZERO-FILLED by the
//    hardware, which is exactly BORDER_CONSTANT(0).  The LDG/STS staging loop remains as the fallback for geometries
//    whose footprints exceed the box.
constexpr int kCubTileW = 64, kCubTileH = 16, kCubSmemWords = 3072;    // 12 KB
constexpr int kCubBoxW = 96, kCubBoxH0 = 24, kCubBoxH1 = 32;           // TMA boxes (words x rows); 96 * 32 = kCubSmemWords

struct CubicArgs {
    CUtensorMap tm0, tm1;            // [images][sh][sw] words, boxes kCubBoxW x kCubBoxH0 / kCubBoxH1 (TMA path only)
    const uint32_t *src; size_t src_img_words; int sw, sh;
    const uint32_t *map; int map_w; cudaTextureObject_t wtex; const int4 *tiles;
    int rx, ry, rw, rh;
    uint32_t *dst; size_t dst_img_words;
    int tx, ty;                      // tiles per image row / column (logical grid = tx x ty x images)
};

// one 256-thread block = tile (bx, by) of image bz; sm = kCubSmemWords words of shared memory
template <bool kTma>
__device__ __forceinline__ void cubic5_body(const CubicArgs &A, uint32_t *__restrict__ sm, uint64_t *bar, int bx, int by, int bz)
{
    const uint32_t *__restrict__ src = A.src, *__restrict__ map = A.map;
    const int sw = A.sw, sh = A.sh, map_w = A.map_w, rx = A.rx, ry = A.ry, rw = A.rw, rh = A.rh;
    const cudaTextureObject_t wtex = A.wtex;
    uint32_t *__restrict__ dst = A.dst;
    const int lane = threadIdx.x, wy = threadIdx.y, tid = wy * 32 + lane;
    const int4 td = __ldg(A.tiles + by * A.tx + bx);   // {x0 (%4 == 0), y0, rows | chunks per row << 16, 2^16 / chunks}
    const int rows = td.z & 0xffff, W4 = td.z >> 16;
    const int P = kTma ? kCubBoxW : ((W4 * 4 + 31) & ~31);              // staged row pitch, words
    if (kTma && tid == 0) {
        mbar_init(bar, 1);
        const bool small = rows <= kCubBoxH0;
        mbar_expect_tx(bar, (unsigned)(kCubBoxW * (small ? kCubBoxH0 : kCubBoxH1) * 4));
        tma_load_3d(sm, small ? &A.tm0 : &A.tm1, td.x, td.y, bz, bar);
    }
    const int y = by * kCubTileH + wy, xb = bx * kCubTileW + lane;
    const uint32_t *s = src + (size_t)bz * A.src_img_words;
    uint32_t m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int xx = xb + 32 * (j & 1), yy = y + 8 * (j >> 1);
        m[j] = (xx < rw && yy < rh) ? __ldg(map + (size_t)(yy + ry) * map_w + rx + xx) : 0u;
    }
    if (!kTma) {
        for (int c = tid; c < rows * W4; c += 256) {
            const int r = (c * td.w) >> 16, q = c - r * W4;
            const int sy = td.y + r, sx = td.x + 4 * q;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if ((unsigned)sy < (unsigned)sh && (unsigned)sx < (unsigned)sw) v = __ldg(reinterpret_cast<const uint4 *>(s + (size_t)sy * sw + sx));
            *reinterpret_cast<uint4 *>(sm + r * P + 4 * q) = v;
        }
    }
    __syncthreads();                                                    // staging done / the mbarrier is initialised
    if (kTma) mbar_wait(bar, 0);                                        // the box has landed
    uint32_t *d = dst + (size_t)bz * A.dst_img_words + (size_t)y * rw + xb;
    const int sbase = -((td.y + 4) * P + td.x + 4);                     // the map stores tx + 4, ty + 4
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (xb + 32 * (j & 1) >= rw || y + 8 * (j >> 1) >= rh) continue;
        const uint32_t e = m[j];
        const int fidx = (int)(e >> 22);                                // (fy << 5) | fx
        const uint4 wa = tex1Dfetch<uint4>(wtex, fidx << 1), wb = tex1Dfetch<uint4>(wtex, (fidx << 1) + 1);
        const uint32_t wpk[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        const uint32_t *p = sm + (int)((e >> 11) & 2047u) * P + (int)(e & 2047u) + sbase;
        uint32_t t[16];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) t[4 * r + q] = p[r * P + q];
        int a0 = 16384, a1 = 16384, a2 = 16384;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t bg = __byte_perm(t[2 * k], t[2 * k + 1], 0x5140);
            a0 = dp2a_su((int)wpk[k], bg, a0);
            a1 = dp2a_su_hi((int)wpk[k], bg, a1);
            a2 = dp2a_su((int)wpk[k], __byte_perm(t[2 * k], t[2 * k + 1], 0x0062), a2);
        }
        uint32_t hi, px;
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(0), "r"(a2 >> 15));
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(px) : "r"(a1 >> 15), "r"(a0 >> 15), "r"(hi));
        d[32 * (j & 1) + (size_t)(8 * (j >> 1)) * rw] = px;
    }
}

template <bool kTma>
__global__ void __launch_bounds__(256) cubic5_kernel(const __grid_constant__ CubicArgs A)
{
    pdl_enter();
    __shared__ __align__(128) uint32_t sm[kCubSmemWords];
    __shared__ __align__(8) uint64_t bar;
    cubic5_body<kTma>(A, sm, &bar, blockIdx.x, blockIdx.y, blockIdx.z);
}

// cv::resize INTER_LINEAR from a word-per-pixel source to packed BGR bytes.  One warp row-chunk =
// 32 consecutive output pixels = 96 bytes, assembled in shared memory and stored as 24 words.
__global__ void __launch_bounds__(256) resize4_kernel(const uint32_t *__restrict__ src, size_t src_img_words, int sstride_words,
                                                      uint8_t *__restrict__ dst, size_t dst_img, int dstride, ResizeTab t)
{
    __shared__ __align__(16) uint8_t sm[8][4][96];
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int lane = threadIdx.x;
    const bool rowok = y < t.dh;
    const uint32_t *s = src + (size_t)blockIdx.z * src_img_words;
    const int yo = rowok ? t.yofs[y] : 0;
    const int sy0 = min(max(yo, 0), t.sh - 1), sy1 = min(max(yo + 1, 0), t.sh - 1);
    const short2 ay = rowok ? t.ya[y] : make_short2(0, 0);
    const uint32_t *r0 = s + (size_t)sy0 * sstride_words, *r1 = s + (size_t)sy1 * sstride_words;
    int xo[4];
    short2 ax[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = blockIdx.x * 128 + 32 * j + lane;
        xo[j] = (rowok && x < t.dw) ? t.xofs[x] : 0;
        ax[j] = (rowok && x < t.dw) ? t.xa[x] : make_short2(0, 0);
    }
    uint32_t p[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x1 = min(xo[j] + 1, t.sw - 1);
        p[j][0] = __ldg(r0 + xo[j]); p[j][1] = __ldg(r0 + x1); p[j][2] = __ldg(r1 + xo[j]); p[j][3] = __ldg(r1 + x1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int axp = (int)(((uint32_t)(uint16_t)ax[j].y << 16) | (uint16_t)ax[j].x);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t sel = 0x0040u + 0x11u * c;                 // byte c of both taps
            const int t0 = dp2a_su(axp, __byte_perm(p[j][0], p[j][1], sel), 0);
            const int t1 = dp2a_su(axp, __byte_perm(p[j][2], p[j][3], sel), 0);
            sm[threadIdx.y][j][lane * 3 + c] = (uint8_t)sat_u8((((ay.x * (t0 >> 4)) >> 16) + ((ay.y * (t1 >> 4)) >> 16) + 2) >> 2);
        }
    }
    __syncwarp();
    if (!rowok) return;
    uint8_t *drow = dst + (size_t)blockIdx.z * dst_img + (size_t)y * dstride;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x0 = blockIdx.x * 128 + 32 * j;
        if (x0 >= t.dw) break;
        uint8_t *o = drow + (size_t)x0 * 3;
        if (x0 + 32 <= t.dw && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
            if (lane < 24) reinterpret_cast<uint32_t *>(o)[lane] = reinterpret_cast<const uint32_t *>(sm[threadIdx.y][j])[lane];
        } else {
            for (int b = lane; b < 3 * min(32, t.dw - x0); b += 32) o[b] = sm[threadIdx.y][j][b];
        }
    }
}

// cv::resize INTER_LINEAR (up-scaling rows), word-per-pixel source -> packed BGR bytes, "row walker".
// cv::resize is separable: output row y blends the horizontally interpolated source rows v = yofs[y] and
// v + 1 (both clamped), and the horizontal result of a source row does not depend on y.  A lane owns ONE
// output column and walks down a band of kResizeBand consecutive v, keeping the horizontal results of the
// two current source rows in registers (3 channels each): every source row is fetched and filtered once per
// band instead of once per output row that touches it (2.4x fewer taps and horizontal products for the
// 889 -> 1080 up-scale), and the column tables are read once per band.  Nothing inside the walk depends on
// a load issued in the same step: source rows are fetched two steps ahead, and the per-row tables (first
// output row of every v -- ResizeTab::ybeg -- and the row coefficients) sit in lane registers and are
// broadcast with shuffles.  The vertical pass is two multiply-high + one shift per channel
// ((a * h) >> 16 == umulhi(a << 16, h) for the non-negative bilinear coefficients).  Four lanes' pixels
// (12 bytes) leave as three aligned words built with one shuffle + one byte permute per lane.
constexpr int kResizeBand = 24;      // virtual source rows per block
constexpr int kResizeRows = 64;      // most output rows one band may produce (row coefficients live in shared memory)

// requires: dw % 32 == 0, dst rows 4-byte aligned, <= kResizeRows output rows per band (checked at launch)
struct ResizeArgs {
    const uint32_t *src; size_t src_img_words; unsigned sstride_words;
    uint8_t *dst; size_t dst_img; unsigned dstride;
    ResizeTab t;
    int gx, gy;                      // logical grid of 128-thread blocks = gx x gy x images
};
constexpr int kResizeSmemWords = kResizeBand + 1 + kResizeRows;

// one 128-thread block (4 warps: wy = 0..3) = column group bx, band by of image bz.
// kWords: the output is one 32-bit word per pixel (B | G << 8 | R << 16; dstride / dst_img still in bytes) instead of
// packed BGR -- the layout the rotation warp's staged gather (warp_tile_kernel<.., kSrc4>) consumes directly when the
// front end is chained in front of a stitcher handle (an internal buffer: the caller never sees it).
template <bool kWords>
__device__ __forceinline__ void resize4_walk_body(const ResizeArgs &A, uint32_t *__restrict__ smem, int bx, int by, int bz, int wy)
{
    const uint32_t *__restrict__ src = A.src;
    uint8_t *__restrict__ dst = A.dst;
    const size_t src_img_words = A.src_img_words, dst_img = A.dst_img;
    const unsigned sstride_words = A.sstride_words, dstride = A.dstride;
    const ResizeTab &t = A.t;
    int *s_yb = reinterpret_cast<int *>(smem);
    uint32_t *s_ay = smem + kResizeBand + 1;
    const int lane = threadIdx.x, tid = wy * 32 + lane;
    const int sh = t.sh;
    const int v0 = by * kResizeBand - 1;                               // v runs over [-1, sh - 1]
    const int nv = min(kResizeBand, sh - v0);
    const int ybase = __ldg(t.ybeg + v0 + 1);
    if (tid <= nv) s_yb[tid] = __ldg(t.ybeg + v0 + 1 + tid);           // first output row of virtual row v0 + tid
    if (tid >= 64) {
        const short2 q = __ldg(t.ya + min(ybase + tid - 64, t.dh - 1));
        s_ay[tid - 64] = ((uint32_t)(uint16_t)q.y << 16) | (uint16_t)q.x;
    }
    __syncthreads();
    const int xw = (bx * 4 + wy) * 32;                                  // first column of this warp
    if (xw >= t.dw) return;
    const int x = xw + lane;
    const unsigned xo = __ldg(t.xofs + x), x1 = min(xo + 1u, (unsigned)t.sw - 1u);
    const short2 ax = __ldg(t.xa + x);
    const int axp = (int)(((uint32_t)(uint16_t)ax.y << 16) | (uint16_t)ax.x);
    const uint32_t *s0 = src + (size_t)bz * src_img_words + xo;
    const uint32_t *s1 = src + (size_t)bz * src_img_words + x1;
    const int j = lane & 3;
    const uint32_t sel = j == 0 ? 0x4210u : (j == 1 ? 0x5421u : 0x6542u);
    // lanes 4k..4k+2 store the three words of pixels 4k..4k+3; lane 4k+3 stores nothing
    uint8_t *olane = dst + (size_t)bz * dst_img + (size_t)ybase * dstride +
                     (kWords ? (size_t)x * 4 : (size_t)xw * 3 + ((lane >> 2) * 3 + j) * 4);
    const uint32_t *ayp = s_ay;
    const unsigned sstride_bytes = sstride_words * 4u;

#define RS_FETCH(v, A, B)                                                        \
    {                                                                            \
        const size_t rb_ = (size_t)((unsigned)min(max((v), 0), sh - 1) * sstride_bytes);  \
        A = __ldg(reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s0) + rb_)); \
        B = __ldg(reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s1) + rb_)); \
    }
#define RS_HCALC(A, B, H0, H1, H2)                                               \
    H0 = (uint32_t)dp2a_su(axp, __byte_perm(A, B, 0x5140u), 0) >> 4;             \
    H1 = (uint32_t)dp2a_su_hi(axp, __byte_perm(A, B, 0x5140u), 0) >> 4;          \
    H2 = (uint32_t)dp2a_su(axp, __byte_perm(A, B, 0x0062u), 0) >> 4;
    // one step: emit the output rows that blend (A, B) = rows (v, v + 1), then row v + 2 replaces A (it is the
    // next step's B) and the prefetch slot is refilled with row v + 4
#define RS_STEP(i, A0, A1, A2, B0, B1, B2, PA, PB)                               \
    {                                                                            \
        _Pragma("unroll 1") for (int n_ = s_yb[(i) + 1] - s_yb[(i)]; n_ > 0; --n_) { \
            const uint32_t w_ = *ayp++;                                          \
            const uint32_t a0_ = w_ << 16, a1_ = w_ & 0xffff0000u;               \
            const uint32_t c0_ = (__umulhi(a0_, A0) + 2u + __umulhi(a1_, B0)) >> 2;   \
            const uint32_t c1_ = (__umulhi(a0_, A1) + 2u + __umulhi(a1_, B1)) >> 2;   \
            const uint32_t c2_ = (__umulhi(a0_, A2) + 2u + __umulhi(a1_, B2)) >> 2;   \
            const uint32_t px_ = __byte_perm(__byte_perm(c0_, c1_, 0x0040), c2_, 0x4410); \
            if (kWords) {                                                        \
                *reinterpret_cast<uint32_t *>(olane) = px_;                      \
            } else {                                                             \
                const uint32_t nx_ = __shfl_down_sync(0xffffffffu, px_, 1);      \
                if (j < 3) *reinterpret_cast<uint32_t *>(olane) = __byte_perm(px_, nx_, sel); \
            }                                                                    \
            olane += dstride;                                                    \
        }                                                                        \
        RS_HCALC(PA, PB, A0, A1, A2)                                             \
        RS_FETCH(v0 + (i) + 4, PA, PB)                                           \
    }

    uint32_t g0, g1, g2, h0, h1, h2, p0a, p0b, p1a, p1b;
    {
        uint32_t a, b, c, e;
        RS_FETCH(v0, a, b) RS_FETCH(v0 + 1, c, e)
        RS_FETCH(v0 + 2, p0a, p0b) RS_FETCH(v0 + 3, p1a, p1b)
        RS_HCALC(a, b, g0, g1, g2) RS_HCALC(c, e, h0, h1, h2)
    }
    for (int i = 0; i < nv; i += 2) {
        RS_STEP(i, g0, g1, g2, h0, h1, h2, p0a, p0b)
        if (i + 1 < nv) RS_STEP(i + 1, h0, h1, h2, g0, g1, g2, p1a, p1b)
    }
#undef RS_STEP
#undef RS_HCALC
#undef RS_FETCH
}

// plain multiply-high: written as PTX so that the compiler does not fold the following add into IMAD.HI's 64-bit
// accumulator operand (which needs a zeroed register pair per use and ends up longer)
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// The same walk with kCols (2 or 4) adjacent output columns per lane (word output only: the chained hot path).  The
// kernel is issue-bound, and a third of its instructions do not depend on the column at all (row-coefficient fetch,
// loop control, the first-output-row table, the row index, the store address): with several pixels per lane they are
// paid once per group, the group leaves as one 8- or 16-byte store, and tap addresses are 32-bit word indices (one add
// + one widening multiply-add per load instead of a 64-bit add chain; the launcher checks the batch has < 2^32 words).
// requires dw % (32 * kCols) == 0 and dst rows aligned to 4 * kCols bytes (checked at launch).
template <int kCols>
__device__ __forceinline__ void resize4_walkn_body(const ResizeArgs &A, uint32_t *__restrict__ smem, int bx, int by, int bz, int wy)
{
    const ResizeTab &t = A.t;
    int *s_yb = reinterpret_cast<int *>(smem);
    uint32_t *s_ay = smem + kResizeBand + 1;
    const int lane = threadIdx.x, tid = wy * 32 + lane;
    const int sh = t.sh;
    const int v0 = by * kResizeBand - 1;
    const int nv = min(kResizeBand, sh - v0);
    const int ybase = __ldg(t.ybeg + v0 + 1);
    if (tid <= nv) s_yb[tid] = __ldg(t.ybeg + v0 + 1 + tid);
    if (tid >= 64) {
        const short2 q = __ldg(t.ya + min(ybase + tid - 64, t.dh - 1));
        s_ay[tid - 64] = ((uint32_t)(uint16_t)q.y << 16) | (uint16_t)q.x;
    }
    __syncthreads();
    const int xw = (bx * 4 + wy) * 32 * kCols;
    if (xw >= t.dw) return;
    const int x = xw + kCols * lane;
    int xo[kCols], ax[kCols];                       // ax: short2 (a0 | a1 << 16) = exactly the dp2a operand
    if (kCols == 4) {
        const int4 q = __ldg(reinterpret_cast<const int4 *>(t.xofs + x)), r = __ldg(reinterpret_cast<const int4 *>(t.xa + x));
        xo[0] = q.x; xo[1] = q.y; xo[kCols - 2] = q.z; xo[kCols - 1] = q.w;
        ax[0] = r.x; ax[1] = r.y; ax[kCols - 2] = r.z; ax[kCols - 1] = r.w;
    } else {
        const int2 q = __ldg(reinterpret_cast<const int2 *>(t.xofs + x)), r = __ldg(reinterpret_cast<const int2 *>(t.xa + x));
        xo[0] = q.x; xo[1] = q.y;
        ax[0] = r.x; ax[1] = r.y;
    }
    // word indices of the taps in row 0 of this image, relative to A.src
    const unsigned img0 = (unsigned)bz * (unsigned)A.src_img_words;
    unsigned ti[2 * kCols];
#pragma unroll
    for (int c = 0; c < kCols; ++c) {
        ti[2 * c] = img0 + (unsigned)xo[c];
        ti[2 * c + 1] = img0 + min((unsigned)xo[c] + 1u, (unsigned)t.sw - 1u);
    }
    const uint32_t *__restrict__ simg = A.src;
    const unsigned sstride_words = A.sstride_words, dstride = A.dstride;
    uint8_t *olane = A.dst + (size_t)bz * A.dst_img + (size_t)ybase * dstride + (size_t)x * 4;
    const uint32_t *ayp = s_ay;

    auto fetch = [&](int v, uint32_t (&P)[2 * kCols]) {
        const unsigned ri = (unsigned)min(max(v, 0), sh - 1) * sstride_words;
#pragma unroll
        for (int k = 0; k < 2 * kCols; ++k) P[k] = __ldg(simg + (ri + ti[k]));
    };
    auto hcalc = [&](const uint32_t (&P)[2 * kCols], uint32_t (&H)[3 * kCols]) {
#pragma unroll
        for (int c = 0; c < kCols; ++c) {
            const uint32_t bg = __byte_perm(P[2 * c], P[2 * c + 1], 0x5140u);
            H[3 * c] = (uint32_t)dp2a_su(ax[c], bg, 0) >> 4;
            H[3 * c + 1] = (uint32_t)dp2a_su_hi(ax[c], bg, 0) >> 4;
            H[3 * c + 2] = (uint32_t)dp2a_su(ax[c], __byte_perm(P[2 * c], P[2 * c + 1], 0x0062u), 0) >> 4;
        }
    };
    // one step: emit the output rows that blend (HA, HB) = rows (v, v + 1), then row v + 2 replaces HA (it is the next
    // step's HB) and the prefetch slot is refilled with row v + 4
    auto step = [&](int i, uint32_t (&HA)[3 * kCols], const uint32_t (&HB)[3 * kCols], uint32_t (&P)[2 * kCols]) {
#pragma unroll 1
        for (int n = s_yb[i + 1] - s_yb[i]; n > 0; --n) {
            const uint32_t w = *ayp++;
            const uint32_t a0 = w << 16, a1 = w & 0xffff0000u;
            uint32_t px[kCols];
#pragma unroll
            for (int c = 0; c < kCols; ++c) {
                const uint32_t c0 = (mul_hi(a0, HA[3 * c]) + mul_hi(a1, HB[3 * c]) + 2u) >> 2;
                const uint32_t c1 = (mul_hi(a0, HA[3 * c + 1]) + mul_hi(a1, HB[3 * c + 1]) + 2u) >> 2;
                const uint32_t c2 = (mul_hi(a0, HA[3 * c + 2]) + mul_hi(a1, HB[3 * c + 2]) + 2u) >> 2;
                px[c] = __byte_perm(__byte_perm(c0, c1, 0x0040), c2, 0x4410);
            }
            if (kCols == 4) *reinterpret_cast<uint4 *>(olane) = make_uint4(px[0], px[1], px[kCols - 2], px[kCols - 1]);
            else *reinterpret_cast<uint2 *>(olane) = make_uint2(px[0], px[1]);
            olane += dstride;
        }
        hcalc(P, HA);
        fetch(v0 + i + 4, P);
    };

    uint32_t g[3 * kCols], hh[3 * kCols], p0[2 * kCols], p1[2 * kCols];
    {
        uint32_t a[2 * kCols], b[2 * kCols];
        fetch(v0, a); fetch(v0 + 1, b);
        fetch(v0 + 2, p0); fetch(v0 + 3, p1);
        hcalc(a, g); hcalc(b, hh);
    }
    for (int i = 0; i < nv; i += 2) {
        step(i, g, hh, p0);
        if (i + 1 < nv) step(i + 1, hh, g, p1);
    }
}

template <int kCols, int kMinBlocks>
__global__ void __launch_bounds__(128, kMinBlocks) resize4_walkn_kernel(const __grid_constant__ ResizeArgs A)
{
    pdl_enter();
    __shared__ uint32_t smem[kResizeSmemWords];
    resize4_walkn_body<kCols>(A, smem, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.y);
}

template <bool kWords>
__global__ void __launch_bounds__(128, 8) resize4_walk_kernel(const __grid_constant__ ResizeArgs A)
{
    pdl_enter();
    __shared__ uint32_t smem[kResizeSmemWords];
    resize4_walk_body<kWords>(A, smem, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.y);
}

thread_local std::string g_front_error;

}  // namespace

// The front end's intermediate images.  A handle has one set of its own (host-buffer API, standalone device calls); a
// stitcher handle that chains the front end in brings ITS OWN set (pano_frontend_scratch_*), so that one front-end handle
// -- which otherwise only holds read-only tables -- can serve several stitchers running on different streams / threads
// at the same time (the SDK facade's two rings share the camera model).
struct pano_front_scratch {
    uint8_t *a = nullptr, *b = nullptr, *c = nullptr;  // generic path: undist-sized 3-channel intermediates
    uint32_t *w = nullptr;                             // fast path: cropped undistorted image, one word per pixel
    uint8_t *argb = nullptr;                           // YUYV ingest: converted 8UC4 frames
    int depth = 0;                                     // images each buffer holds
    int device = 0;
};

struct pano_frontend_ctx {
    pano_frontend_config cfg{};
    std::string err;
    std::vector<float> mapx, mapy;
    std::vector<void *> owned;
    int2 *dmap = nullptr;
    short *dtab = nullptr;
    cudaTextureObject_t wtex = 0;                                  // dtab as a linear uint4 texture
    ResizeTab r_in, r_mid, r_out;       // cam->undist, rect->undist, undist->out
    bool use_r_in = false, use_r_mid = false, use_r_out = false;
    pano_front_scratch own;                                        // intermediates, max_batch deep
    uint32_t *dmap32 = nullptr;                                    // packed map entries (fast path, sources <= 2043 px)
    bool fast4 = false;
    int4 *cub_tiles = nullptr;                                     // per 128x8 tile of the crop rect: staged footprint (cubic5_kernel)
    int cub_tx = 0, cub_ty = 0;
    bool cub_tma = false;                                          // every tile footprint fits the TMA boxes (kCubBoxW x kCubBoxH1)
    cudaEvent_t *prof_ev = nullptr;                                // events when profiling: before cubic, between, after resize, [3] before the YUYV conversion
    int in_px() const { return cfg.src_format == PANO_SRC_YUYV ? 2 : 4; }
    uint8_t *stage_in = nullptr, *stage_out = nullptr;
    int launches = 0;
};

namespace {

int ffail(pano_frontend_ctx *h, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_front_error = buf;
    return PANO_ERR;
}

#define FCK(h, call)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return ffail(h, "%s failed: %s", #call, cudaGetErrorString(e_));   \
    } while (0)

template <typename T>
int falloc(pano_frontend_ctx *h, T **p, size_t count)
{
    FCK(h, cudaMalloc((void **)p, std::max<size_t>(1, count) * sizeof(T)));
    h->owned.push_back(*p);
    return PANO_OK;
}

int scratchAlloc(pano_frontend_ctx *h, pano_front_scratch &s, int depth)
{
    const pano_frontend_config &c = h->cfg;
    const size_t ubytes = (size_t)c.undist_width * c.undist_height * 3 * depth;
    s.depth = depth;
    s.device = c.device;
    if (c.src_format == PANO_SRC_YUYV) FCK(h, cudaMalloc((void **)&s.argb, (size_t)c.cam_src_width * c.cam_src_height * 4 * depth));
    if (h->fast4) FCK(h, cudaMalloc((void **)&s.w, (size_t)c.rect[2] * c.rect[3] * depth * sizeof(uint32_t)));
    // the three undist-sized intermediates of the generic path (3 x 1.6 GB at 256 x 1080p) are only needed where the fast
    // path does not apply (no undistortion, extra resizes) or is refused (camera frames not 16-byte aligned): scratchGeneric
    if (!h->fast4) {
        FCK(h, cudaMalloc((void **)&s.a, ubytes));
        FCK(h, cudaMalloc((void **)&s.b, ubytes));
        FCK(h, cudaMalloc((void **)&s.c, ubytes));
    }
    return PANO_OK;
}

// first use of the generic path with a set that was created for the fast path: allocate now (not possible inside a
// stream capture -- the caller's frames then have to be 16-byte aligned)
int scratchGeneric(pano_frontend_ctx *h, pano_front_scratch &s, cudaStream_t st)
{
    if (s.a) return PANO_OK;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        (void)cudaGetLastError();
        return ffail(h, "pano_frontend_run: camera frames must be 16-byte aligned here (the byte-wise path needs buffers that cannot be allocated inside a stream capture)");
    }
    const size_t ubytes = (size_t)h->cfg.undist_width * h->cfg.undist_height * 3 * s.depth;
    FCK(h, cudaMalloc((void **)&s.a, ubytes));
    FCK(h, cudaMalloc((void **)&s.b, ubytes));
    FCK(h, cudaMalloc((void **)&s.c, ubytes));
    return PANO_OK;
}

void scratchFree(pano_front_scratch &s)
{
    cudaFree(s.a); cudaFree(s.b); cudaFree(s.c); cudaFree(s.w); cudaFree(s.argb);
    s = pano_front_scratch();
}

int makeResize(pano_frontend_ctx *h, ResizeTab &t, int sw, int sh, int dw, int dh)
{
    t.sw = sw; t.sh = sh; t.dw = dw; t.dh = dh;
    std::vector<int> xo, yo;
    std::vector<int16_t> xa0, xa1, ya0, ya1;
    resizeAxis(sw, dw, true, xo, xa0, xa1);
    resizeAxis(sh, dh, false, yo, ya0, ya1);
    std::vector<short2> xa(dw), ya(dh);
    for (int i = 0; i < dw; ++i) xa[i] = make_short2(xa0[i], xa1[i]);
    for (int i = 0; i < dh; ++i) ya[i] = make_short2(ya0[i], ya1[i]);
    if (falloc(h, &t.xofs, dw) || falloc(h, &t.yofs, dh) || falloc(h, &t.xa, dw) || falloc(h, &t.ya, dh)) return PANO_ERR;
    FCK(h, cudaMemcpy(t.xofs, xo.data(), dw * sizeof(int), cudaMemcpyHostToDevice));
    FCK(h, cudaMemcpy(t.yofs, yo.data(), dh * sizeof(int), cudaMemcpyHostToDevice));
    FCK(h, cudaMemcpy(t.xa, xa.data(), dw * sizeof(short2), cudaMemcpyHostToDevice));
    FCK(h, cudaMemcpy(t.ya, ya.data(), dh * sizeof(short2), cudaMemcpyHostToDevice));
    // row walker tables: yofs is non-decreasing and lies in [-1, sh - 1]
    std::vector<int> yb(sh + 2);
    bool mono = true, nonneg = true;
    for (int i = 1; i < dh; ++i) mono = mono && yo[i] >= yo[i - 1];
    for (int i = 0; i < dh; ++i) nonneg = nonneg && ya0[i] >= 0 && ya1[i] >= 0 && yo[i] >= -1 && yo[i] <= sh - 1;
    for (int i = 0; i < dw; ++i) nonneg = nonneg && xa0[i] >= 0 && xa1[i] >= 0;
    for (int v = -1, y = 0; v <= sh; ++v) {
        while (y < dh && yo[y] < v) ++y;
        yb[v + 1] = y;
    }
    if (falloc(h, &t.ybeg, yb.size())) return PANO_ERR;
    FCK(h, cudaMemcpy(t.ybeg, yb.data(), yb.size() * sizeof(int), cudaMemcpyHostToDevice));
    bool fits = true;                                   // every band's output rows fit the kernel's coefficient table
    for (int v = -1; v <= sh - 1; v += kResizeBand) fits = fits && yb[std::min(v + kResizeBand, sh) + 1] - yb[v + 1] <= kResizeRows;
    t.walk = mono && nonneg && fits && dh >= sh && dw % 32 == 0;
    return PANO_OK;
}

inline dim3 grid2(int w, int hh, dim3 b, int z) { return dim3((w + b.x - 1) / b.x, (hh + b.y - 1) / b.y, z); }

}  // namespace

extern "C" {

const char *pano_frontend_last_error(pano_frontend_handle h) { return h ? h->err.c_str() : g_front_error.c_str(); }

int pano_frontend_create(const pano_frontend_config *cfg, pano_frontend_handle *out)
{
    if (!cfg || !out) return ffail(nullptr, "pano_frontend_create: null argument");
    *out = nullptr;
    if (cfg->cam_src_width < 2 || cfg->cam_src_height < 2 || cfg->undist_width < 2 || cfg->undist_height < 2 ||
        cfg->out_width < 2 || cfg->out_height < 2)
        return ffail(nullptr, "bad sizes");
    int ndev = 0;
    if (cfg->src_format != PANO_SRC_BGRA && cfg->src_format != PANO_SRC_YUYV) return ffail(nullptr, "bad src_format");
    if (cfg->src_format == PANO_SRC_YUYV && (cfg->cam_src_width & 1)) return ffail(nullptr, "YUYV frames need an even width");
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ffail(nullptr, "no CUDA device: this library has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return ffail(nullptr, "bad device ordinal");
    pano_frontend_ctx *h = new pano_frontend_ctx();
    h->cfg = *cfg;
    h->cfg.max_batch = std::max(1, cfg->max_batch);
    auto bail = [&]() { g_front_error = h->err; pano_frontend_destroy(h); return PANO_ERR; };
    if (cudaSetDevice(cfg->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return bail(); }
    const int S = h->cfg.max_batch;
    const int uw = cfg->undist_width, uh = cfg->undist_height;
    if (cfg->undistort) {
        const int *rc = cfg->rect;
        if (rc[0] < 0 || rc[1] < 0 || rc[2] < 1 || rc[3] < 1 || rc[0] + rc[2] > uw || rc[1] + rc[3] > uh) {
            h->err = "crop rect outside the undistorted image";
            return bail();
        }
        h->mapx.resize((size_t)uw * uh);
        h->mapy.resize((size_t)uw * uh);
        if (cfg->mapx && cfg->mapy) {
            std::memcpy(h->mapx.data(), cfg->mapx, h->mapx.size() * sizeof(float));
            std::memcpy(h->mapy.data(), cfg->mapy, h->mapy.size() * sizeof(float));
        } else {
            undistortMaps(cfg->K, cfg->D, cfg->newK, uw, uh, h->mapx.data(), h->mapy.data());
        }
        std::vector<int2> fm((size_t)uw * uh);
        for (size_t i = 0; i < fm.size(); ++i) {
            const FixedCoord fc = toFixed(h->mapx[i], h->mapy[i]);
            fm[i] = make_int2(fc.ix * 32 + fc.fx, fc.iy * 32 + fc.fy);
        }
        std::vector<int16_t> tab(1024 * 16);
        cubicTable(tab.data());
        if (falloc(h, &h->dmap, fm.size()) || falloc(h, &h->dtab, tab.size())) return bail();
        if (cudaMemcpy(h->dmap, fm.data(), fm.size() * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(h->dtab, tab.data(), tab.size() * sizeof(short), cudaMemcpyHostToDevice) != cudaSuccess) {
            h->err = "front-end table upload failed";
            return bail();
        }
        {
            cudaResourceDesc rd{};
            rd.resType = cudaResourceTypeLinear;
            rd.res.linear.devPtr = h->dtab;
            rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
            rd.res.linear.sizeInBytes = tab.size() * sizeof(short);
            cudaTextureDesc td{};
            td.readMode = cudaReadModeElementType;
            if (cudaCreateTextureObject(&h->wtex, &rd, &td, nullptr) != cudaSuccess) { h->err = "weight texture creation failed"; return bail(); }
        }
        h->use_r_in = !(cfg->cam_src_width == uw && cfg->cam_src_height == uh);
        h->use_r_mid = !(rc[2] == uw && rc[3] == uh);
        if (h->use_r_in && makeResize(h, h->r_in, cfg->cam_src_width, cfg->cam_src_height, uw, uh)) return bail();
        if (h->use_r_mid && makeResize(h, h->r_mid, rc[2], rc[3], uw, uh)) return bail();
    } else {
        h->use_r_in = !(cfg->cam_src_width == uw && cfg->cam_src_height == uh);
        if (h->use_r_in && makeResize(h, h->r_in, cfg->cam_src_width, cfg->cam_src_height, uw, uh)) return bail();
    }
    h->use_r_out = !(cfg->out_width == uw && cfg->out_height == uh);
    if (h->use_r_out && makeResize(h, h->r_out, uw, uh, cfg->out_width, cfg->out_height)) return bail();
    // fast path: undistort straight from the 8UC4 camera frame, then one resize to the output size
    h->fast4 = cfg->undistort && !h->use_r_in && h->use_r_mid && !h->use_r_out;
    if (scratchAlloc(h, h->own, S)) return bail();
    if (h->fast4 && cfg->cam_src_width + 4 <= 2047 && cfg->cam_src_height + 4 <= 2047) {
        std::vector<uint32_t> pm(h->mapx.size());
        for (size_t i = 0; i < pm.size(); ++i) {
            const FixedCoord fc = toFixed(h->mapx[i], h->mapy[i]);
            const int tx = std::max(-4, std::min(cfg->cam_src_width, fc.ix - 1));
            const int ty = std::max(-4, std::min(cfg->cam_src_height, fc.iy - 1));
            pm[i] = (uint32_t)(tx + 4) | ((uint32_t)(ty + 4) << 11) | ((uint32_t)fc.fx << 22) | ((uint32_t)fc.fy << 27);
        }
        if (falloc(h, &h->dmap32, pm.size())) return bail();
        if (cudaMemcpy(h->dmap32, pm.data(), pm.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            h->err = "front-end packed map upload failed";
            return bail();
        }
        // source footprint of every 128x8 output tile (cubic5_kernel); all tiles must fit its staging buffer
        const int *rc = cfg->rect;
        const int tx = (rc[2] + kCubTileW - 1) / kCubTileW, ty = (rc[3] + kCubTileH - 1) / kCubTileH;
        std::vector<int4> tl((size_t)tx * ty);
        bool fits = cfg->cam_src_width % 4 == 0, tma_fits = true;
        for (int by = 0; by < ty && fits; ++by)
            for (int bx = 0; bx < tx && fits; ++bx) {
                int x0 = INT32_MAX, x1 = INT32_MIN, y0 = INT32_MAX, y1 = INT32_MIN;
                for (int y = by * kCubTileH; y < std::min(rc[3], (by + 1) * kCubTileH); ++y)
                    for (int x = bx * kCubTileW; x < std::min(rc[2], (bx + 1) * kCubTileW); ++x) {
                        const uint32_t e = pm[(size_t)(y + rc[1]) * uw + (x + rc[0])];
                        const int ix = (int)(e & 2047u) - 4, iy = (int)((e >> 11) & 2047u) - 4;
                        x0 = std::min(x0, ix); x1 = std::max(x1, ix + 3);
                        y0 = std::min(y0, iy); y1 = std::max(y1, iy + 3);
                    }
                const int ax0 = (x0 >= 0 ? x0 / 4 : -((-x0 + 3) / 4)) * 4;           // floor to a multiple of 4
                const int w4 = (x1 - ax0) / 4 + 1, rows = y1 - y0 + 1;
                fits = w4 <= 64 && rows * ((w4 * 4 + 31) & ~31) <= kCubSmemWords && rows * w4 < 1024;
                tma_fits = tma_fits && w4 * 4 <= kCubBoxW && rows <= kCubBoxH1;
                tl[(size_t)by * tx + bx] = make_int4(ax0, y0, rows | (w4 << 16), (65536 + w4 - 1) / w4);
            }
        if (getenv("PANO_DEBUG")) fprintf(stderr, "[panob200] cubic tiles %dx%d of %dx%d: %s\n", tx, ty, kCubTileW, kCubTileH, fits ? "staged" : "footprint too large -> untiled kernel");
        if (fits) {
            if (falloc(h, &h->cub_tiles, tl.size())) return bail();
            if (cudaMemcpy(h->cub_tiles, tl.data(), tl.size() * sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess) {
                h->err = "front-end tile table upload failed";
                return bail();
            }
            h->cub_tx = tx; h->cub_ty = ty;
            h->cub_tma = tma_fits;
        }
    }
    *out = h;
    return PANO_OK;
}

int pano_frontend_destroy(pano_frontend_handle h)
{
    if (!h) return PANO_OK;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    if (h->wtex) cudaDestroyTextureObject(h->wtex);
    for (void *p : h->owned) cudaFree(p);
    scratchFree(h->own);
    delete h;
    return PANO_OK;
}

int pano_frontend_get_maps(pano_frontend_handle h, float *mapx, float *mapy)
{
    if (!h || h->mapx.empty()) return ffail(h, "no undistort maps");
    if (mapx) std::memcpy(mapx, h->mapx.data(), h->mapx.size() * sizeof(float));
    if (mapy) std::memcpy(mapy, h->mapy.data(), h->mapy.size() * sizeof(float));
    return PANO_OK;
}

}  // extern "C"

// A private set of the intermediates for a caller that chains this front end into its own stream (capi.cu: one per
// stitcher handle and distinct front end): as deep as the handle's own (max_batch images).
pano_front_scratch *pano_frontend_scratch_create(pano_frontend_handle h)
{
    if (!h) return nullptr;
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) return nullptr;
    pano_front_scratch *s = new pano_front_scratch();
    if (scratchAlloc(h, *s, h->cfg.max_batch)) { scratchFree(*s); delete s; return nullptr; }
    return s;
}

void pano_frontend_scratch_destroy(pano_front_scratch *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    scratchFree(*s);
    delete s;
}


// YUYV ingest: `count` frames (in_img bytes apart) -> dense 8UC4 frames at dst
int pano_frontend_convert(pano_frontend_handle h, const uint8_t *yuyv, size_t in_img, uint8_t *dst, int count, cudaStream_t st)
{
    const pano_frontend_config &c = h->cfg;
    const size_t npx = (size_t)c.cam_src_width * c.cam_src_height, out_img = npx * 4;
    if (npx % 8 == 0 && in_img % 16 == 0 && (reinterpret_cast<uintptr_t>(yuyv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0)
        launch_chain(yuyv_to_bgra_kernel, dim3((unsigned)((npx / 8 + 255) / 256), count), dim3(256), st,
                     reinterpret_cast<const uint4 *>(yuyv), in_img / 16, reinterpret_cast<uint4 *>(dst), out_img / 16, npx / 8);
    else
        yuyv_to_bgra_pair_kernel<<<dim3((unsigned)((npx / 2 + 255) / 256), count), 256, 0, st>>>(yuyv, in_img, dst, out_img, npx / 2);
    FCK(h, cudaGetLastError());
    return PANO_OK;
}
int pano_frontend_in_px(pano_frontend_handle h) { return h ? h->in_px() : 4; }

// Internal entry (also used by capi.cu when a front end is attached to a stitcher handle):
// images may be strided (in_img / o_img bytes between consecutive images).

// out_px: 3 = packed BGR (the public layout), 4 = one word per pixel (internal hand-over to the rotation warp; only
// when pano_frontend_can_words(h))
bool pano_frontend_can_words(pano_frontend_handle h)
{
    static const bool no_words = getenv("PANO_FE_NO_WORDS") != nullptr;   // A/B switch: packed BGR hand-over
    return h && h->fast4 && h->r_mid.walk && !no_words && !getenv("PANO_NO_RESIZE_WALK");
}

int pano_frontend_run(pano_frontend_handle h, const uint8_t *argb, size_t in_img, uint8_t *out, size_t o_img, int batch,
                      cudaStream_t st, int out_px, pano_front_scratch *scratch)
{
    if (!h || !argb || !out || batch < 1) return ffail(h, "pano_frontend_run: bad argument");
    pano_front_scratch &B = scratch ? *scratch : h->own;            // the caller's set of intermediates, or the handle's own
    if (out_px != 3 && !(out_px == 4 && pano_frontend_can_words(h))) return ffail(h, "pano_frontend_run: word output is not available for this geometry");
    FCK(h, cudaSetDevice(h->cfg.device));
    const pano_frontend_config &c = h->cfg;
    const int S = B.depth, uw = c.undist_width, uh = c.undist_height;
    const size_t u_img = (size_t)uw * uh * 3;
    const dim3 blk(32, 8);
    h->launches = 0;
    for (int b0 = 0; b0 < batch; b0 += S) {
        const int nb = std::min(S, batch - b0);
        const uint8_t *src = argb + (size_t)b0 * in_img;
        uint8_t *final_dst = out + (size_t)b0 * o_img;
        size_t img_stride = in_img;                     // bytes between consecutive 8UC4 frames at `src`
        if (c.src_format == PANO_SRC_YUYV) {
            // stage 0 (YUYVCAM builds, :880-886): cv::cvtColor(COLOR_YUV2BGRA_YUYV) into the 8UC4 frame m_argb
            if (h->prof_ev) cudaEventRecord(h->prof_ev[3], st);
            if (pano_frontend_convert(h, src, in_img, B.argb, nb, st)) return PANO_ERR;
            ++h->launches;
            src = B.argb;
            img_stride = (size_t)c.cam_src_width * c.cam_src_height * 4;
        }
        // stage 1: camera frame -> undist-sized 3-channel "tmp" (:903-904 / :924-927)
        const uint8_t *cur = src; int cur_c = 4; int cw = c.cam_src_width, chh = c.cam_src_height; size_t cur_img = img_stride;
        auto target = [&](bool last, uint8_t *scratch) { return last ? final_dst : scratch; };
        // the fast path reads the frames as 16-byte vectors (or through a TMA descriptor): base and image stride must be
        // 16-byte aligned; anything else takes the generic byte-wise kernels below
        if (h->fast4 && (img_stride & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const int *rc = c.rect;
            const size_t w_img = (size_t)rc[2] * rc[3];
            const dim3 cg((rc[2] + 127) / 128, (rc[3] + 7) / 8, nb);
            static const bool tex_w = getenv("PANO_CUBIC_TEX") != nullptr;
            static const bool no_tiled = getenv("PANO_CUBIC_UNTILED") != nullptr;
            static const bool no_walk = getenv("PANO_NO_RESIZE_WALK") != nullptr;
            static const bool no_tma = getenv("PANO_NO_TMA") != nullptr;        // A/B switch: LDG/STS staging loop
            const uint32_t *src4 = reinterpret_cast<const uint32_t *>(src);
            const uint4 *tab4 = reinterpret_cast<const uint4 *>(h->dtab);
            const bool tiled = h->cub_tiles && h->wtex && !no_tiled;
            const bool walk = h->r_mid.walk && !no_walk && (reinterpret_cast<uintptr_t>(final_dst) & 3) == 0 && (o_img & 3) == 0 && ((uw * 3) & 3) == 0;
            if (out_px == 4 && !walk) return ffail(h, "pano_frontend_run: word output needs the row-walking resize");
            CubicArgs ca{};
            ca.src = src4; ca.src_img_words = img_stride / 4; ca.sw = cw; ca.sh = chh; ca.map = h->dmap32; ca.map_w = uw; ca.wtex = h->wtex;
            ca.tiles = h->cub_tiles; ca.rx = rc[0]; ca.ry = rc[1]; ca.rw = rc[2]; ca.rh = rc[3]; ca.dst = B.w; ca.dst_img_words = w_img;
            ca.tx = h->cub_tx; ca.ty = h->cub_ty;
            const bool tma = tiled && h->cub_tma && !no_tma &&
                             tma_encode_words3d(&ca.tm0, src4, cw, chh, nb, (size_t)cw * 4, img_stride, kCubBoxW, kCubBoxH0) &&
                             tma_encode_words3d(&ca.tm1, src4, cw, chh, nb, (size_t)cw * 4, img_stride, kCubBoxW, kCubBoxH1);
            ResizeArgs ra{B.w, w_img, (unsigned)rc[2], final_dst, o_img, (unsigned)(uw * out_px), h->r_mid,
                          (uw + 127) / 128, (rc[3] + 1 + kResizeBand - 1) / kResizeBand};
            if (h->prof_ev) cudaEventRecord(h->prof_ev[0], st);
            if (tma)
                launch_chain(cubic5_kernel<true>, dim3(h->cub_tx, h->cub_ty, nb), blk, st, ca);
            else if (tiled)
                launch_chain(cubic5_kernel<false>, dim3(h->cub_tx, h->cub_ty, nb), blk, st, ca);
            else if (h->dmap32 && tex_w)
                cubic4_kernel<true, true><<<cg, blk, 0, st>>>(src4, img_stride / 4, cw, chh, h->dmap32, uw, tab4, h->wtex, rc[0], rc[1], rc[2], rc[3], B.w, w_img);
            else if (h->dmap32)
                cubic4_kernel<true, false><<<cg, blk, 0, st>>>(src4, img_stride / 4, cw, chh, h->dmap32, uw, tab4, h->wtex, rc[0], rc[1], rc[2], rc[3], B.w, w_img);
            else
                cubic4_kernel<false, false><<<cg, blk, 0, st>>>(src4, img_stride / 4, cw, chh, h->dmap, uw, tab4, h->wtex, rc[0], rc[1], rc[2], rc[3], B.w, w_img);
            if (h->prof_ev) cudaEventRecord(h->prof_ev[1], st);
            static const int cols_env = getenv("PANO_RESIZE_COLS") ? atoi(getenv("PANO_RESIZE_COLS")) : 4;     // A/B switch: output columns per lane
            const bool idx32 = w_img * (size_t)nb < ((size_t)1 << 32);
            if (walk && out_px == 4 && cols_env == 4 && idx32 && uw % 128 == 0 && (reinterpret_cast<uintptr_t>(final_dst) & 15) == 0 && (o_img & 15) == 0)
            {
                static const int occ = getenv("PANO_RESIZE_OCC") ? atoi(getenv("PANO_RESIZE_OCC")) : 7;      // tuning knob: min blocks per SM
                const dim3 g4((uw + 511) / 512, ra.gy, nb);
                if (occ >= 7) launch_chain(resize4_walkn_kernel<4, 7>, g4, dim3(32, 4), st, ra);
                else if (occ == 6) launch_chain(resize4_walkn_kernel<4, 6>, g4, dim3(32, 4), st, ra);
                else launch_chain(resize4_walkn_kernel<4, 5>, g4, dim3(32, 4), st, ra);
            }
            else if (walk && out_px == 4 && cols_env >= 2 && idx32 && uw % 64 == 0 && (reinterpret_cast<uintptr_t>(final_dst) & 7) == 0 && (o_img & 7) == 0)
                launch_chain(resize4_walkn_kernel<2, 8>, dim3((uw + 255) / 256, ra.gy, nb), dim3(32, 4), st, ra);
            else if (walk && out_px == 4)
                launch_chain(resize4_walk_kernel<true>, dim3(ra.gx, ra.gy, nb), dim3(32, 4), st, ra);
            else if (walk)
                launch_chain(resize4_walk_kernel<false>, dim3(ra.gx, ra.gy, nb), dim3(32, 4), st, ra);
            else
                resize4_kernel<<<dim3((uw + 127) / 128, (uh + 7) / 8, nb), blk, 0, st>>>(B.w, w_img, rc[2], final_dst, o_img,
                                                                                         uw * 3, h->r_mid);
            if (h->prof_ev) cudaEventRecord(h->prof_ev[2], st);
            h->launches += 2;
            continue;
        }
        if (out_px != 3) return ffail(h, "pano_frontend_run: word output needs 16-byte aligned camera frames");
        if (scratchGeneric(h, B, st)) return PANO_ERR;
        if (c.undistort) {
            if (h->use_r_in) {
                resize_kernel<4><<<grid2(uw, uh, blk, nb), blk, 0, st>>>(cur, cur_img, cw * 4, B.a, u_img, uw * 3, h->r_in);
                ++h->launches;
                cur = B.a; cur_c = 3; cw = uw; chh = uh; cur_img = u_img;
            }
            // stage 2: cubic remap restricted to the crop rect (:909-916)
            const int *rc = c.rect;
            const bool last2 = !h->use_r_mid && !h->use_r_out;
            uint8_t *d2 = target(last2, B.b);
            const size_t d2_img = last2 ? o_img : (size_t)rc[2] * rc[3] * 3;
            if (cur_c == 4)
                cubic_kernel<4><<<grid2(rc[2], rc[3], blk, nb), blk, 0, st>>>(cur, cur_img, cw, chh, cw * 4, h->dmap, uw, h->dtab,
                                                                              rc[0], rc[1], rc[2], rc[3], d2, d2_img);
            else
                cubic_kernel<3><<<grid2(rc[2], rc[3], blk, nb), blk, 0, st>>>(cur, cur_img, cw, chh, cw * 3, h->dmap, uw, h->dtab,
                                                                              rc[0], rc[1], rc[2], rc[3], d2, d2_img);
            ++h->launches;
            cur = d2; cur_c = 3; cw = rc[2]; chh = rc[3]; cur_img = d2_img;
            // stage 3: resize the crop back to the undistort size (:917)
            if (h->use_r_mid) {
                uint8_t *d3 = target(!h->use_r_out, B.c);
                const size_t d3_img = h->use_r_out ? u_img : o_img;
                resize_kernel<3><<<grid2(uw, uh, blk, nb), blk, 0, st>>>(cur, cur_img, cw * 3, d3, d3_img, uw * 3, h->r_mid);
                ++h->launches;
                cur = d3; cw = uw; chh = uh; cur_img = d3_img;
            }
        } else {
            uint8_t *d1 = target(!h->use_r_out, B.a);
            const size_t d1_img = h->use_r_out ? u_img : o_img;
            if (h->use_r_in) {
                resize_kernel<4><<<grid2(uw, uh, blk, nb), blk, 0, st>>>(cur, cur_img, cw * 4, d1, d1_img, uw * 3, h->r_in);
            } else {
                const size_t npx = (size_t)uw * uh;
                for (int k = 0; k < nb; ++k)
                    drop_alpha_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(cur + k * cur_img, d1 + k * d1_img, npx);
            }
            ++h->launches;
            cur = d1; cur_c = 3; cw = uw; chh = uh; cur_img = d1_img;
        }
        // stage 4: getFrame(.., src=false) resize to the stitcher input (:1094)
        if (h->use_r_out) {
            resize_kernel<3><<<grid2(c.out_width, c.out_height, blk, nb), blk, 0, st>>>(cur, cur_img, cw * 3, final_dst, o_img,
                                                                                         c.out_width * 3, h->r_out);
            ++h->launches;
        }
    }
    FCK(h, cudaGetLastError());
    return PANO_OK;
}

// The whole pixel pipeline as ONE coordinate map, walked backwards: position (x, y) in the front end's output image
// -> position in the 8UC4 camera frame (pixel-centre coordinates, double precision).  Host only; consumed by the
// fused single-gather variant of the compose (pano_set_frontend_mode).  Each cv::resize stage inverts to
// (d + 0.5) * (1 / (dsize / ssize)) - 0.5 clamped to the source (what its index/weight tables encode); the undistort
// stage is the float map itself, bilinearly interpolated between its integer nodes.
void pano_frontend_backmap(pano_frontend_handle h, double *xs, double *ys, size_t count)
{
    const pano_frontend_config &c = h->cfg;
    const int uw = c.undist_width, uh = c.undist_height;
    auto inv = [](double d, int ssize, int dsize) {
        const double s = 1.0 / ((double)dsize / ssize);
        return std::min((double)(ssize - 1), std::max(0.0, (d + 0.5) * s - 0.5));
    };
    for (size_t k = 0; k < count; ++k) {
        double x = xs[k], y = ys[k];
        if (h->use_r_out) { x = inv(x, uw, c.out_width); y = inv(y, uh, c.out_height); }
        if (c.undistort) {
            if (h->use_r_mid) { x = inv(x, c.rect[2], uw); y = inv(y, c.rect[3], uh); }
            x += c.rect[0]; y += c.rect[1];
            x = std::min((double)(uw - 1), std::max(0.0, x));
            y = std::min((double)(uh - 1), std::max(0.0, y));
            const int x0 = std::min(uw - 2, (int)x), y0 = std::min(uh - 2, (int)y);
            const double fx = x - x0, fy = y - y0;
            const float *mx = h->mapx.data() + (size_t)y0 * uw + x0, *my = h->mapy.data() + (size_t)y0 * uw + x0;
            x = (1 - fy) * ((1 - fx) * mx[0] + fx * mx[1]) + fy * ((1 - fx) * mx[uw] + fx * mx[uw + 1]);
            y = (1 - fy) * ((1 - fx) * my[0] + fx * my[1]) + fy * ((1 - fx) * my[uw] + fx * my[uw + 1]);
        }
        if (h->use_r_in) { x = inv(x, c.cam_src_width, uw); y = inv(y, c.cam_src_height, uh); }
        xs[k] = x; ys[k] = y;
    }
}

int pano_frontend_launches(pano_frontend_handle h) { return h ? h->launches : 0; }
int pano_frontend_max_batch(pano_frontend_handle h) { return h ? h->cfg.max_batch : 0; }
// fast path only, one chunk per call: record events around the two kernels; returns their algorithmic bytes per image
bool pano_frontend_set_prof(pano_frontend_handle h, cudaEvent_t *ev, double *cubic_bytes, double *resize_bytes, int out_px)
{
    h->prof_ev = ev;
    if (!h->fast4) return false;
    const pano_frontend_config &c = h->cfg;
    const double rect_px = (double)c.rect[2] * c.rect[3];
    if (cubic_bytes) *cubic_bytes = (double)c.cam_src_width * c.cam_src_height * 4 + rect_px * (h->dmap32 ? 4 : 8) + rect_px * 4;
    if (resize_bytes) *resize_bytes = rect_px * 4 + (double)c.out_width * c.out_height * out_px;
    return true;
}
void pano_frontend_sizes(pano_frontend_handle h, int *in_wh, int *out_wh)
{
    in_wh[0] = h->cfg.cam_src_width; in_wh[1] = h->cfg.cam_src_height;
    out_wh[0] = h->cfg.out_width; out_wh[1] = h->cfg.out_height;
}

extern "C" {

int pano_frontend_process_device(pano_frontend_handle h, const uint8_t *argb, uint8_t *out, int batch, void *stream)
{
    if (!h) return PANO_ERR;
    const pano_frontend_config &c = h->cfg;
    return pano_frontend_run(h, argb, (size_t)c.cam_src_width * c.cam_src_height * h->in_px(), out,
                             (size_t)c.out_width * c.out_height * 3, batch, (cudaStream_t)stream, 3, nullptr);
}

int pano_frontend_process(pano_frontend_handle h, const uint8_t *argb_host, int stride, uint8_t *out_host, int out_stride)
{
    if (!h || !argb_host || !out_host) return ffail(h, "pano_frontend_process: bad argument");
    FCK(h, cudaSetDevice(h->cfg.device));
    const pano_frontend_config &c = h->cfg;
    const size_t in_row = (size_t)c.cam_src_width * h->in_px(), out_row = (size_t)c.out_width * 3;
    if (stride < (int)in_row || out_stride < (int)out_row) return ffail(h, "stride too small");
    if (!h->stage_in) {
        if (falloc(h, &h->stage_in, in_row * c.cam_src_height) || falloc(h, &h->stage_out, out_row * c.out_height)) return PANO_ERR;
    }
    FCK(h, cudaMemcpy2D(h->stage_in, in_row, argb_host, stride, in_row, c.cam_src_height, cudaMemcpyHostToDevice));
    if (pano_frontend_process_device(h, h->stage_in, h->stage_out, 1, nullptr)) return PANO_ERR;
    FCK(h, cudaMemcpy2D(out_host, out_stride, h->stage_out, out_row, out_row, c.out_height, cudaMemcpyDeviceToHost));
    return PANO_OK;
}

}  // extern "C"
