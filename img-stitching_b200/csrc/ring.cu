// Caller step right after ocvStitcher::process in the two-ring rigs: the upper and lower ring
// panoramas are brought to one size, stacked and separated by a black bar --
//   src/master.cpp:321-326        cv::resize(up, up, down.size()); cv::vconcat(up, down, ret);
//                                 cv::rectangle(ret, Rect(0, ret.rows/2 - 5, ret.cols, 10), 0, -1);
//   src/panocamimpl.cpp:354-360   both cropped to Rect(0, finalcut, min width, min height - 2*finalcut);
//                                 cv::vconcat; cv::rectangle(ret, Rect(0, height - 2, width, 4), 0, -1);
// ONE kernel writes the stacked frame: every output byte is written once (resized / copied / bar), nothing is staged.
// The display step that follows it in the renderer, nvrenderAlpha::fit2final (src/nvrenderAlpha.cpp:153-189), is the
// second kernel here (pano_fit_*): the stacked frame scaled by fitscale = min(1, canvas_w / cols) with
// cv::resize(.., Size(), fitscale, fitscale) and pasted centred on the black 1920x1080 canvas.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/panob200.h"
#include "geometry.hpp"

using namespace pano;

namespace {

thread_local std::string g_ring_error;

struct RingArgs {
    int out_w, out_h;          // stacked frame
    int half_h;                // rows of the upper part
    int bar_y0, bar_y1;        // black rows [y0, y1)
    int crop_y;                // first source row of both parts (CROP mode), 0 otherwise
    int resize;                // 1: the upper part is cv::resize(INTER_LINEAR) of `up`
    int up_w, up_h;
    const int *xofs, *yofs;    // cv::resize tables (resize == 1)
    const short2 *xa, *ya;
};

__device__ __forceinline__ int sat_u8(int v) { return max(0, min(255, v)); }

// one thread = 4 consecutive output pixels (12 bytes, three word stores when the row is word aligned)
__global__ void __launch_bounds__(256) ring_kernel(const uint8_t *__restrict__ up, size_t up_img, int up_stride,
                                                   const uint8_t *__restrict__ down, size_t down_img, int down_stride,
                                                   uint8_t *__restrict__ out, size_t out_img, int out_stride, RingArgs a)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= a.out_w || y >= a.out_h) return;
    const int npx = min(4, a.out_w - x0);
    uint8_t px[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) px[i] = 0;
    if (y < a.bar_y0 || y >= a.bar_y1) {
        if (y < a.half_h && a.resize) {
            // cv::resize INTER_LINEAR 8UC3: horizontal pass in 11-bit coefficients, vertical pass on (t >> 4)
            const uint8_t *s = up + (size_t)blockIdx.z * up_img;
            const int yo = a.yofs[y];
            const int sy0 = min(max(yo, 0), a.up_h - 1), sy1 = min(max(yo + 1, 0), a.up_h - 1);
            const short2 ay = a.ya[y];
            const uint8_t *r0 = s + (size_t)sy0 * up_stride, *r1 = s + (size_t)sy1 * up_stride;
            for (int j = 0; j < npx; ++j) {
                const int sx0 = a.xofs[x0 + j], sx1 = min(sx0 + 1, a.up_w - 1);
                const short2 ax = a.xa[x0 + j];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int t0 = __ldg(r0 + sx0 * 3 + c) * ax.x + __ldg(r0 + sx1 * 3 + c) * ax.y;
                    const int t1 = __ldg(r1 + sx0 * 3 + c) * ax.x + __ldg(r1 + sx1 * 3 + c) * ax.y;
                    px[3 * j + c] = (uint8_t)sat_u8((((ay.x * (t0 >> 4)) >> 16) + ((ay.y * (t1 >> 4)) >> 16) + 2) >> 2);
                }
            }
        } else {
            const bool top = y < a.half_h;
            const uint8_t *s = top ? up + (size_t)blockIdx.z * up_img + (size_t)(y + a.crop_y) * up_stride
                                   : down + (size_t)blockIdx.z * down_img + (size_t)(y - a.half_h + a.crop_y) * down_stride;
            s += (size_t)x0 * 3;
            if (npx == 4 && ((reinterpret_cast<uintptr_t>(s) & 3) == 0)) {
                const uint32_t *w = reinterpret_cast<const uint32_t *>(s);
                const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
                *reinterpret_cast<uint32_t *>(px) = w0; *reinterpret_cast<uint32_t *>(px + 4) = w1; *reinterpret_cast<uint32_t *>(px + 8) = w2;
            } else {
                for (int i = 0; i < 3 * npx; ++i) px[i] = __ldg(s + i);
            }
        }
    }
    uint8_t *d = out + (size_t)blockIdx.z * out_img + (size_t)y * out_stride + (size_t)x0 * 3;
    if (npx == 4 && ((reinterpret_cast<uintptr_t>(d) & 3) == 0)) {
        uint32_t *w = reinterpret_cast<uint32_t *>(d);
        w[0] = *reinterpret_cast<uint32_t *>(px); w[1] = *reinterpret_cast<uint32_t *>(px + 4); w[2] = *reinterpret_cast<uint32_t *>(px + 8);
    } else {
        for (int i = 0; i < 3 * npx; ++i) d[i] = px[i];
    }
}

struct FitArgs {
    int canvas_w, canvas_h;    // output
    int ox, oy, w, h;          // paste rectangle inside the canvas
    int in_w, in_h;
    int mode;                  // 0 copy (fitscale == 1), 1 cv::resize INTER_LINEAR, 2 the 2x2 average cv::resize switches to at exactly 1/2
    const int *xofs, *yofs;
    const short2 *xa, *ya;
};

// one thread = one canvas pixel (the canvas is 2 Mpx: a streaming kernel far from any limit)
__global__ void __launch_bounds__(256) fit_kernel(const uint8_t *__restrict__ in, size_t in_img, int in_stride,
                                                  uint8_t *__restrict__ out, size_t out_img, int out_stride, FitArgs a)
{
    const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X >= a.canvas_w || Y >= a.canvas_h) return;
    uint8_t *d = out + (size_t)blockIdx.z * out_img + (size_t)Y * out_stride + (size_t)X * 3;
    const int x = X - a.ox, y = Y - a.oy;
    int v[3] = {0, 0, 0};                                            // canvas.setTo(0) (src/nvrenderAlpha.cpp:11)
    if ((unsigned)x < (unsigned)a.w && (unsigned)y < (unsigned)a.h) {
        const uint8_t *s = in + (size_t)blockIdx.z * in_img;
        if (a.mode == 0) {
            const uint8_t *p = s + (size_t)y * in_stride + (size_t)x * 3;
            v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
        } else if (a.mode == 2) {
            const uint8_t *p = s + (size_t)(2 * y) * in_stride + (size_t)(2 * x) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = (p[c] + p[3 + c] + p[in_stride + c] + p[in_stride + 3 + c] + 2) >> 2;
        } else {
            const int yo = a.yofs[y];
            const int sy0 = min(max(yo, 0), a.in_h - 1), sy1 = min(max(yo + 1, 0), a.in_h - 1);
            const short2 ay = a.ya[y], ax = a.xa[x];
            const int sx0 = a.xofs[x], sx1 = min(sx0 + 1, a.in_w - 1);
            const uint8_t *r0 = s + (size_t)sy0 * in_stride, *r1 = s + (size_t)sy1 * in_stride;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int t0 = __ldg(r0 + sx0 * 3 + c) * ax.x + __ldg(r0 + sx1 * 3 + c) * ax.y;
                const int t1 = __ldg(r1 + sx0 * 3 + c) * ax.x + __ldg(r1 + sx1 * 3 + c) * ax.y;
                v[c] = sat_u8((((ay.x * (t0 >> 4)) >> 16) + ((ay.y * (t1 >> 4)) >> 16) + 2) >> 2);
            }
        }
    }
    d[0] = (uint8_t)v[0]; d[1] = (uint8_t)v[1]; d[2] = (uint8_t)v[2];
}

}  // namespace

struct pano_ring_ctx {
    pano_ring_config cfg{};
    RingArgs args{};
    std::string err;
    std::vector<void *> owned;
    uint8_t *st_up = nullptr, *st_down = nullptr, *st_out = nullptr;   // staging of the host entry point
};

namespace {

int rfail(pano_ring_ctx *h, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_ring_error = buf;
    return PANO_ERR;
}

template <typename T>
bool upload(pano_ring_ctx *h, const std::vector<T> &v, const T **dst)
{
    void *p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(1, v.size()) * sizeof(T)) != cudaSuccess) return false;
    h->owned.push_back(p);
    if (cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    *dst = static_cast<const T *>(p);
    return true;
}

}  // namespace

extern "C" {

const char *pano_ring_last_error(pano_ring_handle h) { return h ? h->err.c_str() : g_ring_error.c_str(); }

int pano_ring_create(const pano_ring_config *cfg, pano_ring_handle *out)
{
    if (!cfg || !out) return rfail(nullptr, "pano_ring_create: null argument");
    *out = nullptr;
    if (cfg->up_width < 1 || cfg->up_height < 1 || cfg->down_width < 1 || cfg->down_height < 1 || cfg->bar < 0)
        return rfail(nullptr, "pano_ring_create: bad sizes");
    if (cfg->mode != PANO_RING_RESIZE && cfg->mode != PANO_RING_CROP) return rfail(nullptr, "pano_ring_create: bad mode");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return rfail(nullptr, "no CUDA device: this library has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return rfail(nullptr, "bad device ordinal");
    pano_ring_ctx *h = new pano_ring_ctx();
    h->cfg = *cfg;
    auto bail = [&](const char *msg) { g_ring_error = msg; pano_ring_destroy(h); return PANO_ERR; };
    if (cudaSetDevice(cfg->device) != cudaSuccess) return bail("cudaSetDevice failed");
    RingArgs &a = h->args;
    a.up_w = cfg->up_width; a.up_h = cfg->up_height;
    if (cfg->mode == PANO_RING_RESIZE) {
        a.out_w = cfg->down_width; a.half_h = cfg->down_height; a.out_h = 2 * cfg->down_height; a.crop_y = 0;
        a.resize = !(cfg->up_width == cfg->down_width && cfg->up_height == cfg->down_height);
        a.bar_y0 = a.out_h / 2 - cfg->bar / 2;                       // Rect(0, rows/2 - 5, cols, 10)
        a.bar_y1 = a.bar_y0 + cfg->bar;
        if (a.resize) {
            std::vector<int> xo, yo;
            std::vector<int16_t> xa0, xa1, ya0, ya1;
            resizeAxis(cfg->up_width, a.out_w, true, xo, xa0, xa1);
            resizeAxis(cfg->up_height, a.half_h, false, yo, ya0, ya1);
            std::vector<short2> xa(xo.size()), ya(yo.size());
            for (size_t i = 0; i < xo.size(); ++i) xa[i] = make_short2(xa0[i], xa1[i]);
            for (size_t i = 0; i < yo.size(); ++i) ya[i] = make_short2(ya0[i], ya1[i]);
            if (!upload(h, xo, &a.xofs) || !upload(h, yo, &a.yofs) || !upload(h, xa, &a.xa) || !upload(h, ya, &a.ya))
                return bail("ring table upload failed");
        }
    } else {
        const int width = std::min(cfg->up_width, cfg->down_width);
        const int height = std::min(cfg->up_height, cfg->down_height) - 2 * cfg->finalcut;
        if (cfg->finalcut < 0 || height < 1) return bail("pano_ring_create: finalcut leaves no rows");
        a.out_w = width; a.half_h = height; a.out_h = 2 * height; a.crop_y = cfg->finalcut; a.resize = 0;
        a.bar_y0 = height - cfg->bar / 2;                            // Rect(0, height - 2, width, 4)
        a.bar_y1 = a.bar_y0 + cfg->bar;
    }
    a.bar_y0 = std::max(0, a.bar_y0);                                // cv::rectangle clips to the image
    a.bar_y1 = std::min(a.out_h, a.bar_y1);
    *out = h;
    return PANO_OK;
}

int pano_ring_destroy(pano_ring_handle h)
{
    if (!h) return PANO_OK;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (void *p : h->owned) cudaFree(p);
    delete h;
    return PANO_OK;
}

int pano_ring_out_size(pano_ring_handle h, int *wh)
{
    if (!h || !wh) return PANO_ERR;
    wh[0] = h->args.out_w; wh[1] = h->args.out_h;
    return PANO_OK;
}

int pano_ring_compose_device(pano_ring_handle h, const uint8_t *up_dev, int up_stride, const uint8_t *down_dev, int down_stride,
                             uint8_t *out_dev, int out_stride, int batch, void *stream)
{
    if (!h || !up_dev || !down_dev || !out_dev || batch < 1) return rfail(h, "pano_ring_compose_device: bad argument");
    const pano_ring_config &c = h->cfg;
    const RingArgs &a = h->args;
    if (up_stride < 3 * c.up_width || down_stride < 3 * c.down_width || out_stride < 3 * a.out_w)
        return rfail(h, "pano_ring_compose_device: stride too small");
    if (cudaSetDevice(c.device) != cudaSuccess) return rfail(h, "cudaSetDevice failed");
    const dim3 block(64, 4), grid(((a.out_w + 3) / 4 + 63) / 64, (a.out_h + 3) / 4, batch);
    ring_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(up_dev, (size_t)up_stride * c.up_height, up_stride, down_dev,
                                                          (size_t)down_stride * c.down_height, down_stride, out_dev,
                                                          (size_t)out_stride * a.out_h, out_stride, a);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return rfail(h, "ring_kernel launch failed: %s", cudaGetErrorString(e));
    return PANO_OK;
}

int pano_ring_compose(pano_ring_handle h, const uint8_t *up_host, int up_stride, const uint8_t *down_host, int down_stride,
                      uint8_t *out_host, int out_stride)
{
    if (!h || !up_host || !down_host || !out_host) return rfail(h, "pano_ring_compose: bad argument");
    const pano_ring_config &c = h->cfg;
    const RingArgs &a = h->args;
    if (up_stride < 3 * c.up_width || down_stride < 3 * c.down_width || out_stride < 3 * a.out_w)
        return rfail(h, "pano_ring_compose: stride too small");
    if (cudaSetDevice(c.device) != cudaSuccess) return rfail(h, "cudaSetDevice failed");
    const size_t ur = (size_t)3 * c.up_width, dr = (size_t)3 * c.down_width, orow = (size_t)3 * a.out_w;
    if (!h->st_up) {
        void *p[3] = {nullptr, nullptr, nullptr};
        if (cudaMalloc(&p[0], ur * c.up_height) != cudaSuccess || cudaMalloc(&p[1], dr * c.down_height) != cudaSuccess ||
            cudaMalloc(&p[2], orow * a.out_h) != cudaSuccess) {
            for (void *q : p) cudaFree(q);
            return rfail(h, "staging allocation failed");
        }
        for (void *q : p) h->owned.push_back(q);
        h->st_up = (uint8_t *)p[0]; h->st_down = (uint8_t *)p[1]; h->st_out = (uint8_t *)p[2];
    }
    if (cudaMemcpy2D(h->st_up, ur, up_host, up_stride, ur, c.up_height, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy2D(h->st_down, dr, down_host, down_stride, dr, c.down_height, cudaMemcpyHostToDevice) != cudaSuccess)
        return rfail(h, "H2D copy failed");
    if (pano_ring_compose_device(h, h->st_up, (int)ur, h->st_down, (int)dr, h->st_out, (int)orow, 1, nullptr)) return PANO_ERR;
    if (cudaMemcpy2D(out_host, out_stride, h->st_out, orow, orow, a.out_h, cudaMemcpyDeviceToHost) != cudaSuccess)
        return rfail(h, "D2H copy failed");
    return PANO_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- fit2final
struct pano_fit_ctx {
    pano_fit_config cfg{};
    FitArgs args{};
    double fitscale = 1.0;
    std::string err;
    std::vector<void *> owned;
    uint8_t *st_in = nullptr, *st_out = nullptr;
};

namespace {
thread_local std::string g_fit_error;
int ffail2(pano_fit_ctx *h, const char *msg) { if (h) h->err = msg; else g_fit_error = msg; return PANO_ERR; }
template <typename T>
bool fupload(pano_fit_ctx *h, const std::vector<T> &v, const T **dst)
{
    void *p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(1, v.size()) * sizeof(T)) != cudaSuccess) return false;
    h->owned.push_back(p);
    if (cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    *dst = static_cast<const T *>(p);
    return true;
}
}  // namespace

extern "C" {

const char *pano_fit_last_error(pano_fit_handle h) { return h ? h->err.c_str() : g_fit_error.c_str(); }

int pano_fit_create(const pano_fit_config *cfg, pano_fit_handle *out)
{
    if (!cfg || !out) return ffail2(nullptr, "pano_fit_create: null argument");
    *out = nullptr;
    if (cfg->in_width < 1 || cfg->in_height < 1 || cfg->canvas_width < 1 || cfg->canvas_height < 1) return ffail2(nullptr, "pano_fit_create: bad sizes");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ffail2(nullptr, "no CUDA device: this library has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return ffail2(nullptr, "bad device ordinal");
    pano_fit_ctx *h = new pano_fit_ctx();
    h->cfg = *cfg;
    auto bail = [&](const char *msg) { g_fit_error = msg; pano_fit_destroy(h); return PANO_ERR; };
    if (cudaSetDevice(cfg->device) != cudaSuccess) return bail("cudaSetDevice failed");
    FitArgs &a = h->args;
    a.canvas_w = cfg->canvas_width; a.canvas_h = cfg->canvas_height; a.in_w = cfg->in_width; a.in_h = cfg->in_height;
    if (cfg->in_width == cfg->canvas_width && cfg->in_height == cfg->canvas_height) {
        a.ox = a.oy = 0; a.w = a.canvas_w; a.h = a.canvas_h; a.mode = 0;          // input.copyTo(canvas)  (:156-160)
        *out = h;
        return PANO_OK;
    }
    // :166-181 -- fitscale only shrinks, and only by the width
    h->fitscale = cfg->in_width > cfg->canvas_width ? cfg->canvas_width * 1.0 / cfg->in_width : 1.0;
    a.w = (int)std::nearbyint(cfg->in_width * h->fitscale);          // cv::resize: dsize = saturate_cast<int>(ssize * fx)
    a.h = (int)std::nearbyint(cfg->in_height * h->fitscale);
    a.ox = (cfg->canvas_width - a.w) / 2;
    a.oy = (cfg->canvas_height - a.h) / 2;
    if (a.w < 1 || a.h < 1 || a.ox < 0 || a.oy < 0 || a.ox + a.w > a.canvas_w || a.oy + a.h > a.canvas_h)
        return bail("pano_fit_create: the scaled frame does not fit the canvas (cv::Mat::operator()(Rect) would assert, :187)");
    if (a.w == cfg->in_width && a.h == cfg->in_height) {
        a.mode = 0;                                                   // cv::resize with equal sizes is a copy
    } else {
        const double sx = 1.0 / h->fitscale;
        const int isx = (int)std::nearbyint(sx);
        const bool area_fast = std::abs(sx - isx) < 2.220446049250313e-16 && isx == 2 && a.w * 2 == cfg->in_width && a.h * 2 == cfg->in_height;
        if (area_fast) {
            a.mode = 2;                                               // INTER_LINEAR at exactly 1/2 runs as the 2x2 INTER_AREA average
        } else {
            a.mode = 1;
            std::vector<int> xo, yo;
            std::vector<int16_t> xa0, xa1, ya0, ya1;
            resizeAxis(cfg->in_width, a.w, true, xo, xa0, xa1, h->fitscale);
            resizeAxis(cfg->in_height, a.h, false, yo, ya0, ya1, h->fitscale);
            std::vector<short2> xa(xo.size()), ya(yo.size());
            for (size_t i = 0; i < xo.size(); ++i) xa[i] = make_short2(xa0[i], xa1[i]);
            for (size_t i = 0; i < yo.size(); ++i) ya[i] = make_short2(ya0[i], ya1[i]);
            if (!fupload(h, xo, &a.xofs) || !fupload(h, yo, &a.yofs) || !fupload(h, xa, &a.xa) || !fupload(h, ya, &a.ya))
                return bail("fit table upload failed");
        }
    }
    *out = h;
    return PANO_OK;
}

int pano_fit_destroy(pano_fit_handle h)
{
    if (!h) return PANO_OK;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (void *p : h->owned) cudaFree(p);
    delete h;
    return PANO_OK;
}

int pano_fit_geometry(pano_fit_handle h, int *rect, double *fitscale)
{
    if (!h) return PANO_ERR;
    if (rect) { rect[0] = h->args.ox; rect[1] = h->args.oy; rect[2] = h->args.w; rect[3] = h->args.h; }
    if (fitscale) *fitscale = h->fitscale;
    return PANO_OK;
}

int pano_fit_compose_device(pano_fit_handle h, const uint8_t *in_dev, int in_stride, uint8_t *canvas_dev, int canvas_stride, int batch, void *stream)
{
    if (!h || !in_dev || !canvas_dev || batch < 1) return ffail2(h, "pano_fit_compose_device: bad argument");
    const FitArgs &a = h->args;
    if (in_stride < 3 * a.in_w || canvas_stride < 3 * a.canvas_w) return ffail2(h, "pano_fit_compose_device: stride too small");
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) return ffail2(h, "cudaSetDevice failed");
    const dim3 block(64, 4), grid((a.canvas_w + 63) / 64, (a.canvas_h + 3) / 4, batch);
    fit_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in_dev, (size_t)in_stride * a.in_h, in_stride, canvas_dev,
                                                         (size_t)canvas_stride * a.canvas_h, canvas_stride, a);
    if (cudaGetLastError() != cudaSuccess) return ffail2(h, "fit_kernel launch failed");
    return PANO_OK;
}

int pano_fit_compose(pano_fit_handle h, const uint8_t *in_host, int in_stride, uint8_t *canvas_host, int canvas_stride)
{
    if (!h || !in_host || !canvas_host) return ffail2(h, "pano_fit_compose: bad argument");
    const FitArgs &a = h->args;
    if (in_stride < 3 * a.in_w || canvas_stride < 3 * a.canvas_w) return ffail2(h, "pano_fit_compose: stride too small");
    if (cudaSetDevice(h->cfg.device) != cudaSuccess) return ffail2(h, "cudaSetDevice failed");
    const size_t ir = (size_t)3 * a.in_w, orow = (size_t)3 * a.canvas_w;
    if (!h->st_in) {
        void *p[2] = {nullptr, nullptr};
        if (cudaMalloc(&p[0], ir * a.in_h) != cudaSuccess || cudaMalloc(&p[1], orow * a.canvas_h) != cudaSuccess) {
            for (void *q : p) cudaFree(q);
            return ffail2(h, "staging allocation failed");
        }
        for (void *q : p) h->owned.push_back(q);
        h->st_in = (uint8_t *)p[0]; h->st_out = (uint8_t *)p[1];
    }
    if (cudaMemcpy2D(h->st_in, ir, in_host, in_stride, ir, a.in_h, cudaMemcpyHostToDevice) != cudaSuccess) return ffail2(h, "H2D copy failed");
    if (pano_fit_compose_device(h, h->st_in, (int)ir, h->st_out, (int)orow, 1, nullptr)) return PANO_ERR;
    if (cudaMemcpy2D(canvas_host, canvas_stride, h->st_out, orow, orow, a.canvas_h, cudaMemcpyDeviceToHost) != cudaSuccess)
        return ffail2(h, "D2H copy failed");
    return PANO_OK;
}

}  // extern "C"
