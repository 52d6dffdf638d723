// Caller step right after ocvStitcher::process in the two-ring rigs: the upper and lower ring
// panoramas are brought to one size, stacked and separated by a black bar --
//   src/master.cpp:321-326        cv::resize(up, up, down.size()); cv::vconcat(up, down, ret);
//                                 cv::rectangle(ret, Rect(0, ret.rows/2 - 5, ret.cols, 10), 0, -1);
//   src/panocamimpl.cpp:354-360   both cropped to Rect(0, finalcut, min width, min height - 2*finalcut);
//                                 cv::vconcat; cv::rectangle(ret, Rect(0, height - 2, width, 4), 0, -1);
// ONE kernel writes the stacked frame: every output byte is written once (resized / copied / bar), nothing is staged.
#include <cuda_runtime.h>

#include <algorithm>
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
