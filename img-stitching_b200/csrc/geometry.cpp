// Host-side init-time table builders.  Compile with -ffp-contract=off (see Makefile).
#include "geometry.hpp"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <emmintrin.h>

namespace pano {

namespace {

// round-half-even with x86 "integer indefinite" on overflow, as cvRound does
inline int roundEven(float v) { return _mm_cvtss_si32(_mm_set_ss(v)); }
inline int clampS16(int v) { return std::max(-32768, std::min(32767, v)); }

// 3x3 float product with float accumulation, left to right (cv::gemm on CV_32F 3x3)
void mul3(const float *a, const float *b, float *c)
{
    for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) {
            float acc = a[3 * r] * b[q];
            acc += a[3 * r + 1] * b[3 + q];
            acc += a[3 * r + 2] * b[6 + q];
            c[3 * r + q] = acc;
        }
}

// cv::invert for 3x3 CV_32F: cofactor formula evaluated in double
void inv3(const float *m, float *out)
{
    auto M = [&](int r, int c) { return static_cast<double>(m[3 * r + c]); };
    double det = M(0, 0) * (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1)) -
                 M(0, 1) * (M(1, 0) * M(2, 2) - M(1, 2) * M(2, 0)) +
                 M(0, 2) * (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0));
    if (det == 0.0) {
        std::fill(out, out + 9, 0.f);
        return;
    }
    det = 1.0 / det;
    const double cof[9] = {
        (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1)) * det, (M(0, 2) * M(2, 1) - M(0, 1) * M(2, 2)) * det,
        (M(0, 1) * M(1, 2) - M(0, 2) * M(1, 1)) * det, (M(1, 2) * M(2, 0) - M(1, 0) * M(2, 2)) * det,
        (M(0, 0) * M(2, 2) - M(0, 2) * M(2, 0)) * det, (M(0, 2) * M(1, 0) - M(0, 0) * M(1, 2)) * det,
        (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0)) * det, (M(0, 1) * M(2, 0) - M(0, 0) * M(2, 1)) * det,
        (M(0, 0) * M(1, 1) - M(0, 1) * M(1, 0)) * det};
    for (int i = 0; i < 9; ++i) out[i] = static_cast<float>(cof[i]);
}

inline int reflect101(int p, int n)
{
    if (static_cast<unsigned>(p) < static_cast<unsigned>(n)) return p;
    if (n == 1) return 0;
    do {
        p = p < 0 ? -p : 2 * n - 2 - p;
    } while (static_cast<unsigned>(p) >= static_cast<unsigned>(n));
    return p;
}

}  // namespace

// ------------------------------------------------------------------------- warper

void RotationWarper::setCamera(const float K[9], const float R[9])
{
    float kinv[9];
    std::copy(K, K + 9, k_);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) rinv_[3 * r + c] = R[3 * c + r];
    inv3(K, kinv);
    mul3(R, kinv, r_kinv_);
    mul3(K, rinv_, k_rinv_);
}

void RotationWarper::forward(float x, float y, float &u, float &v) const
{
    const float *m = r_kinv_;
    const float px = m[0] * x + m[1] * y + m[2];
    const float py = m[3] * x + m[4] * y + m[5];
    const float pz = m[6] * x + m[7] * y + m[8];
    u = scale_ * atan2f(px, pz);
    if (kind_ == 0) {
        const float w = py / sqrtf(px * px + py * py + pz * pz);
        v = scale_ * (static_cast<float>(M_PI) - acosf(w == w ? w : 0));
    } else {
        v = scale_ * py / sqrtf(px * px + pz * pz);
    }
}

void RotationWarper::backward(float u, float v, float &x, float &y) const
{
    u /= scale_;
    v /= scale_;
    float px, py, pz;
    if (kind_ == 0) {
        const float sinv = sinf(static_cast<float>(M_PI) - v);
        px = sinv * sinf(u);
        py = cosf(static_cast<float>(M_PI) - v);
        pz = sinv * cosf(u);
    } else {
        px = sinf(u);
        py = v;
        pz = cosf(u);
    }
    const float *m = k_rinv_;
    x = m[0] * px + m[1] * py + m[2] * pz;
    y = m[3] * px + m[4] * py + m[5] * pz;
    const float z = m[6] * px + m[7] * py + m[8] * pz;
    if (z > 0) {
        x /= z;
        y /= z;
    } else {
        x = y = -1;
    }
}

Rect RotationWarper::warpRoi(int W, int H) const
{
    float lo_u = FLT_MAX, lo_v = FLT_MAX, hi_u = -FLT_MAX, hi_v = -FLT_MAX;
    auto visit = [&](float x, float y) {
        float u, v;
        forward(x, y, u, v);
        lo_u = std::min(lo_u, u); lo_v = std::min(lo_v, v);
        hi_u = std::max(hi_u, u); hi_v = std::max(hi_v, v);
    };
    for (float x = 0; x < W; ++x) {
        visit(x, 0.f);
        visit(x, static_cast<float>(H - 1));
    }
    for (int y = 0; y < H; ++y) {
        visit(0.f, static_cast<float>(y));
        visit(static_cast<float>(W - 1), static_cast<float>(y));
    }
    int tlx = static_cast<int>(lo_u), tly = static_cast<int>(lo_v);
    int brx = static_cast<int>(hi_u), bry = static_cast<int>(hi_v);
    if (kind_ == 0) {
        // the sphere's poles may fall inside the image (SphericalWarper::detectResultRoi)
        float flx = static_cast<float>(tlx), fly = static_cast<float>(tly);
        float fhx = static_cast<float>(brx), fhy = static_cast<float>(bry);
        for (int pole = 0; pole < 2; ++pole) {
            const float x = rinv_[1], y = pole ? -rinv_[4] : rinv_[4], z = rinv_[7];
            if (!(y > 0.f)) continue;
            const float ix = (k_[0] * x + k_[1] * y) / z + k_[2];
            const float iy = k_[4] * y / z + k_[5];
            if (ix > 0.f && ix < W && iy > 0.f && iy < H) {
                const float pv = pole ? 0.f : static_cast<float>(M_PI * scale_);
                flx = std::min(flx, 0.f); fly = std::min(fly, pv);
                fhx = std::max(fhx, 0.f); fhy = std::max(fhy, pv);
            }
        }
        tlx = static_cast<int>(flx); tly = static_cast<int>(fly);
        brx = static_cast<int>(fhx); bry = static_cast<int>(fhy);
    }
    Rect r;
    r.x = tlx; r.y = tly; r.w = brx - tlx + 1; r.h = bry - tly + 1;
    return r;
}

void RotationWarper::buildMaps(int, int, const Rect &roi, float *xmap, float *ymap) const
{
    for (int v = 0; v < roi.h; ++v)
        for (int u = 0; u < roi.w; ++u) {
            float x, y;
            backward(static_cast<float>(roi.x + u), static_cast<float>(roi.y + v), x, y);
            xmap[static_cast<size_t>(v) * roi.w + u] = x;
            ymap[static_cast<size_t>(v) * roi.w + u] = y;
        }
}

FixedCoord toFixed(float mx, float my)
{
    const int sx = roundEven(mx * 32.f), sy = roundEven(my * 32.f);
    FixedCoord f;
    f.ix = clampS16(sx >> 5);
    f.iy = clampS16(sy >> 5);
    f.fx = sx & 31;
    f.fy = sy & 31;
    return f;
}

uint32_t foldReflect(int i, int f, int n)
{
    // taps of cv::remap are reflect(i), reflect(i+1) with weights (32-f), f.
    // BORDER_REFLECT maps p -> p mod 2n mirrored; walk both taps then re-express.
    auto refl = [n](int p) {
        if (n == 1) return 0;
        while (static_cast<unsigned>(p) >= static_cast<unsigned>(n)) p = p < 0 ? -p - 1 : 2 * n - 1 - p;
        return p;
    };
    const int a = refl(i), b = refl(i + 1);
    if (b == a + 1) return static_cast<uint32_t>(32 * a + f);          // ordinary orientation
    if (b == a - 1) {                                                   // mirrored segment
        if (f == 0) return static_cast<uint32_t>(32 * a);
        return static_cast<uint32_t>(32 * b + (32 - f));
    }
    return static_cast<uint32_t>(32 * a);                               // turning point: a == b
}

// ------------------------------------------------------------------------- blender geometry

Rect resultRoi(const std::vector<Rect> &rois)
{
    int x0 = INT_MAX, y0 = INT_MAX, x1 = INT_MIN, y1 = INT_MIN;
    for (const Rect &r : rois) {
        x0 = std::min(x0, r.x); y0 = std::min(y0, r.y);
        x1 = std::max(x1, r.x + r.w); y1 = std::max(y1, r.y + r.h);
    }
    Rect o;
    o.x = x0; o.y = y0; o.w = x1 - x0; o.h = y1 - y0;
    return o;
}

int multibandPrepare(const Rect &roi, int num_bands, int &pw, int &ph)
{
    const double max_len = static_cast<double>(std::max(roi.w, roi.h));
    const int nb = std::min(num_bands, static_cast<int>(std::ceil(std::log(max_len) / std::log(2.0))));
    const int unit = 1 << nb;
    pw = roi.w + (unit - roi.w % unit) % unit;
    ph = roi.h + (unit - roi.h % unit) % unit;
    return nb;
}

FeedRect multibandFeedRect(const Rect &roi, int pw, int ph, int nb, const Rect &img)
{
    const int unit = 1 << nb, gap = 3 * unit;
    const int dbrx = roi.x + pw, dbry = roi.y + ph;
    int tlx = std::max(roi.x, img.x - gap), tly = std::max(roi.y, img.y - gap);
    int brx = std::min(dbrx, img.x + img.w + gap), bry = std::min(dbry, img.y + img.h + gap);
    tlx = roi.x + (((tlx - roi.x) >> nb) << nb);
    tly = roi.y + (((tly - roi.y) >> nb) << nb);
    int w = brx - tlx, h = bry - tly;
    w += (unit - w % unit) % unit;
    h += (unit - h % unit) % unit;
    brx = tlx + w; bry = tly + h;
    const int dx = std::max(brx - dbrx, 0), dy = std::max(bry - dbry, 0);
    tlx -= dx; brx -= dx; tly -= dy; bry -= dy;
    FeedRect fr;
    fr.top = img.y - tly;
    fr.left = img.x - tlx;
    fr.bottom = bry - img.y - img.h;
    fr.right = brx - img.x - img.w;
    fr.rect.x = tlx - roi.x; fr.rect.y = tly - roi.y; fr.rect.w = w; fr.rect.h = h;
    return fr;
}

// ------------------------------------------------------------------------- weights

// cv::pyrDown on CV_32F with the evaluation order of OpenCV 4.x's universal-intrinsics build with 4-lane float vectors
// (x86-64 SSE baseline, NEON): modules/imgproc/src/pyramids.cpp evaluates the SAME 1-4-6-4-1 sums in two different
// orders depending on the column -- the vector bodies (PyrDownVecH / PyrDownVecV: c*6 + ((b+d)*4 + (a+e)) and
// (r1+r3+r2)*4 + (r0+r4+(r2+r2))) and the scalar loops around them (c*6 + (b+d)*4 + a + e, left to right) -- and
// float addition is not associative, so bit-exact weights need the column rule as well:
//   horizontal: dst column 0 and every column from hv_end on are scalar (left border, vector tail, right border), the
//               vector body runs in groups of 4 from column 1 while a whole group fits below
//               width0 = min((sw - 3) / 2 + 1, dw);
//   vertical:   groups of 4 from column 0 while a whole group fits below dw, scalar tail.
// Pinned against cv2 4.13 over all sizes from 1 x 1 to 48 x 48 and the bench shapes (tests/test_host_tables.py).
void pyrDownColumnRule(int sw, int *hv_end, int *vv_end)
{
    const int dw = (sw + 1) / 2;
    const int width0 = std::min((sw - 3) / 2 + 1, dw);
    const int groups = width0 >= 5 ? (width0 - 5) / 4 + 1 : 0;
    *hv_end = 1 + 4 * groups;
    *vv_end = dw & ~3;
}

void pyrDownF32(const float *src, int sw, int sh, float *dst)
{
    const int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    int hv_end, vv_end;
    pyrDownColumnRule(sw, &hv_end, &vv_end);
    std::vector<float> hrow(static_cast<size_t>(dw) * sh);
    for (int y = 0; y < sh; ++y) {
        const float *s = src + static_cast<size_t>(y) * sw;
        for (int x = 0; x < dw; ++x) {
            const float a = s[reflect101(2 * x - 2, sw)], b = s[reflect101(2 * x - 1, sw)], c = s[2 * x],
                        d = s[reflect101(2 * x + 1, sw)], e = s[reflect101(2 * x + 2, sw)];
            float &h = hrow[static_cast<size_t>(y) * dw + x];
            if (x >= 1 && x < hv_end) h = c * 6 + ((b + d) * 4 + (a + e));
            else h = c * 6 + (b + d) * 4 + a + e;
        }
    }
    for (int y = 0; y < dh; ++y) {
        const float *r0 = &hrow[static_cast<size_t>(reflect101(2 * y - 2, sh)) * dw];
        const float *r1 = &hrow[static_cast<size_t>(reflect101(2 * y - 1, sh)) * dw];
        const float *r2 = &hrow[static_cast<size_t>(2 * y) * dw];
        const float *r3 = &hrow[static_cast<size_t>(reflect101(2 * y + 1, sh)) * dw];
        const float *r4 = &hrow[static_cast<size_t>(reflect101(2 * y + 2, sh)) * dw];
        float *o = dst + static_cast<size_t>(y) * dw;
        for (int x = 0; x < dw; ++x) {
            if (x < vv_end) o[x] = ((r1[x] + r3[x] + r2[x]) * 4 + (r0[x] + r4[x] + (r2[x] + r2[x]))) * (1.f / 256.f);
            else o[x] = (r2[x] * 6 + (r1[x] + r3[x]) * 4 + r0[x] + r4[x]) * (1.f / 256.f);
        }
    }
}

void featherWeight(const uint8_t *mask, int w, int h, int stride, float sharpness, float *out)
{
    // two-pass 3x3 chamfer with (1, 2) costs == exact city-block distance to the nearest zero
    const int BIG = INT_MAX >> 2;
    const int pw = w + 2;
    std::vector<int> d(static_cast<size_t>(pw) * (h + 2), BIG);
    auto at = [&](int x, int y) -> int & { return d[static_cast<size_t>(y + 1) * pw + x + 1]; };
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            if (mask[static_cast<size_t>(y) * stride + x] == 0) { at(x, y) = 0; continue; }
            int t = std::min(std::min(at(x - 1, y - 1) + 2, at(x, y - 1) + 1),
                             std::min(at(x + 1, y - 1) + 2, at(x - 1, y) + 1));
            at(x, y) = std::min(t, BIG);
        }
    for (int y = h - 1; y >= 0; --y)
        for (int x = w - 1; x >= 0; --x) {
            int t = at(x, y);
            if (t > 1) {
                t = std::min(t, std::min(std::min(at(x + 1, y + 1) + 2, at(x, y + 1) + 1),
                                         std::min(at(x - 1, y + 1) + 2, at(x + 1, y) + 1)));
                at(x, y) = t;
            }
            const float dist = t >= BIG / 2 ? FLT_MAX : static_cast<float>(t);
            const float v = dist * sharpness;
            out[static_cast<size_t>(y) * w + x] = v > 1.f ? 1.f : v;
        }
}

// ------------------------------------------------------------------------- front end tables

void undistortMaps(const double K[9], const double D[4], const double A[9], int w, int h, float *mapx, float *mapy)
{
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3];
    // inverse of the new camera matrix (cofactor form, double)
    double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
                 A[2] * (A[3] * A[7] - A[4] * A[6]);
    det = 1.0 / det;
    const double ir[9] = {(A[4] * A[8] - A[5] * A[7]) * det, (A[2] * A[7] - A[1] * A[8]) * det,
                          (A[1] * A[5] - A[2] * A[4]) * det, (A[5] * A[6] - A[3] * A[8]) * det,
                          (A[0] * A[8] - A[2] * A[6]) * det, (A[2] * A[3] - A[0] * A[5]) * det,
                          (A[3] * A[7] - A[4] * A[6]) * det, (A[1] * A[6] - A[0] * A[7]) * det,
                          (A[0] * A[4] - A[1] * A[3]) * det};
    for (int v = 0; v < h; ++v) {
        double X = v * ir[1] + ir[2], Y = v * ir[4] + ir[5], Wh = v * ir[7] + ir[8];
        for (int u = 0; u < w; ++u, X += ir[0], Y += ir[3], Wh += ir[6]) {
            const double iw = 1.0 / Wh, x = X * iw, y = Y * iw;
            const double x2 = x * x, y2 = y * y, r2 = x2 + y2, xy2 = 2 * x * y;
            const double radial = 1 + (k2 * r2 + k1) * r2;
            const double xd = x * radial + p1 * xy2 + p2 * (r2 + 2 * x2);
            const double yd = y * radial + p1 * (r2 + 2 * y2) + p2 * xy2;
            mapx[static_cast<size_t>(v) * w + u] = static_cast<float>(fx * xd + cx);
            mapy[static_cast<size_t>(v) * w + u] = static_cast<float>(fy * yd + cy);
        }
    }
}

void cubicTable(int16_t *tab)
{
    const float A = -0.75f;
    float c1[32][4];
    for (int i = 0; i < 32; ++i) {
        const float t = static_cast<float>(i) * (1.f / 32.f);
        c1[i][0] = ((A * (t + 1) - 5 * A) * (t + 1) + 8 * A) * (t + 1) - 4 * A;
        c1[i][1] = ((A + 2) * t - (A + 3)) * t * t + 1;
        c1[i][2] = ((A + 2) * (1 - t) - (A + 3)) * (1 - t) * (1 - t) + 1;
        c1[i][3] = 1.f - c1[i][0] - c1[i][1] - c1[i][2];
    }
    for (int fy = 0; fy < 32; ++fy)
        for (int fx = 0; fx < 32; ++fx) {
            int16_t *t = tab + (fy * 32 + fx) * 16;
            int sum = 0;
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) {
                    const int iv = clampS16(roundEven(c1[fy][r] * c1[fx][c] * 32768.f));
                    t[4 * r + c] = static_cast<int16_t>(iv);
                    sum += iv;
                }
            if (sum == 32768) continue;
            // OpenCV spreads the rounding residue onto one of the four central taps
            int lo = 10, hi = 10;  // (2,2)
            const int cand[4] = {10, 11, 14, 15};
            for (int q = 0; q < 4; ++q) {
                const int k = cand[q];
                if (t[k] < t[lo]) lo = k;
                else if (t[k] > t[hi]) hi = k;
            }
            const int diff = sum - 32768;
            if (diff < 0) t[hi] = static_cast<int16_t>(t[hi] - diff);
            else t[lo] = static_cast<int16_t>(t[lo] - diff);
        }
}

void resizeAxis(int ssize, int dsize, bool clamp_frac, std::vector<int> &ofs,
                std::vector<int16_t> &a0, std::vector<int16_t> &a1, double inv_scale_in)
{
    ofs.resize(dsize); a0.resize(dsize); a1.resize(dsize);
    // cv::resize called with dsize derives inv_scale from the sizes; called with fx / fy it keeps the caller's factor
    const double inv_scale = inv_scale_in > 0 ? inv_scale_in : static_cast<double>(dsize) / ssize;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = static_cast<float>((d + 0.5) * scale - 0.5);
        int s = static_cast<int>(std::floor(f));
        f -= s;
        if (clamp_frac) {
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        }
        ofs[d] = s;
        a0[d] = static_cast<int16_t>(clampS16(roundEven((1.f - f) * 2048.f)));
        a1[d] = static_cast<int16_t>(clampS16(roundEven(f * 2048.f)));
    }
}

// OpenCV resize.cpp, interpolationLinear<uchar>::getCoeffs: softdouble arithmetic == IEEE double operations;
// ufixedpoint16(x) = cvRound(x * 256) (round half to even); outside the source range the edge sample is replicated.
// inv_scale <= 0: cv::resize was given dsize (inv_scale = dsize / ssize); otherwise the caller's fx / fy is used as is.
void linearExactAxis(int ssize, int dsize, std::vector<int> &ofs, std::vector<int> &c1, double inv_scale_in)
{
    ofs.assign(dsize, 0); c1.assign(dsize, 0);
    const double inv_scale = inv_scale_in > 0 ? inv_scale_in : static_cast<double>(dsize) / ssize;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        const double fval = scale * (d + 0.5) - 0.5;
        const int ival = static_cast<int>(std::floor(fval));
        if (ival >= 0 && ssize > 1) {
            if (ival < ssize - 1) {
                ofs[d] = ival;
                c1[d] = static_cast<int>(std::nearbyint((fval - ival) * 256.0));
            } else {
                ofs[d] = ssize - 1;
            }
        }
    }
}

// ------------------------------------------------------------------------- seam-finder inputs (init-time, host)
// cv::resize(src, dst, Size(), fx, fy, INTER_LINEAR_EXACT) on 8-bit images (include/ocvstitcher.hpp:988, 1228): the
// horizontal pass accumulates in 8.8 fixed point, the vertical pass in 16.16, the result is rounded half up.
void resizeLinearExactU8(const uint8_t *src, int w, int h, int stride, int ch, double fx, double fy, int dw, int dh,
                         uint8_t *dst, int dstride)
{
    std::vector<int> xo, xc, yo, yc;
    linearExactAxis(w, dw, xo, xc, fx);
    linearExactAxis(h, dh, yo, yc, fy);
    std::vector<uint32_t> row0(static_cast<size_t>(dw) * ch), row1(static_cast<size_t>(dw) * ch);
    auto hrow = [&](int y, std::vector<uint32_t> &out) {
        const uint8_t *s = src + static_cast<size_t>(y) * stride;
        for (int x = 0; x < dw; ++x) {
            const int x0 = xo[x], x1 = std::min(x0 + 1, w - 1), c1 = xc[x];
            for (int c = 0; c < ch; ++c) out[static_cast<size_t>(x) * ch + c] = s[x0 * ch + c] * (256 - c1) + s[x1 * ch + c] * c1;
        }
    };
    int have0 = -1, have1 = -1;
    for (int y = 0; y < dh; ++y) {
        const int y0 = yo[y], y1 = std::min(y0 + 1, h - 1), c1 = yc[y];
        if (have0 != y0) { if (have1 == y0) { row0.swap(row1); std::swap(have0, have1); } else { hrow(y0, row0); have0 = y0; } }
        if (have1 != y1) { hrow(y1, row1); have1 = y1; }
        uint8_t *d = dst + static_cast<size_t>(y) * dstride;
        const std::vector<uint32_t> &b = (y1 == y0) ? row0 : row1;
        for (int i = 0; i < dw * ch; ++i) d[i] = static_cast<uint8_t>((row0[i] * (256 - c1) + b[i] * c1 + 32768u) >> 16);
    }
}

// cv::remap(INTER_LINEAR, BORDER_REFLECT) on 8-bit images through float maps: 1/32-pixel coordinates,
// weights (32-fy | fy) x (32-fx | fx), (sum + 512) >> 10 == OpenCV's 15-bit table form.
void remapBilinearReflectU8(const uint8_t *src, int w, int h, int stride, int ch, const float *xmap, const float *ymap,
                            int mw, int mh, uint8_t *dst, int dstride)
{
    auto refl = [](int p, int n) {
        if (n == 1) return 0;
        while (static_cast<unsigned>(p) >= static_cast<unsigned>(n)) p = p < 0 ? -p - 1 : 2 * n - 1 - p;
        return p;
    };
    for (int y = 0; y < mh; ++y)
        for (int x = 0; x < mw; ++x) {
            const FixedCoord f = toFixed(xmap[static_cast<size_t>(y) * mw + x], ymap[static_cast<size_t>(y) * mw + x]);
            const int x0 = refl(f.ix, w), x1 = refl(f.ix + 1, w), y0 = refl(f.iy, h), y1 = refl(f.iy + 1, h);
            const int w00 = (32 - f.fy) * (32 - f.fx), w01 = (32 - f.fy) * f.fx, w10 = f.fy * (32 - f.fx), w11 = f.fy * f.fx;
            const uint8_t *r0 = src + static_cast<size_t>(y0) * stride, *r1 = src + static_cast<size_t>(y1) * stride;
            uint8_t *d = dst + static_cast<size_t>(y) * dstride + static_cast<size_t>(x) * ch;
            for (int c = 0; c < ch; ++c)
                d[c] = static_cast<uint8_t>((w00 * r0[x0 * ch + c] + w01 * r0[x1 * ch + c] + w10 * r1[x0 * ch + c] + w11 * r1[x1 * ch + c] + 512) >> 10);
        }
}

// warp of an all-255 mask with INTER_NEAREST / BORDER_CONSTANT (:1013-1017, 1085, 1241): 255 where the rounded source
// position lies inside the frame
void warpedFullMask(const float *xmap, const float *ymap, int mw, int mh, int src_w, int src_h, uint8_t *dst, int dstride)
{
    for (int y = 0; y < mh; ++y)
        for (int x = 0; x < mw; ++x) {
            const int ix = clampS16(roundEven(xmap[static_cast<size_t>(y) * mw + x]));
            const int iy = clampS16(roundEven(ymap[static_cast<size_t>(y) * mw + x]));
            dst[static_cast<size_t>(y) * dstride + x] = (static_cast<unsigned>(ix) < static_cast<unsigned>(src_w) &&
                                                        static_cast<unsigned>(iy) < static_cast<unsigned>(src_h)) ? 255 : 0;
        }
}

double seamWorkAspect(int src_w, int src_h)
{
    return std::min(1.0, std::sqrt(1e5 / (static_cast<double>(src_w) * src_h)));      // include/ocvstitcher.hpp:298
}

}  // namespace pano
