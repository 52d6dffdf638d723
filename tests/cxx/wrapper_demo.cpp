// C++ caller of the drop-in wrapper include/ocvstitcher_b200.hpp -- what src/replay.cpp:206-288 does with the
// reference's header-only ocvStitcher (init -> calibration -> process per frame-set -> updateMask), here over the
// C ABI.  Test harness only: tests/test_cxx_wrapper.py writes the inputs, runs this program and compares the
// panoramas with the oracle.
//
//   wrapper_demo <in.bin> <out.bin>
// in.bin  (little endian): int32 n, W, H, warp_kind, blender, num_bands, cut[4], nsets; float32 scale; float32 K[9n], R[9n];
//                          per camera: int32 mw, mh, then mw*mh mask bytes; then nsets * n frames of W*H*3 bytes
// out.bin: int32 ow, oh, nsets; nsets panoramas of ow*oh*3 bytes
// exit code: 0 ok, 2 usage / io, 3 init failed (message on stderr), 4 process failed
#include <cstdint>
#include <cstdio>
#include <vector>

#include "ocvstitcher_b200.hpp"

namespace {
template <typename T>
bool rd(FILE *f, T *p, size_t n) { return fread(p, sizeof(T), n, f) == n; }
}  // namespace

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hdr[11];
    float scale;
    if (!rd(f, hdr, 11) || !rd(f, &scale, 1)) return 2;
    const int n = hdr[0], W = hdr[1], H = hdr[2], nsets = hdr[10];
    pano::StitcherParams p;
    p.num_images = n; p.width = W; p.height = H; p.warp_kind = hdr[3]; p.blender = hdr[4]; p.num_bands = hdr[5];
    for (int k = 0; k < 4; ++k) p.cut[k] = hdr[6 + k];
    p.warped_image_scale = scale;
    p.K.resize(9 * n); p.R.resize(9 * n);
    if (!rd(f, p.K.data(), p.K.size()) || !rd(f, p.R.data(), p.R.size())) return 2;
    std::vector<std::vector<unsigned char>> mask_store(n);
    std::vector<pano::Image> masks(n);
    for (int i = 0; i < n; ++i) {
        int32_t wh[2];
        if (!rd(f, wh, 2)) return 2;
        mask_store[i].resize((size_t)wh[0] * wh[1]);
        if (!rd(f, mask_store[i].data(), mask_store[i].size())) return 2;
        masks[i] = pano::Image{mask_store[i].data(), wh[0], wh[1], wh[0]};
    }
    pano::ocvStitcher st;
    if (st.init(p) != pano::RET_OK || st.calibration(&masks) != pano::RET_OK) {
        fprintf(stderr, "init failed: %s\n", st.lastError().c_str());
        return 3;
    }
    const int ow = st.outWidth(), oh = st.outHeight();
    FILE *o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 2; }
    const int32_t oh3[3] = {ow, oh, nsets};
    fwrite(oh3, sizeof(int32_t), 3, o);
    std::vector<std::vector<unsigned char>> frames(n, std::vector<unsigned char>((size_t)W * H * 3));
    std::vector<unsigned char> pano_buf((size_t)ow * oh * 3);
    for (int s = 0; s < nsets; ++s) {
        std::vector<pano::Image> imgs(n);
        for (int i = 0; i < n; ++i) {
            if (!rd(f, frames[i].data(), frames[i].size())) return 2;
            imgs[i] = pano::Image{frames[i].data(), W, H, W * 3};
        }
        pano::Image out{pano_buf.data(), ow, oh, ow * 3};
        if (st.process(imgs, out) != pano::RET_OK) { fprintf(stderr, "process failed: %s\n", st.lastError().c_str()); return 4; }
        if (s == 0 && st.updateMask(masks) != pano::RET_OK) return 4;      // the every-200-frames refresh path (:1218)
        fwrite(pano_buf.data(), 1, pano_buf.size(), o);
    }
    fclose(o);
    fclose(f);
    return 0;
}
