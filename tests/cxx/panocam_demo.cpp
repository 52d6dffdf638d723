// C++ caller of the SDK facade pano::panocam (include/ocvstitcher_b200.hpp), shaped like src/demo.cpp / the FSM's use of
// include/panocam.h: init -> (frames in) -> calibration with the application's seam search -> getPanoFrame -> fit2final.
// Test harness only (tests/test_cxx_wrapper.py writes the inputs and compares with the oracle chain).
//
//   panocam_demo <in.bin> <out.bin>
// in.bin : int32 n, W, H, num_bands, rect[4], finalcut, canvas_w, canvas_h; float64 K[9], D[4], newK[9];
//          float32 scale[2]; float32 Ks[n*9], Rs[n*9]; then 2*n camera frames of W*H*4 bytes (8UC4)
// out.bin: int32 w, h; the stacked frame (w*h*3); the canvas (canvas_w*canvas_h*3)
// The "seam finder" of this demo keeps the central 60 % of every low-resolution mask's columns -- a stand-in with the
// GraphCutSeamFinder's interface (it narrows the masks it is given) that the test can restate exactly.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "ocvstitcher_b200.hpp"

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hdr[11];
    double kd[22];
    float sc[2];
    if (fread(hdr, sizeof(int32_t), 11, f) != 11 || fread(kd, sizeof(double), 22, f) != 22 || fread(sc, sizeof(float), 2, f) != 2) return 2;
    const int n = hdr[0], W = hdr[1], H = hdr[2];
    std::vector<float> Ks(9 * n), Rs(9 * n);
    if (fread(Ks.data(), sizeof(float), Ks.size(), f) != Ks.size() || fread(Rs.data(), sizeof(float), Rs.size(), f) != Rs.size()) return 2;

    pano::PanoCamParams p;
    p.num_images = n;
    p.finalcut = hdr[8];
    p.cam.camSrcWidth = p.cam.undistoredWidth = p.cam.outPutWidth = W;
    p.cam.camSrcHeight = p.cam.undistoredHeight = p.cam.outPutHeight = H;
    for (int k = 0; k < 4; ++k) p.cam.rect[k] = hdr[4 + k];
    for (int k = 0; k < 9; ++k) { p.cam.K[k] = kd[k]; p.cam.newK[k] = kd[13 + k]; }
    for (int k = 0; k < 4; ++k) p.cam.distorParams[k] = kd[9 + k];
    for (int r = 0; r < 2; ++r) {
        pano::StitcherParams &s = p.stitcher[r];
        s.width = W; s.height = H; s.num_images = n; s.K = Ks; s.R = Rs; s.warped_image_scale = sc[r];
        s.blender = PANO_BLEND_MULTIBAND; s.num_bands = hdr[3];
    }
    pano::panocam cam;
    if (cam.init(p) != pano::RET_OK) { fprintf(stderr, "panocam init failed: %s\n", cam.lastError().c_str()); return 3; }
    std::vector<std::vector<unsigned char>> frames(2 * n, std::vector<unsigned char>((size_t)W * H * 4));
    for (int i = 0; i < 2 * n; ++i) {
        if (fread(frames[i].data(), 1, frames[i].size(), f) != frames[i].size()) return 2;
        cam.setCamFrame(i, pano::Image{frames[i].data(), W, H, W * 4});
    }
    fclose(f);
    auto finder = [](pano::SeamInputs &s) {
        for (size_t i = 0; i < s.masks.size(); ++i) {
            const int w = s.sizes[2 * i], h = s.sizes[2 * i + 1], lo = w / 5, hi = w - w / 5;
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x)
                    if (x < lo || x >= hi) s.masks[i][(size_t)y * w + x] = 0;
        }
    };
    if (cam.calibration(finder) != pano::RET_OK) { fprintf(stderr, "calibration failed: %s\n", cam.lastError().c_str()); return 3; }
    std::vector<unsigned char> ret((size_t)cam.outWidth() * cam.outHeight() * 3);
    pano::Image ret_img{ret.data(), cam.outWidth(), cam.outHeight(), cam.outWidth() * 3};
    // repeatedly: later calls replay the captured graphs, and the two rings -- two threads, two streams, ONE shared
    // front-end handle -- must give the same bytes every time whatever their relative timing
    std::vector<unsigned char> first;
    for (int rep = 0; rep < 24; ++rep) {
        if (cam.getPanoFrame(ret_img) != pano::RET_OK) { fprintf(stderr, "getPanoFrame failed: %s\n", cam.lastError().c_str()); return 4; }
        if (rep == 0) first = ret;
        else if (ret != first) { fprintf(stderr, "getPanoFrame call %d differs from the first call\n", rep); return 5; }
    }
    pano::FitCanvas fit;
    if (fit.init(ret_img.width, ret_img.height, hdr[9], hdr[10]) != pano::RET_OK) { fprintf(stderr, "fit: %s\n", fit.lastError().c_str()); return 3; }
    std::vector<unsigned char> canvas((size_t)hdr[9] * hdr[10] * 3);
    pano::Image canvas_img{canvas.data(), hdr[9], hdr[10], hdr[9] * 3};
    if (fit.fit2final(ret_img, canvas_img) != pano::RET_OK) return 4;
    FILE *o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 2; }
    const int32_t wh[2] = {ret_img.width, ret_img.height};
    fwrite(wh, sizeof(int32_t), 2, o);
    fwrite(ret.data(), 1, ret.size(), o);
    fwrite(canvas.data(), 1, canvas.size(), o);
    fclose(o);
    return 0;
}
