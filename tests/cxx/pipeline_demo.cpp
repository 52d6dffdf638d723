// C++ caller of the whole widened path through include/ocvstitcher_b200.hpp, shaped like the two-ring rig of
// src/panocamimpl.cpp:154-360 / src/master.cpp:258-326: calibration files -> per-ring stitcher with the nvCam
// front end chained in -> process(up), process(down) -> resize + vconcat + separator.  Test harness only
// (tests/test_cxx_wrapper.py writes the inputs and compares the stacked frame with the oracle).
//
//   pipeline_demo <cfgdir/> <in.bin> <out.bin>
// cfgdir holds cameraparaout_1.txt (upper ring) and cameraparaout_2.txt (lower ring); the parsed parameters of ring 1
// are re-saved as cameraparaout_9.txt (saveCameraParams round trip).
// in.bin: int32 n, W, H, num_bands, rect[4]; float64 K[9], D[4], newK[9]; then 2 * n camera frames of W*H*4 bytes (8UC4)
// out.bin: int32 w, h; the stacked frame (w*h*3 bytes)
#include <cstdint>
#include <cstdio>
#include <vector>

#include "ocvstitcher_b200.hpp"

int main(int argc, char **argv)
{
    if (argc != 4) { fprintf(stderr, "usage: %s cfgdir/ in.bin out.bin\n", argv[0]); return 2; }
    FILE *f = fopen(argv[2], "rb");
    if (!f) { perror(argv[2]); return 2; }
    int32_t hdr[8];
    double kd[22];
    if (fread(hdr, sizeof(int32_t), 8, f) != 8 || fread(kd, sizeof(double), 22, f) != 22) return 2;
    const int n = hdr[0], W = hdr[1], H = hdr[2];

    pano::CamParams cp;
    cp.camSrcWidth = cp.undistoredWidth = cp.outPutWidth = W;
    cp.camSrcHeight = cp.undistoredHeight = cp.outPutHeight = H;
    for (int k = 0; k < 4; ++k) cp.rect[k] = hdr[4 + k];
    for (int k = 0; k < 9; ++k) { cp.K[k] = kd[k]; cp.newK[k] = kd[13 + k]; }
    for (int k = 0; k < 4; ++k) cp.distorParams[k] = kd[9 + k];
    pano::nvCamFrontEnd cam;
    if (cam.init(cp) != pano::RET_OK) { fprintf(stderr, "front end init failed: %s\n", cam.lastError().c_str()); return 3; }

    pano::ocvStitcher ring[2];
    std::vector<unsigned char> pano_buf[2];
    pano::Image pano_img[2];
    for (int r = 0; r < 2; ++r) {
        pano::StitcherParams p;
        std::string err;
        if (pano::loadCameraParams(argv[1], 1 + r, n, p, &err) != pano::RET_OK) { fprintf(stderr, "%s\n", err.c_str()); return 3; }
        if (r == 0 && pano::saveCameraParams(argv[1], 9, p) != pano::RET_OK) return 3;
        p.width = W; p.height = H; p.blender = PANO_BLEND_MULTIBAND; p.num_bands = hdr[3];
        if (ring[r].init(p) != pano::RET_OK || ring[r].calibration() != pano::RET_OK || ring[r].attachFrontEnd(cam.handle()) != pano::RET_OK) {
            fprintf(stderr, "ring %d init failed: %s\n", r, ring[r].lastError().c_str());
            return 3;
        }
        pano_buf[r].resize((size_t)ring[r].outWidth() * ring[r].outHeight() * 3);
        pano_img[r] = pano::Image{pano_buf[r].data(), ring[r].outWidth(), ring[r].outHeight(), ring[r].outWidth() * 3};
    }
    std::vector<std::vector<unsigned char>> frames(n, std::vector<unsigned char>((size_t)W * H * 4));
    for (int r = 0; r < 2; ++r) {
        std::vector<pano::Image> imgs(n);
        for (int i = 0; i < n; ++i) {
            if (fread(frames[i].data(), 1, frames[i].size(), f) != frames[i].size()) return 2;
            imgs[i] = pano::Image{frames[i].data(), W, H, W * 4};
        }
        if (ring[r].process(imgs, pano_img[r]) != pano::RET_OK) { fprintf(stderr, "process failed: %s\n", ring[r].lastError().c_str()); return 4; }
    }
    fclose(f);
    pano::RingComposer rc;
    if (rc.init(pano_img[0].width, pano_img[0].height, pano_img[1].width, pano_img[1].height) != pano::RET_OK) {
        fprintf(stderr, "ring composer: %s\n", rc.lastError().c_str());
        return 3;
    }
    std::vector<unsigned char> ret((size_t)rc.outWidth() * rc.outHeight() * 3);
    pano::Image ret_img{ret.data(), rc.outWidth(), rc.outHeight(), rc.outWidth() * 3};
    if (rc.compose(pano_img[0], pano_img[1], ret_img) != pano::RET_OK) return 4;
    FILE *o = fopen(argv[3], "wb");
    if (!o) { perror(argv[3]); return 2; }
    const int32_t wh[2] = {rc.outWidth(), rc.outHeight()};
    fwrite(wh, sizeof(int32_t), 2, o);
    fwrite(ret.data(), 1, ret.size(), o);
    fclose(o);
    return 0;
}
