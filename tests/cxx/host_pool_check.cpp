// Unit check of the host worker pool behind pano_process' pageable-buffer path (img-stitching_b200/csrc/host_pool.*):
// strided 2-D copies, with and without streaming stores, odd sizes and alignments, tickets waited in and out of order.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../img-stitching_b200/csrc/host_pool.hpp"

using namespace pano;

static int check_case(HostPool *p, size_t row_bytes, size_t rows, size_t spad, size_t dpad, size_t soff, size_t doff, bool stream, int chunks)
{
    const size_t ss = row_bytes + spad, ds = row_bytes + dpad;
    std::vector<uint8_t> src(ss * rows + soff + 64), dst(ds * rows + doff + 64, 0xEE), want(dst);
    for (size_t i = 0; i < src.size(); ++i) src[i] = (uint8_t)(i * 131 + (i >> 9) * 7 + 3);
    for (size_t r = 0; r < rows; ++r) std::memcpy(&want[doff + r * ds], &src[soff + r * ss], row_bytes);
    std::vector<int> tickets;
    const size_t band = (rows + chunks - 1) / chunks;
    for (int k = 0; k < chunks; ++k) {
        const size_t r0 = std::min(rows, k * band), nr = std::min(rows, r0 + band) - r0;
        tickets.push_back(host_pool_copy2d(p, &dst[doff + r0 * ds], ds, &src[soff + r0 * ss], ss, row_bytes, nr, stream));
    }
    for (int k = chunks - 1; k >= 0; k -= 2) host_pool_wait(p, tickets[k]);      // out of order, some never waited singly
    host_pool_wait_all(p);
    if (dst != want) {
        std::printf("MISMATCH row_bytes=%zu rows=%zu spad=%zu dpad=%zu soff=%zu doff=%zu stream=%d chunks=%d\n", row_bytes, rows, spad, dpad,
                    soff, doff, (int)stream, chunks);
        return 1;
    }
    return 0;
}

int main()
{
    int bad = 0, n = 0;
    for (int threads : {1, 3, 8}) {
        HostPool *p = host_pool_create(threads);
        for (bool stream : {false, true})
            for (size_t row_bytes : {(size_t)1, (size_t)31, (size_t)4096, (size_t)5760, (size_t)7681, (size_t)100003})
                for (size_t rows : {(size_t)1, (size_t)5, (size_t)37})
                    for (size_t pad : {(size_t)0, (size_t)13})
                        for (size_t off : {(size_t)0, (size_t)1, (size_t)17}) {
                            bad += check_case(p, row_bytes, rows, pad, pad ? 64 - pad : 0, off, 31 - off, stream, 1 + (int)(rows % 4));
                            ++n;
                        }
        host_pool_wait_all(p);      // nothing pending: must return at once
        host_pool_destroy(p);
    }
    std::printf("host_pool: %d cases, %d mismatches\n", n, bad);
    return bad ? 1 : 0;
}
