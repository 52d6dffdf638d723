// A few persistent host threads that move frames between the caller's PAGEABLE buffers (what a cv::Mat holds) and the
// handle's pinned staging memory, so that pano_process does not hand pageable pointers to the CUDA driver (whose own
// single-threaded staging copy runs at ~12 GB/s: 4.1 ms per config-2 frame-set against 1.1 ms from pinned memory).
#pragma once
#include <cstddef>
#include <cstdint>

namespace pano {

class HostPool;
HostPool *host_pool_create(int threads);
void host_pool_destroy(HostPool *p);
// rows x row_bytes from src (pitch sstride) to dst (pitch dstride) on a worker thread; returns a ticket.
// stream_stores: dst is only read by the copy engine afterwards -> non-temporal stores (x86-64 with AVX2)
int host_pool_copy2d(HostPool *p, void *dst, size_t dstride, const void *src, size_t sstride, size_t row_bytes, size_t rows,
                     bool stream_stores = false);
void host_pool_wait(HostPool *p, int ticket);       // blocks until that copy has finished
void host_pool_wait_all(HostPool *p);               // ... until every submitted copy has finished; tickets restart at 0

}  // namespace pano
