#include "host_pool.hpp"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace pano {

namespace {
struct Job { int ticket; void *dst; size_t dstride; const void *src; size_t sstride, row_bytes, rows; bool stream; };

// Copy whose destination is only ever read by the copy engine (the pinned bounce buffer of the input frames): streaming
// stores keep the destination lines out of the cache hierarchy and skip the read-for-ownership, a third of the memory
// traffic of a plain memcpy.  The fence orders the streaming stores before the hand-over to the thread that starts the DMA.
#if defined(__x86_64__)
__attribute__((target("avx2"))) void copy_stream_avx2(char *dst, const char *src, size_t n)
{
    size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
    if (head > n) head = n;
    std::memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
    }
    for (; i + 32 <= n; i += 32)
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)));
    _mm_sfence();
    std::memcpy(dst + i, src + i, n - i);
}
const bool g_avx2 = __builtin_cpu_supports("avx2");
#endif

void copy_bytes(char *dst, const char *src, size_t n, bool stream)
{
#if defined(__x86_64__)
    if (stream && g_avx2 && n >= 4096) { copy_stream_avx2(dst, src, n); return; }
#endif
    (void)stream;
    std::memcpy(dst, src, n);
}
}

class HostPool {
public:
    explicit HostPool(int n)
    {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    ~HostPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        work_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int submit(const Job &j)
    {
        int ticket;
        {
            std::lock_guard<std::mutex> lk(m_);
            ticket = (int)done_.size();
            done_.push_back(0);
            Job q = j;
            q.ticket = ticket;
            jobs_.push_back(q);
        }
        work_.notify_one();
        return ticket;
    }
    void wait(int ticket)
    {
        std::unique_lock<std::mutex> lk(m_);
        finished_.wait(lk, [&] { return ticket < (int)done_.size() && done_[ticket]; });
    }
    void wait_all()
    {
        std::unique_lock<std::mutex> lk(m_);
        finished_.wait(lk, [&] {
            for (char d : done_)
                if (!d) return false;
            return true;
        });
        done_.clear();
    }

private:
    void run()
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m_);
                work_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
                if (jobs_.empty()) return;      // stop requested and nothing left
                j = jobs_.front();
                jobs_.pop_front();
            }
            if (j.sstride == j.row_bytes && j.dstride == j.row_bytes) {
                copy_bytes(static_cast<char *>(j.dst), static_cast<const char *>(j.src), j.row_bytes * j.rows, j.stream);
            } else {
                for (size_t r = 0; r < j.rows; ++r)
                    copy_bytes(static_cast<char *>(j.dst) + r * j.dstride, static_cast<const char *>(j.src) + r * j.sstride, j.row_bytes, j.stream);
            }
            {
                std::lock_guard<std::mutex> lk(m_);
                done_[j.ticket] = 1;
            }
            finished_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable work_, finished_;
    std::deque<Job> jobs_;
    std::vector<char> done_;
    bool stop_ = false;
};

HostPool *host_pool_create(int threads) { return new HostPool(threads < 1 ? 1 : threads); }
void host_pool_destroy(HostPool *p) { delete p; }
int host_pool_copy2d(HostPool *p, void *dst, size_t dstride, const void *src, size_t sstride, size_t row_bytes, size_t rows,
                     bool stream_stores)
{
    return p->submit(Job{0, dst, dstride, src, sstride, row_bytes, rows, stream_stores});
}
void host_pool_wait(HostPool *p, int ticket) { p->wait(ticket); }
void host_pool_wait_all(HostPool *p) { p->wait_all(); }

}  // namespace pano
