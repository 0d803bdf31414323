// Host-side helpers of the C ABI (include/aznet_b200.h): no CUDA in this file, compiled by g++.
//
// azn_host_f32_to_bf16: round-to-nearest-even f32 -> bf16 of a host array on a pool of worker threads.
// The reference hands its 'fc' net the conv5_3 map as f32 blobs (lib/detect/test.py:228-236) and the engine keeps
// maps as bf16, so the batched host-facing call (aznet_b200/pipeline.py) may narrow a batch BEFORE it crosses PCIe:
// half the bytes over the link that bounds that call.  Same rounding as the device conversion
// (__float2bfloat16_rn: ties to even, NaN -> 0x7fff), so both routes put identical bits in HBM.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

extern "C" int azn_host_f32_to_bf16(const float *src, uint16_t *dst, size_t n, int threads);
extern "C" int azn_host_threads(void);
void azn_set_error(const char *fmt, ...);

namespace {

inline uint16_t cvt_one(uint32_t x) {
    if ((x & 0x7fffffffu) > 0x7f800000u) return 0x7fffu;                  // NaN, as __float2bfloat16_rn
    return (uint16_t)((x + 0x7fffu + ((x >> 16) & 1u)) >> 16);            // ties to even; overflow rounds to inf
}

void cvt_scalar(const float *src, uint16_t *dst, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        uint32_t x;
        memcpy(&x, src + i, 4);
        dst[i] = cvt_one(x);
    }
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void cvt_avx2(const float *src, uint16_t *dst, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31u)) {                      // up to a 32-byte boundary of the output
        uint32_t x;
        memcpy(&x, src + i, 4);
        dst[i++] = cvt_one(x);
    }
    const __m256i bias = _mm256_set1_epi32(0x7fff), one = _mm256_set1_epi32(1);
    const __m256i absm = _mm256_set1_epi32(0x7fffffff), inf = _mm256_set1_epi32(0x7f800000);
    const __m256i nanv = _mm256_set1_epi32(0x7fff);
    for (; i + 16 <= n; i += 16) {
        __m256i a = _mm256_loadu_si256((const __m256i *)(src + i));
        __m256i b = _mm256_loadu_si256((const __m256i *)(src + i + 8));
        const __m256i na = _mm256_cmpgt_epi32(_mm256_and_si256(a, absm), inf);      // |x| > inf (signed compare is fine: both < 2^31)
        const __m256i nb = _mm256_cmpgt_epi32(_mm256_and_si256(b, absm), inf);
        a = _mm256_srli_epi32(_mm256_add_epi32(_mm256_add_epi32(a, bias), _mm256_and_si256(_mm256_srli_epi32(a, 16), one)), 16);
        b = _mm256_srli_epi32(_mm256_add_epi32(_mm256_add_epi32(b, bias), _mm256_and_si256(_mm256_srli_epi32(b, 16), one)), 16);
        a = _mm256_blendv_epi8(a, nanv, na);
        b = _mm256_blendv_epi8(b, nanv, nb);
        __m256i p = _mm256_packus_epi32(a, b);                            // [a0..3 b0..3 | a4..7 b4..7]
        p = _mm256_permute4x64_epi64(p, 0xD8);                            // [a0..7 b0..7]
        _mm256_stream_si256((__m256i *)(dst + i), p);                     // the output is read next by the DMA engine, not by us
    }
    _mm_sfence();
    for (; i < n; ++i) {
        uint32_t x;
        memcpy(&x, src + i, 4);
        dst[i] = cvt_one(x);
    }
}
#endif

void cvt_range(const float *src, uint16_t *dst, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) { cvt_avx2(src, dst, n); return; }
#endif
    cvt_scalar(src, dst, n);
}

// A small persistent pool: the call is made once per batch (a few milliseconds of work), thread start-up per call
// would be a visible fraction of it.  One job at a time (callers are serialised by the mutex).
class Pool {
public:
    ~Pool() {
        {
            std::lock_guard<std::mutex> g(m_);
            quit_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void run(const float *src, uint16_t *dst, size_t n, int threads) {
        std::lock_guard<std::mutex> call(call_m_);
        const size_t grain = 4096;                                       // elements; keeps every part 32-byte aligned on the output
        size_t parts = (n + grain - 1) / grain;
        if (parts > (size_t)threads) parts = (size_t)threads;
        if (parts <= 1) { cvt_range(src, dst, n); return; }
        ensure((int)parts - 1);
        {
            std::lock_guard<std::mutex> g(m_);
            src_ = src; dst_ = dst; n_ = n; parts_ = parts; pending_ = (int)parts - 1;
            ++epoch_;
        }
        cv_.notify_all();
        part(0);                                                         // the caller works too
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [&] { return pending_ == 0; });
    }

private:
    void part(size_t k) {
        const size_t grain = 4096;
        const size_t blocks = (n_ + grain - 1) / grain;
        const size_t b0 = blocks * k / parts_, b1 = blocks * (k + 1) / parts_;
        const size_t lo = b0 * grain, hi = b1 * grain < n_ ? b1 * grain : n_;
        if (hi > lo) cvt_range(src_ + lo, dst_ + lo, hi - lo);
    }
    void ensure(int n_workers) {
        while ((int)workers_.size() < n_workers) {
            const int id = (int)workers_.size() + 1;
            long seen;
            {
                std::lock_guard<std::mutex> g(m_);
                seen = epoch_;
            }
            workers_.emplace_back([this, id, seen]() mutable {
                for (;;) {
                    std::unique_lock<std::mutex> g(m_);
                    cv_.wait(g, [&] { return epoch_ != seen; });
                    seen = epoch_;
                    if (quit_) return;
                    const bool mine = (size_t)id < parts_;
                    g.unlock();
                    if (!mine) continue;
                    part((size_t)id);
                    g.lock();
                    if (--pending_ == 0) done_.notify_one();
                }
            });
        }
    }
    std::mutex m_, call_m_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    const float *src_ = nullptr;
    uint16_t *dst_ = nullptr;
    size_t n_ = 0, parts_ = 0;
    int pending_ = 0;
    long epoch_ = 0;
    bool quit_ = false;
};

Pool &pool() {
    static Pool *p = new Pool();          // leaked on purpose: worker threads must not be joined from a static destructor at exit
    return *p;
}

}  // namespace

extern "C" int azn_host_threads(void) {
    const unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

extern "C" int azn_host_f32_to_bf16(const float *src, uint16_t *dst, size_t n, int threads) {
    if ((!src || !dst) && n) {
        azn_set_error("azn_host_f32_to_bf16: null buffer");
        return 1;                                                        // AZN_ERR_INVALID
    }
    if (threads <= 0) threads = azn_host_threads();
    if (threads > 64) threads = 64;
    pool().run(src, dst, n, threads);
    return 0;
}
