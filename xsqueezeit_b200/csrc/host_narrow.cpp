// host_narrow.cpp -- host side of the PCIe boundary for int32 genotype rows.
//
// bcf_get_genotypes widens the record's int8 FORMAT/GT payload to int32 (htslib vcf.c:4728-4795) and
// fill_genotype_array hands int32 back (accessor_internals.hpp:399-413): 4 bytes per genotype over PCIe
// for values that fit one byte whenever a record has at most 63 alleles.  The C ABI therefore moves such
// rows across the bus in the BCF int8 encoding (0x80 = missing, 0x81 = vector end) and converts on the
// host, chunk by chunk, on a small worker pool that runs beside the DMA.  This is a TRANSPORT encoding
// only: no genotype arithmetic happens here (all of it is in the kernels), and a value that does not fit
// makes the caller fall back to moving int32.
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <immintrin.h>
#include <sched.h>

#include "host_narrow.hpp"

namespace xsi {
namespace {

// ---- int32 -> int8: low 7 bits | sign bit.  Returns the OR of everything that would be lost. ----
//   v in [0,127] -> v;  INT32_MIN -> 0x80;  INT32_MIN+1 -> 0x81;  anything else is "bad"
inline uint32_t narrow_scalar(const int32_t* src, int8_t* dst, size_t n) {
    uint32_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t u = (uint32_t)src[i];
        dst[i] = (int8_t)((u & 0x7Fu) | ((u >> 24) & 0x80u));
        bad |= (u & 0x7FFFFF80u) | ((uint32_t)(src[i] >> 31) & u & 0x7Eu);
    }
    return bad;
}

__attribute__((target("avx512f,avx512bw"))) uint32_t narrow_avx512(const int32_t* src, int8_t* dst, size_t n) {
    const __m512i m7f = _mm512_set1_epi32(0x7F), m80 = _mm512_set1_epi32(0x80), mhi = _mm512_set1_epi32(0x7FFFFF80),
                  m7e = _mm512_set1_epi32(0x7E);
    __m512i bad = _mm512_setzero_si512();
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
#pragma GCC unroll 4
        for (int k = 0; k < 4; ++k) {
            const __m512i v = _mm512_loadu_si512(src + i + 16 * k);
            const __m512i b = _mm512_or_si512(_mm512_and_si512(v, m7f), _mm512_and_si512(_mm512_srli_epi32(v, 24), m80));
            bad = _mm512_or_si512(bad, _mm512_and_si512(v, mhi));
            bad = _mm512_or_si512(bad, _mm512_and_si512(_mm512_and_si512(_mm512_srai_epi32(v, 31), v), m7e));
            _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 16 * k), _mm512_cvtepi32_epi8(b));
        }
    }
    uint32_t r = _mm512_test_epi32_mask(bad, bad) ? 1u : 0u;
    return r | narrow_scalar(src + i, dst + i, n - i);
}

__attribute__((target("avx2"))) uint32_t narrow_avx2(const int32_t* src, int8_t* dst, size_t n) {
    const __m256i m7f = _mm256_set1_epi32(0x7F), m80 = _mm256_set1_epi32(0x80), mhi = _mm256_set1_epi32(0x7FFFFF80),
                  m7e = _mm256_set1_epi32(0x7E);
    const __m256i perm = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    __m256i bad = _mm256_setzero_si256();
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        __m256i b[4];
        for (int k = 0; k < 4; ++k) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 8 * k));
            b[k] = _mm256_or_si256(_mm256_and_si256(v, m7f), _mm256_and_si256(_mm256_srli_epi32(v, 24), m80));
            bad = _mm256_or_si256(bad, _mm256_and_si256(v, mhi));
            bad = _mm256_or_si256(bad, _mm256_and_si256(_mm256_and_si256(_mm256_srai_epi32(v, 31), v), m7e));
        }
        // values are 0..255: unsigned-saturating packs keep them; the packs interleave 128-bit lanes
        const __m256i p01 = _mm256_packus_epi32(b[0], b[1]), p23 = _mm256_packus_epi32(b[2], b[3]);
        const __m256i p = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(p01, p23), perm);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), p);
    }
    uint32_t r = _mm256_testz_si256(bad, bad) ? 0u : 1u;
    return r | narrow_scalar(src + i, dst + i, n - i);
}

// ---- int8 -> int32: (b & 0x7F) | (b & 0x80) << 24, written with streaming stores (the destination is the
// caller's row buffer, far larger than any cache: no read-for-ownership traffic) ----
inline void widen_scalar(const int8_t* src, int32_t* dst, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const uint32_t b = (uint8_t)src[i];
        dst[i] = (int32_t)((b & 0x7Fu) | ((b & 0x80u) << 24));
    }
}

__attribute__((target("avx512f,avx512bw"))) void widen_avx512(const int8_t* src, int32_t* dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63)) { widen_scalar(src + i, dst + i, 1); ++i; }
    const __m512i m7f = _mm512_set1_epi32(0x7F), m80 = _mm512_set1_epi32(0x80);
    for (; i + 16 <= n; i += 16) {
        const __m512i b = _mm512_cvtepu8_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)));
        const __m512i v = _mm512_or_si512(_mm512_and_si512(b, m7f), _mm512_slli_epi32(_mm512_and_si512(b, m80), 24));
        _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + i), v);
    }
    widen_scalar(src + i, dst + i, n - i);
    _mm_sfence();
}

__attribute__((target("avx2"))) void widen_avx2(const int8_t* src, int32_t* dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) { widen_scalar(src + i, dst + i, 1); ++i; }
    const __m256i m7f = _mm256_set1_epi32(0x7F), m80 = _mm256_set1_epi32(0x80);
    for (; i + 8 <= n; i += 8) {
        const __m256i b = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(src + i)));
        const __m256i v = _mm256_or_si256(_mm256_and_si256(b, m7f), _mm256_slli_epi32(_mm256_and_si256(b, m80), 24));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), v);
    }
    widen_scalar(src + i, dst + i, n - i);
    _mm_sfence();
}

int isa_level() {
    static const int lvl = [] {
        __builtin_cpu_init();
        if (const char* s = getenv("XSI_HOST_ISA")) return atoi(s);
        if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) return 2;
        if (__builtin_cpu_supports("avx2")) return 1;
        return 0;
    }();
    return lvl;
}

// ---- worker pool: fork-join over [0, n_tasks), one job at a time (callers from several contexts queue) ----
class Pool {
public:
    explicit Pool(unsigned n) {
        for (unsigned t = 0; t + 1 < n; ++t) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    unsigned size() const { return (unsigned)workers_.size() + 1; }
    void run(size_t n_tasks, const std::function<void(size_t)>& fn) {
        if (n_tasks == 0) return;
        std::lock_guard<std::mutex> job(job_m_);
        if (n_tasks == 1 || workers_.empty()) { for (size_t i = 0; i < n_tasks; ++i) fn(i); return; }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn; n_ = n_tasks; next_.store(0); left_.store(n_tasks); active_ = true; ++gen_;
        }
        cv_.notify_all();
        work();  // the calling thread takes its share
        std::unique_lock<std::mutex> g(m_);
        done_cv_.wait(g, [this] { return left_.load() == 0 && busy_ == 0; });
        active_ = false;  // a worker that wakes up late finds no job and goes back to sleep
        fn_ = nullptr;
    }

private:
    void work() {
        for (;;) {
            const size_t i = next_.fetch_add(1);
            if (i >= n_) break;
            (*fn_)(i);
            left_.fetch_sub(1);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                if (!active_) continue;
                ++busy_;
            }
            work();
            {
                std::lock_guard<std::mutex> g(m_);
                --busy_;
            }
            done_cv_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, job_m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(size_t)>* fn_ = nullptr;
    size_t n_ = 0;
    std::atomic<size_t> next_{0}, left_{0};
    unsigned busy_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false, active_ = false;
};

Pool& pool() {
    static Pool* p = [] {
        unsigned n = 0;
        if (const char* s = getenv("XSI_HOST_THREADS")) n = (unsigned)atoi(s);
        if (n == 0) {
            cpu_set_t set;
            CPU_ZERO(&set);
            if (sched_getaffinity(0, sizeof set, &set) == 0) n = (unsigned)CPU_COUNT(&set);
            if (n == 0) n = std::thread::hardware_concurrency();
            if (n == 0) n = 1;
            if (n > 32) n = 32;
        }
        return new Pool(n);  // lives until process exit (worker threads must not outlive their pool object)
    }();
    return *p;
}

constexpr size_t SLICE = 1u << 18;  // elements per task

}  // namespace

unsigned host_threads() { return pool().size(); }

void host_parallel_for(size_t n_tasks, const std::function<void(size_t)>& fn) { pool().run(n_tasks, fn); }

bool narrow_i32_to_i8(const int32_t* src, int8_t* dst, size_t n) {
    const int lvl = isa_level();
    auto one = [lvl](const int32_t* s, int8_t* d, size_t m) {
        return lvl >= 2 ? narrow_avx512(s, d, m) : lvl == 1 ? narrow_avx2(s, d, m) : narrow_scalar(s, d, m);
    };
    if (n <= SLICE) return one(src, dst, n) == 0;
    std::atomic<uint32_t> bad{0};
    const size_t tasks = (n + SLICE - 1) / SLICE;
    pool().run(tasks, [&](size_t t) {
        const size_t a = t * SLICE, m = std::min(SLICE, n - a);
        if (one(src + a, dst + a, m)) bad.store(1, std::memory_order_relaxed);
    });
    return bad.load() == 0;
}

void widen_rows_i8_to_i32(const int8_t* src, size_t src_stride, int32_t* dst, size_t dst_stride, const uint32_t* len,
                          size_t n_rows) {
    const int lvl = isa_level();
    auto one = [lvl](const int8_t* s, int32_t* d, size_t m) {
        if (lvl >= 2) widen_avx512(s, d, m);
        else if (lvl == 1) widen_avx2(s, d, m);
        else widen_scalar(s, d, m);
    };
    size_t total = 0;
    for (size_t r = 0; r < n_rows; ++r) total += len[r];
    if (total <= SLICE) { for (size_t r = 0; r < n_rows; ++r) one(src + r * src_stride, dst + r * dst_stride, len[r]); return; }
    // tasks = (row, slice of the row)
    size_t max_len = 0;
    for (size_t r = 0; r < n_rows; ++r) max_len = std::max<size_t>(max_len, len[r]);
    const size_t per_row = (max_len + SLICE - 1) / SLICE;
    const size_t rows_per_task = per_row > 1 ? 1 : std::max<size_t>(1, SLICE / std::max<size_t>(1, max_len));
    if (per_row > 1) {
        pool().run(n_rows * per_row, [&](size_t t) {
            const size_t r = t / per_row, a = (t % per_row) * SLICE;
            if (a < len[r]) one(src + r * src_stride + a, dst + r * dst_stride + a, std::min<size_t>(SLICE, len[r] - a));
        });
    } else {
        pool().run((n_rows + rows_per_task - 1) / rows_per_task, [&](size_t t) {
            const size_t r0 = t * rows_per_task, r1 = std::min(n_rows, r0 + rows_per_task);
            for (size_t r = r0; r < r1; ++r) one(src + r * src_stride, dst + r * dst_stride, len[r]);
        });
    }
}

}  // namespace xsi
