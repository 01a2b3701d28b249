// hostconv.cpp -- host-side staging helper of the host-pointer batch entry points.
//
// In the reference pipeline descriptors are "doubles holding float32 values"
// (M/sift/siftdescriptor.c:500-527 computes them in float and copies them into a double matrix), so
// a 128 x K descriptor block can cross PCIe at half the bytes: the block is narrowed to float into
// pinned staging by a small thread pool WHILE the previous chunk is in flight, every value is
// checked to survive the round trip ((double)(float)x == x), and the device widens it again --
// the arithmetic on the device is unchanged (double accumulation, siftmatch.c:61).  A block with a
// value that does not survive is sent as doubles.  No arithmetic of the path runs on the host.
#include <immintrin.h>
#include <stddef.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace pre3 {

class HostPool {
 public:
  explicit HostPool(int n) : stop_(false), gen_(0), pending_(0) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); });
  }
  ~HostPool() {
    {
      std::unique_lock<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // fn(part, nparts) on every worker and on the caller; returns when all parts are done
  void run(const std::function<void(int, int)>& fn) {
    const int nparts = size() + 1;
    {
      std::unique_lock<std::mutex> lk(m_);
      fn_ = &fn;
      pending_ = size();
      ++gen_;
    }
    cv_.notify_all();
    fn(nparts - 1, nparts);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(int id) {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void(int, int)>* fn;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        fn = fn_;
      }
      (*fn)(id, size() + 1);
      {
        std::unique_lock<std::mutex> lk(m_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  bool stop_;
  unsigned long gen_;
  int pending_;
  const std::function<void(int, int)>* fn_ = nullptr;
};

__attribute__((target("avx2"))) static bool narrow_avx2(const double* s, float* d, size_t n) {
  __m256d bad = _mm256_setzero_pd();
  size_t i = 0;
  // the staging buffer is written once and read by the DMA engine only: stream the stores past the cache
  const bool stream = (reinterpret_cast<size_t>(d) & 31) == 0;
  for (; i + 8 <= n; i += 8) {
    const __m256d a = _mm256_loadu_pd(s + i), b = _mm256_loadu_pd(s + i + 4);
    const __m128 fa = _mm256_cvtpd_ps(a), fb = _mm256_cvtpd_ps(b);
    bad = _mm256_or_pd(bad, _mm256_cmp_pd(_mm256_cvtps_pd(fa), a, _CMP_NEQ_UQ));
    bad = _mm256_or_pd(bad, _mm256_cmp_pd(_mm256_cvtps_pd(fb), b, _CMP_NEQ_UQ));
    if (stream)
      _mm256_stream_ps(d + i, _mm256_set_m128(fb, fa));
    else
      _mm256_storeu_ps(d + i, _mm256_set_m128(fb, fa));
  }
  if (stream) _mm_sfence();
  bool ok = _mm256_movemask_pd(bad) == 0;
  for (; i < n; ++i) {
    const float f = (float)s[i];
    d[i] = f;
    ok = ok && ((double)f == s[i]);
  }
  return ok;
}

static bool narrow_scalar(const double* s, float* d, size_t n) {
  bool ok = true;
  for (size_t i = 0; i < n; ++i) {
    const float f = (float)s[i];
    d[i] = f;
    ok = ok && ((double)f == s[i]);  // NaN fails the test: sent as doubles
  }
  return ok;
}

HostPool* host_pool_create(int threads) { return new HostPool(threads); }
void host_pool_destroy(HostPool* p) { delete p; }
int host_pool_size(const HostPool* p) { return p->size(); }

// d[i] = (float)s[i]; true iff every value survives the round trip
bool host_narrow(HostPool* pool, const double* s, float* d, size_t n) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  std::atomic<int> ok(1);
  pool->run([&](int part, int nparts) {
    const size_t per = ((n + nparts - 1) / nparts + 63) & ~(size_t)63;
    const size_t lo = per * part < n ? per * part : n, hi = lo + per < n ? lo + per : n;
    if (hi > lo && !(avx2 ? narrow_avx2(s + lo, d + lo, hi - lo) : narrow_scalar(s + lo, d + lo, hi - lo))) ok = 0;
  });
  return ok.load() != 0;
}

}  // namespace pre3
