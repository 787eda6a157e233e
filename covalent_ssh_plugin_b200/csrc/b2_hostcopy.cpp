// b2_hostcopy.cpp — see b2_hostcopy.h.  Host code only (compiled by the host compiler).
#include "b2_hostcopy.h"

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace b2host {

namespace {

// 32-byte non-temporal stores; dst must be 32 B aligned, n a multiple of 32.
__attribute__((target("avx2"))) void stream_avx2(uint8_t *dst, const uint8_t *src, size_t n) {
  const __m256i *s = reinterpret_cast<const __m256i *>(src);
  __m256i *d = reinterpret_cast<__m256i *>(dst);
  size_t v = n / 32;
  for (; v >= 4; v -= 4, s += 4, d += 4) {
    const __m256i a = _mm256_loadu_si256(s), b = _mm256_loadu_si256(s + 1);
    const __m256i c = _mm256_loadu_si256(s + 2), e = _mm256_loadu_si256(s + 3);
    _mm256_stream_si256(d, a);
    _mm256_stream_si256(d + 1, b);
    _mm256_stream_si256(d + 2, c);
    _mm256_stream_si256(d + 3, e);
  }
  for (; v; --v, ++s, ++d) _mm256_stream_si256(d, _mm256_loadu_si256(s));
}

bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}

constexpr size_t kStreamMin = 4096;   // below this a plain memcpy wins

class Pool {
 public:
  static Pool &get() {
    static Pool *p = new Pool();   // leaked on purpose: worker threads may outlive static destructors
    return *p;
  }
  int threads() const { return n_threads_; }

  void run(const CopyPiece *pieces, size_t n) {
    size_t total = 0;
    for (size_t i = 0; i < n; ++i) total += pieces[i].bytes;
    if (n_threads_ <= 1 || n < 2 || total < (size_t(1) << 20)) {
      for (size_t i = 0; i < n; ++i) copy_stream(pieces[i].dst, pieces[i].src, pieces[i].bytes);
      return;
    }
    std::lock_guard<std::mutex> one(run_mu_);
    start_workers();
    {
      std::lock_guard<std::mutex> lk(mu_);
      pieces_ = pieces;
      n_ = n;
      next_.store(0, std::memory_order_relaxed);
      working_ = (int)workers_.size();
      ++gen_;
    }
    cv_work_.notify_all();
    drain();
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return working_ == 0; });
    pieces_ = nullptr;
  }

 private:
  Pool() {
    int want = 8;
    if (const char *e = getenv("B2_STAGE_THREADS")) want = atoi(e);
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0) want = std::min(want, hw);
    n_threads_ = std::max(1, want);
  }

  void start_workers() {
    if (!workers_.empty() || n_threads_ <= 1) return;
    for (int i = 0; i < n_threads_ - 1; ++i) {   // the caller is the n-th thread
      workers_.emplace_back([this] { worker(); });
      workers_.back().detach();
    }
  }

  void drain() {
    for (;;) {
      const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
      if (i >= n_) break;
      copy_stream(pieces_[i].dst, pieces_[i].src, pieces_[i].bytes);
    }
  }

  void worker() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
      }
      drain();
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--working_ == 0) cv_done_.notify_all();
      }
    }
  }

  int n_threads_ = 1;
  std::vector<std::thread> workers_;
  std::mutex run_mu_, mu_;
  std::condition_variable cv_work_, cv_done_;
  const CopyPiece *pieces_ = nullptr;
  size_t n_ = 0;
  std::atomic<size_t> next_{0};
  int working_ = 0;
  uint64_t gen_ = 0;
};

}  // namespace

void copy_stream(uint8_t *dst, const uint8_t *src, size_t bytes) {
  if (bytes < kStreamMin || !have_avx2()) {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
  if (head) {
    memcpy(dst, src, head);
    dst += head;
    src += head;
    bytes -= head;
  }
  const size_t body = bytes & ~size_t(31);
  stream_avx2(dst, src, body);
  if (bytes - body) memcpy(dst + body, src + body, bytes - body);
  _mm_sfence();
}

void copy_pieces(const CopyPiece *pieces, size_t n) { Pool::get().run(pieces, n); }

int pool_threads() { return Pool::get().threads(); }

}  // namespace b2host
