// hostpack.cpp -- see hostpack.h.  Plain C++ (no CUDA): AVX2 path selected at run time, scalar path otherwise.
#include "hostpack.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <unistd.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace pa {

namespace {

// ---- a small persistent pool: parallel_for over [0, n_tasks) ----
class Pool {
 public:
  explicit Pool(int n) : n_(n) {
    for (int i = 0; i < n_; ++i) workers_.emplace_back([this] { run(); });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> g(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return n_; }
  void parallel_for(int n_tasks, const std::function<void(int)>& fn) {
    if (n_tasks <= 0) return;
    // one job at a time: the job state below (fn_, next_, total_, pending_) is shared, and ctypes callers run without
    // the GIL, so two threads (one per GPU, say) may arrive together -- the second waits for the first job to drain
    std::lock_guard<std::mutex> one_job(job_);
    std::unique_lock<std::mutex> g(m_);
    fn_ = &fn; next_ = 0; total_ = n_tasks; pending_ = n_tasks; ++epoch_;
    cv_.notify_all();
    done_.wait(g, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void run() {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> g(m_);
    for (;;) {
      cv_.wait(g, [&] { return stop_ || (epoch_ != seen && next_ < total_); });
      if (stop_) return;
      while (next_ < total_) {
        const int task = next_++;
        const std::function<void(int)>* fn = fn_;
        g.unlock();
        (*fn)(task);
        g.lock();
        if (--pending_ == 0) done_.notify_all();
      }
      seen = epoch_;
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::mutex m_, job_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int next_ = 0, total_ = 0, pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

Pool& pool() {
  // worker threads do not survive fork(): a child process gets a pool of its own (the parent's object is left alone)
  static Pool* p = nullptr;
  static pid_t owner = 0;
  static std::mutex guard;
  std::lock_guard<std::mutex> g(guard);
  if (!p || owner != getpid()) { p = new Pool(host_pack_threads()); owner = getpid(); }
  return *p;
}

inline bool acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// one block of up to 32 bases -> low word, high word; returns validity
inline bool block_scalar(const uint8_t* p, unsigned n, uint32_t* lo, uint32_t* hi) {
  uint32_t l = 0, h = 0;
  bool ok = true;
  for (unsigned i = 0; i < n; ++i) {
    const uint8_t c = p[i];
    ok &= acgt(c);
    l |= (uint32_t)((c >> 1) & 1u) << i;
    h |= (uint32_t)((c >> 2) & 1u) << i;
  }
  *lo = l; *hi = h;
  return ok;
}

bool pack_range_scalar(const uint8_t* bases, const uint64_t* read_off, uint64_t lo, uint64_t a, uint64_t b, uint32_t* planes) {
  bool ok = true;
  const uint64_t base0 = read_off[lo];
  for (uint64_t i = a; i < b; ++i) {
    const uint64_t o = read_off[i], L = read_off[i + 1] - o, nw = (L + 31) / 32;
    uint32_t* w = planes + 2 * ((o - base0) / 32 + (i - lo));
    for (uint64_t c = 0; c < nw; ++c)
      ok &= block_scalar(bases + o + 32 * c, (unsigned)std::min<uint64_t>(32, L - 32 * c), w + c, w + nw + c);
  }
  return ok;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) bool pack_range_avx2(const uint8_t* bases, const uint64_t* read_off, uint64_t lo, uint64_t a,
                                                     uint64_t b, uint32_t* planes, uint64_t safe_end) {
  // safe_end: one past the last base of the chunk -- full 32-byte loads must end at or before it
  const uint64_t base0 = read_off[lo];
  // validity by table: pshufb indexes with the low nibble of a byte (and yields 0 for bytes >= 0x80); only A, C, G, T
  // find themselves at their own nibble ('A' 0x41 -> 1, 'C' 0x43 -> 3, 'T' 0x54 -> 4, 'G' 0x47 -> 7)
  const __m256i lut = _mm256_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1,
                                       -1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1);
  __m256i all_ok = _mm256_set1_epi8(-1);   // AND of the per-byte validity of every full block
  uint32_t bad = 0;
  for (uint64_t i = a; i < b; ++i) {
    const uint64_t o = read_off[i], L = read_off[i + 1] - o, nw = (L + 31) / 32;
    uint32_t* w = planes + 2 * ((o - base0) / 32 + (i - lo));
    const uint8_t* p = bases + o;
    uint64_t c = 0;
    for (; 32 * (c + 1) <= L; ++c) {
      const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 32 * c));
      all_ok = _mm256_and_si256(all_ok, _mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, v), v));
      w[c] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6));        // ASCII bit 1 -> bit 7 of its byte
      w[nw + c] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5));   // ASCII bit 2 -> bit 7
    }
    if (c < nw) {
      const unsigned rem = (unsigned)(L - 32 * c);   // 1 .. 31 bases in the last block
      if (o + 32 * c + 32 <= safe_end) {
        // the 32-byte load runs into the next read (same buffer): mask what is not ours
        const uint32_t keep = (1u << rem) - 1;
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 32 * c));
        const __m256i ok = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, v), v);
        bad |= ~(uint32_t)_mm256_movemask_epi8(ok) & keep;
        w[c] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6)) & keep;
        w[nw + c] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5)) & keep;
      } else {
        uint32_t l, h;
        if (!block_scalar(p + 32 * c, rem, &l, &h)) bad = 1;
        w[c] = l; w[nw + c] = h;
      }
    }
  }
  bad |= ~(uint32_t)_mm256_movemask_epi8(all_ok);
  return bad == 0;
}
#endif

}  // namespace

void host_parallel_for(int n_tasks, const std::function<void(int)>& fn) {
  if (n_tasks <= 1) { for (int t = 0; t < n_tasks; ++t) fn(t); return; }
  pool().parallel_for(n_tasks, fn);
}

int host_pack_threads() {
  static int n = [] {
    if (const char* e = getenv("PA_PACK_THREADS")) { int v = atoi(e); if (v >= 1) return std::min(v, 256); }
    unsigned hc = std::thread::hardware_concurrency();
    if (!hc) hc = 8;
    // one process per GPU: the ranks of a node share its cores (torchrun exports LOCAL_WORLD_SIZE)
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) { int w = atoi(e); if (w > 1) hc = std::max(1u, hc / (unsigned)w); }
    return (int)std::max(1u, std::min(hc, 64u));
  }();
  return n;
}

bool scan_offsets(const uint64_t* read_off, uint64_t lo, uint64_t hi, uint64_t* max_len, uint64_t* min_len, int n_threads) {
  *max_len = 0; *min_len = 0;
  if (hi <= lo) return true;
  const uint64_t n = hi - lo;
  int tasks = (int)std::min<uint64_t>((uint64_t)std::max(1, n_threads), (n + 65535) / 65536);
  if (tasks < 1) tasks = 1;
  std::vector<uint64_t> mx(tasks, 0), mn(tasks, UINT64_MAX);
  std::atomic<bool> ok{true};
  auto work = [&](int t) {
    const uint64_t a = lo + n * (uint64_t)t / tasks, b = lo + n * (uint64_t)(t + 1) / tasks;
    uint64_t m = 0, s = UINT64_MAX;
    bool good = true;
    for (uint64_t i = a; i < b; ++i) {
      good &= read_off[i + 1] >= read_off[i];
      const uint64_t len = read_off[i + 1] - read_off[i];
      m = std::max(m, len); s = std::min(s, len);
    }
    mx[t] = m; mn[t] = s;
    if (!good) ok.store(false, std::memory_order_relaxed);
  };
  if (tasks == 1) work(0); else pool().parallel_for(tasks, work);
  uint64_t s = UINT64_MAX;
  for (int t = 0; t < tasks; ++t) { *max_len = std::max(*max_len, mx[t]); s = std::min(s, mn[t]); }
  *min_len = s == UINT64_MAX ? 0 : s;
  return ok.load();
}

bool pack_reads_planes(const uint8_t* bases, const uint64_t* read_off, uint64_t lo, uint64_t hi, uint32_t* planes,
                       int n_threads) {
  if (hi <= lo) return true;
#if defined(__x86_64__)
  static const bool cpu_avx2 = __builtin_cpu_supports("avx2");
  const char* scalar = getenv("PA_PACK_SCALAR");   // tests: force the portable path
  const bool have_avx2 = cpu_avx2 && !(scalar && *scalar == '1');
#else
  const bool have_avx2 = false;
#endif
  const uint64_t n = hi - lo;
  int tasks = (int)std::min<uint64_t>((uint64_t)std::max(1, n_threads) * 4, (n + 4095) / 4096);
  if (tasks < 1) tasks = 1;
  std::atomic<bool> ok{true};
  auto work = [&](int t) {
    const uint64_t a = lo + n * (uint64_t)t / tasks, b = lo + n * (uint64_t)(t + 1) / tasks;
    bool good;
#if defined(__x86_64__)
    good = have_avx2 ? pack_range_avx2(bases, read_off, lo, a, b, planes, read_off[hi]) : pack_range_scalar(bases, read_off, lo, a, b, planes);
#else
    good = pack_range_scalar(bases, read_off, lo, a, b, planes);
#endif
    if (!good) ok.store(false, std::memory_order_relaxed);
  };
  if (tasks == 1 || n_threads <= 1) { for (int t = 0; t < tasks; ++t) work(t); }
  else pool().parallel_for(tasks, work);
  return ok.load();
}

}  // namespace pa
