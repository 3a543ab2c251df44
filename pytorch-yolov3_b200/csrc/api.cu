// api.cu — library-level entry points: version, errors, device check, launch counter.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include <immintrin.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace y3 {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches += n; }

static thread_local int g_pdl = -1;  // per thread; -1: not decided yet (environment), else y3_set_pdl's value

bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("Y3_NO_PDL");
    g_pdl = (e && e[0] == '1') ? 0 : 1;
  }
  return g_pdl == 1;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }

int num_sms() {
  static int cached_dev = -1;
  static int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
    cached_sms = sms;
    cached_dev = dev;
  }
  return cached_sms;
}

// memcpy with non-temporal stores: the staging buffer is written once and read by the DMA engine, never by
// a core, so the destination lines need neither a read-for-ownership nor a place in the caches (one third
// less DRAM traffic per staged byte, which is what eight ranks staging 33 MB per batch compete for).
__attribute__((target("avx2"))) static void stream_copy_avx2(char* dst, const char* src, size_t len) {
  const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
  if (head >= len) { memcpy(dst, src, len); return; }
  memcpy(dst, src, head);
  dst += head; src += head; len -= head;
  size_t i = 0;
  for (; i + 128 <= len; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < len) memcpy(dst + i, src + i, len - i);
}

static bool use_stream_copy() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("Y3_STAGE_NT");
    v = (e ? e[0] == '1' : true) && __builtin_cpu_supports("avx2") ? 1 : 0;
  }
  return v == 1;
}

// Host staging pool of y3_stage_images: parked worker threads (created once, never joined — the pool
// is leaked on purpose so nothing is torn down under a waiting thread at process exit) copy 256 KB
// chunks of the stacked batch handed out by an atomic counter; the caller copies along with them.
class StagePool {
 public:
  void run(char* dst, const void* const* srcs, int n, long long bytes_each, int threads) {
    const long long total = (long long)n * bytes_each;
    const long long chunks = (total + kChunk - 1) / kChunk;
    int helpers = (threads < 1 ? 1 : (threads > kMaxThreads ? kMaxThreads : threads)) - 1;
    if (helpers > chunks - 1) helpers = (int)(chunks - 1);
    std::unique_lock<std::mutex> call(call_mu_);  // one batch at a time
    dst_ = dst; srcs_ = srcs; bytes_each_ = bytes_each; total_ = total; chunks_ = chunks;
    nt_ = use_stream_copy();
    next_.store(0, std::memory_order_relaxed);
    if (helpers > 0) {
      std::lock_guard<std::mutex> lk(mu_);
      while ((int)workers_.size() < helpers) {
        workers_.emplace_back(&StagePool::worker, this, (int)workers_.size());
        workers_.back().detach();
      }
      wanted_ = helpers;
      running_ = helpers;
      ++generation_;
    }
    if (helpers > 0) cv_work_.notify_all();
    copy_chunks();
    if (helpers > 0) {
      std::unique_lock<std::mutex> lk(mu_);
      cv_done_.wait(lk, [&] { return running_ == 0; });
    }
  }

 private:
  static constexpr long long kChunk = 256 << 10;
  static constexpr int kMaxThreads = 16;

  void copy_chunks() {
    for (;;) {
      const long long c = next_.fetch_add(1, std::memory_order_relaxed);
      if (c >= chunks_) return;
      long long off = c * kChunk;
      long long left = total_ - off < kChunk ? total_ - off : kChunk;
      while (left > 0) {  // a chunk may straddle images
        const long long img = off / bytes_each_, in = off - img * bytes_each_;
        const long long len = bytes_each_ - in < left ? bytes_each_ - in : left;
        if (nt_) stream_copy_avx2(dst_ + off, static_cast<const char*>(srcs_[img]) + in, (size_t)len);
        else memcpy(dst_ + off, static_cast<const char*>(srcs_[img]) + in, (size_t)len);
        off += len;
        left -= len;
      }
    }
  }

  void worker(int index) {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] { return generation_ != seen && index < wanted_; });
        seen = generation_;
      }
      copy_chunks();
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--running_ == 0) cv_done_.notify_one();
      }
    }
  }

  std::mutex call_mu_, mu_;
  std::condition_variable cv_work_, cv_done_;
  std::vector<std::thread> workers_;
  unsigned long long generation_ = 0;
  int wanted_ = 0, running_ = 0;
  char* dst_ = nullptr;
  const void* const* srcs_ = nullptr;
  long long bytes_each_ = 0, total_ = 0, chunks_ = 0;
  bool nt_ = false;
  std::atomic<long long> next_{0};
};

StagePool& stage_pool() {
  static StagePool* pool = new StagePool();
  return *pool;
}

}  // namespace y3

extern "C" {

int y3_abi_version(void) { return Y3_ABI_VERSION; }

int y3_set_pdl(int on) {
  const int before = y3::pdl_enabled() ? 1 : 0;
  y3::set_pdl(on);
  return before;
}

const char* y3_last_error(void) { return y3::g_err; }

int y3_check_device(int dev) {
  int major = 0, minor = 0;
  Y3_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  Y3_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    y3::set_error("device %d is sm_%d%d; libyolov3_b200 is built for sm_100a only and has no fallback",
                  dev, major, minor);
    return Y3_EARCH;
  }
  return Y3_OK;
}

int y3_stage_images(void* dst, const void* const* srcs, int32_t n, int64_t bytes_each, int32_t threads) {
  Y3_CHECK_ARG(dst && srcs && n > 0 && bytes_each > 0, "stage_images: bad arguments");
  for (int i = 0; i < n; ++i) Y3_CHECK_ARG(srcs[i] != nullptr, "stage_images: image %d is null", i);
  y3::stage_pool().run(static_cast<char*>(dst), srcs, n, bytes_each, threads);
  return Y3_OK;
}

int y3_debug_set_trap_record(void* host_mapped) {
  unsigned long long* p = static_cast<unsigned long long*>(host_mapped);
  Y3_CUDA_OK(y3::conv_umma_set_trap_record(p));
  Y3_CUDA_OK(y3::conv_patch_set_trap_record(p));
  Y3_CUDA_OK(y3::conv_chain_set_trap_record(p));
  return Y3_OK;
}

long long y3_launch_count(void) { return y3::g_launches; }
void y3_reset_launch_count(void) { y3::g_launches = 0; }

}  // extern "C"
