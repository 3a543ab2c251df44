// api.cu — library-level entry points: version, errors, device check, launch counter.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

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

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("Y3_NO_PDL");
    cached = (e && e[0] == '1') ? 0 : 1;
  }
  return cached == 1;
}

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

}  // namespace y3

extern "C" {

int y3_abi_version(void) { return Y3_ABI_VERSION; }

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
  const long long total = (long long)n * bytes_each;
  int t = threads < 1 ? 1 : (threads > 16 ? 16 : threads);
  const long long by_size = total >> 21;  // a thread is worth starting for ~2 MB of copying
  if (t > by_size) t = by_size < 1 ? 1 : (int)by_size;
  if (t > n) t = n;
  auto work = [=](int w) {
    for (int i = w; i < n; i += t) memcpy(static_cast<char*>(dst) + (long long)i * bytes_each, srcs[i], (size_t)bytes_each);
  };
  if (t == 1) { work(0); return Y3_OK; }
  std::vector<std::thread> pool;
  pool.reserve(t - 1);
  for (int w = 1; w < t; ++w) pool.emplace_back(work, w);
  work(0);
  for (auto& th : pool) th.join();
  return Y3_OK;
}

long long y3_launch_count(void) { return y3::g_launches; }
void y3_reset_launch_count(void) { y3::g_launches = 0; }

}  // extern "C"
