// api.cu — library-level entry points: version, errors, device check, launch counter.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

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

long long y3_launch_count(void) { return y3::g_launches; }
void y3_reset_launch_count(void) { y3::g_launches = 0; }

}  // extern "C"
