// common.cuh — shared host/device helpers for libyolov3_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/yolov3_b200.h"

namespace y3 {

// ---- host-side error plumbing (thread-local message, int status) ---------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define Y3_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      y3::set_error(__VA_ARGS__);               \
      return Y3_EINVAL;                         \
    }                                           \
  } while (0)

#define Y3_CUDA_OK(expr)                                                      \
  do {                                                                        \
    cudaError_t e_ = (expr);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      y3::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),   \
                    __FILE__, __LINE__);                                      \
      return Y3_ECUDA;                                                        \
    }                                                                         \
  } while (0)

// cudaGetLastError after a launch; counts the launch for y3_launch_count().
#define Y3_LAUNCH_OK(name)                                                    \
  do {                                                                        \
    cudaError_t e_ = cudaGetLastError();                                      \
    if (e_ != cudaSuccess) {                                                  \
      y3::set_error("launch of %s failed: %s", name, cudaGetErrorString(e_)); \
      return Y3_ECUDA;                                                        \
    }                                                                         \
    y3::count_launch();                                                       \
  } while (0)

// Tiled (non-im2col) bf16 tensor map of rank 2..5 (conv_umma.cu).  `map` is a 64B-aligned CUtensorMap;
// strides_bytes has rank-1 entries (dimension 0 is dense); swizzle_bytes is 128, 64 or 32.
int encode_tiled_map(void* map, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes);

// 3x3 / stride 1 / pad 1 layers whose tiles stay >= 93 % full in the row-padded virtual space run on the
// shared-memory-patch kernel (conv_patch.cu); returns -1 when the layer should use the im2col kernel.
int conv3x3_patch_try(const ::y3_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual,
                      void* y, cudaStream_t stream);

// watchdog record pointer of each translation unit that waits on mbarriers (ptx.cuh: mbar_timeout)
cudaError_t conv_umma_set_trap_record(unsigned long long* host_mapped);
cudaError_t conv_patch_set_trap_record(unsigned long long* host_mapped);
cudaError_t conv_chain_set_trap_record(unsigned long long* host_mapped);

int num_sms();  // SM count of the current device (cached per device)
bool pdl_enabled();  // programmatic dependent launch on every kernel (Y3_NO_PDL=1 / y3_set_pdl(0) turn it off)
void set_pdl(int on);

#ifdef __CUDACC__
// Every kernel of the library is launched with the programmatic-stream-serialization attribute:
// kernel i+1's CTAs may become resident (and run their prologue) while kernel i drains, and block
// in pdl_enter()/pdl_wait() until kernel i has completed and its writes are visible.  Captured
// into a CUDA graph the attribute becomes a programmatic dependency edge.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// ---- device helpers ---------------------------------------------------------
#ifdef __CUDACC__

// Programmatic dependent launch: let the next kernel of the stream start its prologue, then wait
// until the previous kernel has completed (no-ops when launched without the attribute).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_launch_dependents(); pdl_wait(); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte streaming global accesses (activations are touched once per layer).
__device__ __forceinline__ uint4 ld_nc_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_16(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
// max over packed bf16 pairs
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a);
  __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162 m = __hmax2(x, y);
  return *reinterpret_cast<uint32_t*>(&m);
}
__device__ __forceinline__ uint4 bf16x8_max(const uint4& a, const uint4& b) {
  return make_uint4(bf16x2_max(a.x, b.x), bf16x2_max(a.y, b.y),
                    bf16x2_max(a.z, b.z), bf16x2_max(a.w, b.w));
}

#endif  // __CUDACC__

}  // namespace y3
