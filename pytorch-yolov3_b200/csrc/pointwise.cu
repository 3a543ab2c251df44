// pointwise.cu — HBM-bound NHWC bf16 kernels: max-pool (with the reference's right/bottom
// zero padding), fused SPP, shortcut add, channel-slice copy, nearest x2 upsample, input
// packing.  One thread moves one 16-byte vector (8 channels) per step; grids are sized as a
// multiple of the SM count with a grid-stride loop.
#include "common.cuh"

namespace y3 {

static inline int grid_for(long long work_items, int block) {
  long long blocks = (work_items + block - 1) / block;
  long long cap = (long long)num_sms() * 16;  // 16 resident 256-thread CTAs' worth per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---- a6: MaxPool2d (yolov3/darknet.py:16-29) --------------------------------------------
// zero_pad != 0: window [h, h+k) x [w, w+k) on an input zero-padded right/bottom (stride 1);
// else plain floor-mode pooling with stride `stride` (window always inside the input).
__global__ void maxpool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                               int n, int h, int w, int cvec, int ld_x, int ld_y, int k, int stride,
                               int ho, int wo, int zero_pad) {
  pdl_enter();
  const long long total = (long long)n * ho * wo * cvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int ow = (int)(pix % wo);
    pix /= wo;
    const int oh = (int)(pix % ho);
    const int img = (int)(pix / ho);
    const int h0 = oh * stride, w0 = ow * stride;
    const int h1 = min(h0 + k, h), w1 = min(w0 + k, w);
    // a clipped window means padded zeros take part in the max
    const bool clipped = zero_pad && (h0 + k > h || w0 + k > w);
    uint4 m;
    bool first = true;
    if (clipped) { m = make_uint4(0u, 0u, 0u, 0u); first = false; }
    for (int ih = h0; ih < h1; ++ih) {
      const __nv_bfloat16* row = x + ((long long)(img * h + ih) * w) * ld_x + cv * 8;
      for (int iw = w0; iw < w1; ++iw) {
        const uint4 v = ld_nc_16(row + (long long)iw * ld_x);
        m = first ? v : bf16x8_max(m, v);
        first = false;
      }
    }
    st_16(y + ((long long)(img * ho + oh) * wo + ow) * ld_y + cv * 8, m);
  }
}

// ---- SPP: k5/k9/k13 stride-1 pools of one input in one pass -----------------------------
// Windows are nested ([h,h+5) c [h,h+9) c [h,h+13)), so the three maxima are accumulated in
// one sweep of the 13x13 window.
__global__ void spp3_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y5,
                            __nv_bfloat16* __restrict__ y9, __nv_bfloat16* __restrict__ y13, int n,
                            int h, int w, int cvec, int ld_x, int ld_y) {
  pdl_enter();
  const long long total = (long long)n * h * w * cvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int ow = (int)(pix % w);
    pix /= w;
    const int oh = (int)(pix % h);
    const int img = (int)(pix / h);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    // ring r in {5, 9, 13}: elements with max(dh, dw) < r.  Out-of-range taps are zeros.
    uint4 m5, m9, m13;
    bool f5 = true, f9 = true, f13 = true;
    for (int dh = 0; dh < 13; ++dh) {
      const int ih = oh + dh;
      for (int dw = 0; dw < 13; ++dw) {
        const int iw = ow + dw;
        uint4 v = zero;
        if (ih < h && iw < w)
          v = ld_nc_16(x + ((long long)(img * h + ih) * w + iw) * ld_x + cv * 8);
        const int ring = max(dh, dw);
        if (ring < 5) { m5 = f5 ? v : bf16x8_max(m5, v); f5 = false; }
        else if (ring < 9) { m9 = f9 ? v : bf16x8_max(m9, v); f9 = false; }
        else { m13 = f13 ? v : bf16x8_max(m13, v); f13 = false; }
      }
    }
    m9 = bf16x8_max(m9, m5);
    m13 = bf16x8_max(m13, m9);
    const long long o = ((long long)(img * h + oh) * w + ow) * ld_y + cv * 8;
    st_16(y5 + o, m5);
    st_16(y9 + o, m9);
    st_16(y13 + o, m13);
  }
}

// The same three pools for feature maps that fit in shared memory (every SPP block in practice:
// 13x13 at 416, 19x19 at 608): one CTA per (image, group of CVB channel vectors) reads its slice of the
// map ONCE, and since one-sided windows compose — [h, h+5) of [h', h'+5) is [h, h+9), zero padding
// included — y9 = pool5(y5) and y13 = pool5(y9), each pool5 a separable row pass + column pass in
// shared memory.  HBM traffic = the algorithmic 1 read + 3 writes (spp3_kernel re-reads every tap from
// L2: 169 loads per output).
template <int CVB>
__global__ void __launch_bounds__(256)
spp3_tile_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y5,
                 __nv_bfloat16* __restrict__ y9, __nv_bfloat16* __restrict__ y13, int h, int w, int cvec,
                 int ld_x, int ld_y) {
  pdl_enter();
  extern __shared__ uint4 spp_smem[];
  const int hw = h * w, total = hw * CVB;
  uint4* const S = spp_smem;
  uint4* const T = spp_smem + total;
  const int groups = cvec / CVB;
  const int img = blockIdx.x / groups;
  const int c0 = (blockIdx.x - img * groups) * CVB * 8;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  const __nv_bfloat16* const xin = x + (long long)img * hw * ld_x + c0;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int p = e / CVB, cv = e - p * CVB;
    S[e] = ld_nc_16(xin + (long long)p * ld_x + cv * 8);
  }
  __syncthreads();
#pragma unroll 1
  for (int stage = 0; stage < 3; ++stage) {
    __nv_bfloat16* const out = (stage == 0 ? y5 : stage == 1 ? y9 : y13) + (long long)img * hw * ld_y + c0;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {  // row pass: [w, w+5), zeros right of the map
      const int p = e / CVB;
      const int lim = min(5, w - (p % w));
      uint4 m = S[e];
      for (int d = 1; d < lim; ++d) m = bf16x8_max(m, S[e + d * CVB]);
      if (lim < 5) m = bf16x8_max(m, zero);
      T[e] = m;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < total; e += blockDim.x) {  // column pass: [h, h+5), zeros below the map
      const int p = e / CVB, cv = e - p * CVB;
      const int lim = min(5, h - p / w);
      uint4 m = T[e];
      for (int d = 1; d < lim; ++d) m = bf16x8_max(m, T[e + d * w * CVB]);
      if (lim < 5) m = bf16x8_max(m, zero);
      S[e] = m;  // input of the next, wider pool
      st_16(out + (long long)p * ld_y + cv * 8, m);
    }
    __syncthreads();
  }
}

// ---- shortcut add (yolov3/darknet.py:376-379), unfused form ------------------------------
__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                           __nv_bfloat16* __restrict__ y, long long pixels, int cvec, int ld_a,
                           int ld_b, int ld_y) {
  pdl_enter();
  const long long total = pixels * cvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    const long long pix = i / cvec;
    const uint4 va = ld_nc_16(a + pix * ld_a + cv * 8);
    const uint4 vb = ld_nc_16(b + pix * ld_b + cv * 8);
    const uint32_t ua[4] = {va.x, va.y, va.z, va.w}, ub[4] = {vb.x, vb.y, vb.z, vb.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = unpack_bf16x2(ua[j]), fb = unpack_bf16x2(ub[j]);
      o[j] = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
    }
    st_16(y + pix * ld_y + cv * 8, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// ---- route / torch.cat slice copy (yolov3/darknet.py:369-375), unfused form --------------
__global__ void copy_channels_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                     long long pixels, int cvec, int ld_x, int ld_y) {
  pdl_enter();
  const long long total = pixels * cvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    const long long pix = i / cvec;
    st_16(y + pix * ld_y + cv * 8, ld_nc_16(x + pix * ld_x + cv * 8));
  }
}

// ---- nn.Upsample(scale 2, nearest) (yolov3/darknet.py:299-305), unfused form --------------
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                  int n, int h, int w, int cvec, int ld_x, int ld_y) {
  pdl_enter();
  const int ho = 2 * h, wo = 2 * w;
  const long long total = (long long)n * ho * wo * cvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int ow = (int)(pix % wo);
    pix /= wo;
    const int oh = (int)(pix % ho);
    const int img = (int)(pix / ho);
    const uint4 v = ld_nc_16(x + ((long long)(img * h + (oh >> 1)) * w + (ow >> 1)) * ld_x + cv * 8);
    st_16(y + ((long long)(img * ho + oh) * wo + ow) * ld_y + cv * 8, v);
  }
}

// ---- input packing -----------------------------------------------------------------------
// float32 NCHW -> NHWC bf16, channels zero-padded to c_pad (multiple of 8).
__global__ void pack_nchw_f32_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int n,
                                     int c, int hw, int c_pad) {
  pdl_enter();
  const long long total = (long long)n * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / hw);
    const int px = (int)(i - (long long)img * hw);
    const float* src = x + (long long)img * c * hw + px;
    __nv_bfloat16* dst = y + i * c_pad;
    for (int c0 = 0; c0 < c_pad; c0 += 8) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (c0 + j < c) ? __ldg(src + (long long)(c0 + j) * hw) : 0.f;
      st_16(dst + c0, make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                 pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7])));
    }
  }
}

// uint8 BGR HWC -> RGB, fp32 divide by 255 (yolov3/inference.py:332-333), bf16 NHWC padded.
__global__ void pack_bgr_u8_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                   long long pixels, int c_pad) {
  pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels;
       i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* src = x + i * 3;
    const float b = __fdiv_rn((float)src[0], 255.0f);
    const float g = __fdiv_rn((float)src[1], 255.0f);
    const float r = __fdiv_rn((float)src[2], 255.0f);
    __nv_bfloat16* dst = y + i * c_pad;
    st_16(dst, make_uint4(pack_bf16x2(r, g), pack_bf16x2(b, 0.f), 0u, 0u));
    for (int c0 = 8; c0 < c_pad; c0 += 8) st_16(dst + c0, make_uint4(0u, 0u, 0u, 0u));
  }
}

// First-layer im2col: a 3x3 / stride 1 / pad 1 convolution over <= 3 input channels has K = 27,
// which the tensor cores take as one K=32 block.  These kernels write, per pixel, the 27 taps
// (order r, s, c — the weight layout [Cout][R][S][Cin]) + zero padding as one 64-byte row, so
// the first convolution runs as a plain GEMM instead of 9 narrow (32-byte-row) im2col TMA loads.
template <int C, typename Load>
__device__ __forceinline__ void im2col3x3_pixel(Load load, int h, int w, int y, int x,
                                                __nv_bfloat16* __restrict__ dst, int k_pad) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = y + r - 1;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ix = x + s - 1;
      const bool in = iy >= 0 && iy < h && ix >= 0 && ix < w;
#pragma unroll
      for (int cc = 0; cc < C; ++cc) v[(r * 3 + s) * C + cc] = in ? load(iy, ix, cc) : 0.f;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (q * 8 < k_pad)
      st_16(dst + q * 8, make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                    pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7])));
}

template <int C>
__global__ void im2col3x3_nchw_f32_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int n,
                                          int h, int w, int k_pad) {
  pdl_enter();
  const long long total = (long long)n * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % w);
    const int py = (int)((i / w) % h);
    const int img = (int)(i / ((long long)w * h));
    const float* base = x + (long long)img * C * h * w;
    im2col3x3_pixel<C>([&](int iy, int ix, int cc) { return __ldg(base + ((long long)cc * h + iy) * w + ix); }, h, w,
                       py, px, y + i * k_pad, k_pad);
  }
}

__global__ void im2col3x3_bgr_u8_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h,
                                        int w, int k_pad) {
  pdl_enter();
  const long long total = (long long)n * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % w);
    const int py = (int)((i / w) % h);
    const int img = (int)(i / ((long long)w * h));
    const uint8_t* base = x + (long long)img * h * w * 3;
    // channel cc of the network input is RGB: R = byte 2, G = byte 1, B = byte 0 (inference.py:332)
    // v * fl(1/255) rounds to the same bf16 as the reference's fp32 v / 255 for all 256 byte
    // values (checked exhaustively in tests/test_host_logic.py), without 27 IEEE divides per pixel
    im2col3x3_pixel<3>([&](int iy, int ix, int cc) {
      return __fmul_rn((float)__ldg(base + ((long long)iy * w + ix) * 3 + (2 - cc)), 0.003921568859368563f); },
                       h, w, py, px, y + i * k_pad, k_pad);
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace y3

using namespace y3;

extern "C" {

int y3_maxpool(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t ld_x,
               int32_t ld_y, int32_t ksize, int32_t stride, void* stream) {
  Y3_CHECK_ARG(x && y, "maxpool: null pointer");
  Y3_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "maxpool: bad shape n=%d h=%d w=%d c=%d", n, h, w, c);
  Y3_CHECK_ARG(ksize >= 1 && stride >= 1, "maxpool: bad ksize=%d stride=%d", ksize, stride);
  Y3_CHECK_ARG(ld_x >= c && ld_y >= c && ld_x % 8 == 0 && ld_y % 8 == 0, "maxpool: bad pitch");
  Y3_CHECK_ARG(aligned16(x) && aligned16(y), "maxpool: pointers must be 16-byte aligned");
  const int zero_pad = (ksize > 1 && stride == 1) ? 1 : 0;
  int ho, wo;
  if (zero_pad) { ho = h; wo = w; }
  else {
    Y3_CHECK_ARG(h >= ksize && w >= ksize, "maxpool: input smaller than window");
    ho = (h - ksize) / stride + 1;
    wo = (w - ksize) / stride + 1;
  }
  const long long work = (long long)n * ho * wo * (c / 8);
  Y3_CUDA_OK(launch_kernel(maxpool_kernel, dim3(grid_for(work, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c / 8, ld_x, ld_y, ksize, stride, ho, wo, zero_pad));
  Y3_LAUNCH_OK("maxpool_kernel");
  return Y3_OK;
}

int y3_spp3(const void* x, void* y5, void* y9, void* y13, int32_t n, int32_t h, int32_t w, int32_t c,
            int32_t ld_x, int32_t ld_y, void* stream) {
  Y3_CHECK_ARG(x && y5 && y9 && y13, "spp3: null pointer");
  Y3_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "spp3: bad shape");
  Y3_CHECK_ARG(ld_x >= c && ld_y >= c && ld_x % 8 == 0 && ld_y % 8 == 0, "spp3: bad pitch");
  Y3_CHECK_ARG(aligned16(x) && aligned16(y5) && aligned16(y9) && aligned16(y13), "spp3: alignment");
  constexpr int CVB = 4;  // 64 channels per CTA: 2 x (h*w*64 B) of shared memory
  const size_t smem = 2 * (size_t)h * w * CVB * sizeof(uint4);
  if ((c / 8) % CVB == 0 && smem <= 200 * 1024 && (long long)n * (c / 8 / CVB) < (1ll << 31)) {
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
      Y3_CUDA_OK(cudaFuncSetAttribute(spp3_tile_kernel<CVB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      smem_set = 200 * 1024;
    }
    Y3_CUDA_OK(launch_kernel(spp3_tile_kernel<CVB>, dim3(n * (c / 8 / CVB)), dim3(256), smem, (cudaStream_t)stream,
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y5, (__nv_bfloat16*)y9, (__nv_bfloat16*)y13, h, w, c / 8, ld_x, ld_y));
    Y3_LAUNCH_OK("spp3_tile_kernel");
    return Y3_OK;
  }
  const long long work = (long long)n * h * w * (c / 8);
  Y3_CUDA_OK(launch_kernel(spp3_kernel, dim3(grid_for(work, 128)), dim3(128), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y5, (__nv_bfloat16*)y9, (__nv_bfloat16*)y13, n, h, w, c / 8, ld_x, ld_y));
  Y3_LAUNCH_OK("spp3_kernel");
  return Y3_OK;
}

int y3_add(const void* a, const void* b, void* y, int64_t pixels, int32_t c, int32_t ld_a, int32_t ld_b,
           int32_t ld_y, void* stream) {
  Y3_CHECK_ARG(a && b && y, "add: null pointer");
  Y3_CHECK_ARG(pixels > 0 && c > 0 && c % 8 == 0, "add: bad shape");
  Y3_CHECK_ARG(ld_a >= c && ld_b >= c && ld_y >= c && ld_a % 8 == 0 && ld_b % 8 == 0 && ld_y % 8 == 0, "add: bad pitch");
  Y3_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(y), "add: alignment");
  Y3_CUDA_OK(launch_kernel(add_kernel, dim3(grid_for(pixels * (c / 8), 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, pixels, c / 8, ld_a, ld_b, ld_y));
  Y3_LAUNCH_OK("add_kernel");
  return Y3_OK;
}

int y3_copy_channels(const void* x, void* y, int64_t pixels, int32_t c, int32_t ld_x, int32_t ld_y, void* stream) {
  Y3_CHECK_ARG(x && y, "copy_channels: null pointer");
  Y3_CHECK_ARG(pixels > 0 && c > 0 && c % 8 == 0, "copy_channels: bad shape");
  Y3_CHECK_ARG(ld_x >= c && ld_y >= c && ld_x % 8 == 0 && ld_y % 8 == 0, "copy_channels: bad pitch");
  Y3_CHECK_ARG(aligned16(x) && aligned16(y), "copy_channels: alignment");
  Y3_CUDA_OK(launch_kernel(copy_channels_kernel, dim3(grid_for(pixels * (c / 8), 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, pixels, c / 8, ld_x, ld_y));
  Y3_LAUNCH_OK("copy_channels_kernel");
  return Y3_OK;
}

int y3_upsample2x(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t ld_x,
                  int32_t ld_y, void* stream) {
  Y3_CHECK_ARG(x && y, "upsample2x: null pointer");
  Y3_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "upsample2x: bad shape");
  Y3_CHECK_ARG(ld_x >= c && ld_y >= c && ld_x % 8 == 0 && ld_y % 8 == 0, "upsample2x: bad pitch");
  Y3_CHECK_ARG(aligned16(x) && aligned16(y), "upsample2x: alignment");
  const long long work = (long long)n * 4 * h * w * (c / 8);
  Y3_CUDA_OK(launch_kernel(upsample2x_kernel, dim3(grid_for(work, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c / 8, ld_x, ld_y));
  Y3_LAUNCH_OK("upsample2x_kernel");
  return Y3_OK;
}

int y3_pack_nchw_f32(const float* x, void* y, int32_t n, int32_t c, int32_t h, int32_t w, int32_t c_pad,
                     void* stream) {
  Y3_CHECK_ARG(x && y, "pack_nchw_f32: null pointer");
  Y3_CHECK_ARG(n > 0 && c > 0 && h > 0 && w > 0 && c_pad >= c && c_pad % 8 == 0, "pack_nchw_f32: bad shape");
  Y3_CHECK_ARG(aligned16(y), "pack_nchw_f32: alignment");
  const long long work = (long long)n * h * w;
  Y3_CUDA_OK(launch_kernel(pack_nchw_f32_kernel, dim3(grid_for(work, 256)), dim3(256), 0, (cudaStream_t)stream, 
      x, (__nv_bfloat16*)y, n, c, h * w, c_pad));
  Y3_LAUNCH_OK("pack_nchw_f32_kernel");
  return Y3_OK;
}

int y3_pack_bgr_u8(const uint8_t* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c_pad, void* stream) {
  Y3_CHECK_ARG(x && y, "pack_bgr_u8: null pointer");
  Y3_CHECK_ARG(n > 0 && h > 0 && w > 0 && c_pad >= 8 && c_pad % 8 == 0, "pack_bgr_u8: bad shape");
  Y3_CHECK_ARG(aligned16(y), "pack_bgr_u8: alignment");
  const long long work = (long long)n * h * w;
  Y3_CUDA_OK(launch_kernel(pack_bgr_u8_kernel, dim3(grid_for(work, 256)), dim3(256), 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, work, c_pad));
  Y3_LAUNCH_OK("pack_bgr_u8_kernel");
  return Y3_OK;
}

int y3_im2col3x3_nchw_f32(const float* x, void* y, int32_t n, int32_t c, int32_t h, int32_t w, int32_t k_pad,
                          void* stream) {
  Y3_CHECK_ARG(x && y, "im2col3x3_nchw_f32: null pointer");
  Y3_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && 9 * c <= k_pad && k_pad <= 32 && k_pad % 8 == 0,
               "im2col3x3_nchw_f32: need 9*c <= k_pad <= 32 (c=%d k_pad=%d)", c, k_pad);
  Y3_CHECK_ARG(aligned16(y), "im2col3x3_nchw_f32: alignment");
  const long long work = (long long)n * h * w;
  const int grid = grid_for(work, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (c == 3) Y3_CUDA_OK(launch_kernel(im2col3x3_nchw_f32_kernel<3>, dim3(grid), dim3(256), 0, st, x, (__nv_bfloat16*)y, n, h, w, k_pad));
  else if (c == 2) Y3_CUDA_OK(launch_kernel(im2col3x3_nchw_f32_kernel<2>, dim3(grid), dim3(256), 0, st, x, (__nv_bfloat16*)y, n, h, w, k_pad));
  else Y3_CUDA_OK(launch_kernel(im2col3x3_nchw_f32_kernel<1>, dim3(grid), dim3(256), 0, st, x, (__nv_bfloat16*)y, n, h, w, k_pad));
  Y3_LAUNCH_OK("im2col3x3_nchw_f32_kernel");
  return Y3_OK;
}

int y3_im2col3x3_bgr_u8(const uint8_t* x, void* y, int32_t n, int32_t h, int32_t w, int32_t k_pad, void* stream) {
  Y3_CHECK_ARG(x && y, "im2col3x3_bgr_u8: null pointer");
  Y3_CHECK_ARG(n > 0 && h > 0 && w > 0 && k_pad == 32, "im2col3x3_bgr_u8: k_pad must be 32");
  Y3_CHECK_ARG(aligned16(y), "im2col3x3_bgr_u8: alignment");
  const long long work = (long long)n * h * w;
  Y3_CUDA_OK(launch_kernel(im2col3x3_bgr_u8_kernel, dim3(grid_for(work, 256)), dim3(256), 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, n, h, w,
                                                                                  k_pad));
  Y3_LAUNCH_OK("im2col3x3_bgr_u8_kernel");
  return Y3_OK;
}

}  // extern "C"
