// umma_shift_probe.cu — hardware probe (run on the GPU box), not part of the product path.
//
// Question: may the start address of a tcgen05 K-major swizzled smem descriptor be offset by whole
// ROWS (not 1024B-aligned), and may SBO (the stride between 8-row groups) be any multiple of the
// row pitch?  Both are needed to feed a 3x3 convolution's nine taps from ONE smem window of the
// input (A re-use) instead of nine im2col loads.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I pytorch-yolov3_b200/csrc \
//        tools/umma_shift_probe.cu -o gpurun_out/umma_shift_probe && gpurun_out/umma_shift_probe
//
// For every (swizzle span, start row offset, SBO, base_offset policy) it runs one M=128 x N=64 MMA
// chain and reports, per configuration, whether D[m] == A[row0 + (m/8)*SBO_rows + m%8] * B^T.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "ptx.cuh"

using namespace y3;

static constexpr int A_ROWS = 512;  // rows resident in smem
static constexpr int N = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

struct ProbeParams {
  int span_bytes;     // 128 / 64 / 32
  int a_off_bytes;    // start offset of A from its 1024B-aligned base
  int sbo_bytes;      // stride between 8-row groups
  int base_off;       // descriptor bits 49..51
  int dst_off_bytes;  // TMA destination offset of A inside the buffer (tests non-1024B-aligned TMA dst)
  float* out;         // [128][64]
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;                    // up to 512 rows x 128 B = 64 KB (+ dst offset slack 8 KB)
  const uint32_t b_base = base + 72 * 1024;        // 64 rows x 128 B = 8 KB
  const uint32_t bar = base + 81 * 1024;           // full barrier
  const uint32_t mma_bar = bar + 8;
  const uint32_t tmem_slot = bar + 16;
  uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kelems = p.span_bytes / 2;

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(mma_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc<1>(tmem_slot, 64);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;

  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar, A_ROWS * p.span_bytes + N * p.span_bytes);
    ptx::tma_load_2d(a_base + p.dst_off_bytes, &tmap_a, bar, 0, 0);
    ptx::tma_load_2d(a_base + p.dst_off_bytes + 256 * p.span_bytes, &tmap_a, bar, 0, 256);
    ptx::tma_load_2d(b_base, &tmap_b, bar, 0, 0);
    ptx::mbar_wait(bar, 0);
    ptx::tc_fence_after();
    const uint64_t layout = p.span_bytes == 128 ? 2 : p.span_bytes == 64 ? 4 : 6;
    const uint64_t hi_a = ((uint64_t(p.sbo_bytes) >> 4) << 32) | (1ull << 46) | (uint64_t(p.base_off & 7) << 49) | (layout << 61);
    const uint64_t hi_b = ((uint64_t(8 * p.span_bytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
    const uint32_t a_addr = a_base + p.dst_off_bytes + p.a_off_bytes;
    const uint64_t da = hi_a | (1ull << 16) | uint64_t((a_addr >> 4) & 0x3FFFu);
    const uint64_t db = hi_b | (1ull << 16) | uint64_t((b_base >> 4) & 0x3FFFu);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    for (int k = 0; k < kelems / 16; ++k) ptx::umma_bf16_ss<1>(tmem, da + 2u * k, db + 2u * k, idesc, k != 0);
    ptx::umma_commit<1>(mma_bar);
  }
  __syncwarp();
  ptx::mbar_wait(mma_bar, 0);
  ptx::tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    ptx::tmem_ld_x16(tmem + (uint32_t(warp * 32) << 16) + c0, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) p.out[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<1>(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

int main() {
  EncodeTiledFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  enc = (EncodeTiledFn)fn;

  srand(7);
  int bad_total = 0;
  for (int span : {128, 64, 32}) {
    const int kel = span / 2;
    std::vector<__nv_bfloat16> ha(A_ROWS * kel), hb(N * kel);
    std::vector<float> fa(A_ROWS * kel), fb(N * kel);
    for (size_t i = 0; i < ha.size(); ++i) { fa[i] = float(rand() % 9 - 4); ha[i] = __float2bfloat16(fa[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { fb[i] = float(rand() % 9 - 4); hb[i] = __float2bfloat16(fb[i]); }
    __nv_bfloat16 *da, *db;
    float* dout;
    CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dout, 128 * N * 4));
    CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMapSwizzle sw = span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUtensorMap ta, tb;
    {
      cuuint64_t dims[2] = {(cuuint64_t)kel, A_ROWS}; cuuint64_t str[1] = {(cuuint64_t)span};
      cuuint32_t box[2] = {(cuuint32_t)kel, 256}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode A failed %d\n", (int)r); return 1; }
      cuuint64_t dimb[2] = {(cuuint64_t)kel, N}; cuuint32_t boxb[2] = {(cuuint32_t)kel, N};
      r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dimb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 84 * 1024));
    // configurations: start row offsets, SBO in rows, TMA destination offsets (in rows)
    const int row_offs[] = {0, 1, 2, 3, 5, 8, 9, 17, 54, 107};
    const int sbo_rows[] = {8, 12, 16, 17, 24, 35};
    const int dst_rows[] = {0, 1, 5};
    for (int dst : dst_rows)
      for (int sbo : sbo_rows)
        for (int ro : row_offs) {
          if (ro + 15 * sbo + 8 > A_ROWS) continue;
          for (int policy = 0; policy < 2; ++policy) {
            ProbeParams p;
            p.span_bytes = span;
            p.dst_off_bytes = dst * span;
            p.a_off_bytes = ro * span;
            p.sbo_bytes = sbo * span;
            p.out = dout;
            // policy 0: base_offset field 0; policy 1: (start address >> 7) & 7 of the offset from the 1024B-aligned base
            p.base_off = policy == 0 ? 0 : (((p.dst_off_bytes + p.a_off_bytes) >> 7) & 7);
            if (policy == 1 && p.base_off == 0) continue;
            CK(cudaMemset(dout, 0, 128 * N * 4));
            probe_kernel<<<1, 128, 84 * 1024>>>(ta, tb, p);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("span %d dst %d sbo %d ro %d policy %d: CUDA error %s\n", span, dst, sbo, ro, policy, cudaGetErrorString(e)); return 2; }
            std::vector<float> out(128 * N);
            CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
            int bad_rows = 0, first_bad = -1;
            for (int m = 0; m < 128; ++m) {
              const int r = ro + (m / 8) * sbo + (m % 8);  // row index relative to the TMA destination
              bool ok = true;
              for (int n = 0; n < N && ok; ++n) {
                float acc = 0;
                for (int k = 0; k < kel; ++k) acc += fa[r * kel + k] * fb[n * kel + k];
                ok = acc == out[m * N + n];
              }
              if (!ok) { ++bad_rows; if (first_bad < 0) first_bad = m; }
            }
            printf("span %3d dst_row %d sbo_rows %2d row_off %3d base_off %d : %s (bad rows %d, first %d)\n", span, dst, sbo, ro,
                   p.base_off, bad_rows ? "MISMATCH" : "ok", bad_rows, first_bad);
            bad_total += bad_rows != 0;
          }
        }
    cudaFree(da); cudaFree(db); cudaFree(dout);
  }
  printf("configurations with mismatches: %d\n", bad_total);
  return 0;
}
