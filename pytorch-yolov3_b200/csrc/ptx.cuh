// ptx.cuh — inline-PTX wrappers for the Blackwell (sm_100a) async machinery:
// mbarrier, TMA (cp.async.bulk.tensor, tiled + im2col), tcgen05 (alloc / mma /
// commit / ld / fences).  One thin function per instruction, no policy.
#pragma once
#include <stdint.h>

namespace y3 {
namespace ptx {

// ---- mbarrier -----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// cluster helpers (CTA pairs)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// arrive on a barrier given by a shared::cluster address (own CTA's shared::cta addresses qualify)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Blocking wait with a watchdog: a protocol bug traps (-> CUDA error on the
// host) instead of hanging the GPU box.  ~8 s at 2 GHz before giving up (far beyond any
// legitimate stall: a whole 64-image step takes 4 ms).  Before trapping, the waiting thread leaves a
// record (source file id / line of the wait, block, thread, barrier, parity) in the host-mapped buffer
// registered through y3_debug_set_trap_record, if any: the context is dead after the trap, the host
// memory is not.
#ifndef Y3_FILE_ID
#define Y3_FILE_ID 0
#endif
static __device__ unsigned long long* g_trap_rec = nullptr;  // one copy per translation unit
static inline cudaError_t set_trap_record_tu(unsigned long long* host_mapped) {
  return cudaMemcpyToSymbol(g_trap_rec, &host_mapped, sizeof(host_mapped));
}
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, int line) {
  unsigned long long* rec = g_trap_rec;
  if (rec) {
    // words 1-4: the LAST thread that timed out; words 8-11: the FIRST one; word 5: how many did
    const unsigned long long n = atomicAdd_system(rec + 5, 1ull);
    unsigned long long* r = n == 0 ? rec + 7 : rec;
    r[1] = ((unsigned long long)Y3_FILE_ID << 32) | (unsigned)line;
    r[2] = ((unsigned long long)threadIdx.x << 32) | blockIdx.x;
    r[3] = ((unsigned long long)parity << 32) | bar;
    r[4] = ((unsigned long long)blockDim.x << 32) | gridDim.x;
    rec[0] = 0x59335452415021ull;  // "Y3TRAP!"
    __threadfence_system();
    // give the other stuck threads a moment to leave their record before the context dies
    const long long t0 = clock64();
    while (clock64() - t0 < 40000000ll) {}
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait_at(uint32_t bar, uint32_t parity, int line) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t it = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++it) & 0x3FFu) == 0 && clock64() - t0 > 16000000000ll) mbar_timeout(bar, parity, line);
  }
}
#define mbar_wait(bar, parity) mbar_wait_at(bar, parity, __LINE__)

// ---- TMA --------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tiled load: box lands at smem_dst, completion bytes on mbarrier `bar`.
// CG == 2 (cta_group::2): `bar` may be a shared::cluster address of the PEER CTA's barrier.
template <int CG = 1>
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar,
                                            int32_t c0, int32_t c1) {
  if constexpr (CG == 2) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  }
}
// 2-D tiled store: smem box -> global (bulk async group; rows/cols outside the tensor are clipped).
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
// 4-D tiled load over an NHWC tensor, coordinates (c, w, h, n); out-of-range elements are zero-filled
// (negative coordinates allowed) — used to fetch a spatial patch with its halo in one request.
template <int CG = 1>
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const void* tmap, uint32_t bar, int32_t c,
                                            int32_t w, int32_t h, int32_t n) {
  if constexpr (CG == 2) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
        : "memory");
  }
}
// 4-D tiled store of a (c, w, h, n) box (elements outside the tensor are clipped).
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t smem_src, int32_t c, int32_t w,
                                             int32_t h, int32_t n) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c), "r"(w), "r"(h), "r"(n)
               : "memory");
}
// named barrier over `count` threads (a multiple of 32); id 0 is __syncthreads' own
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// 4-D im2col load over an NHWC tensor: (c, w, h, n) is the first base pixel in
// INPUT coordinates (output pixel * stride - pad), (off_w, off_h) the filter tap.
template <int CG = 1>
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t smem_dst, const void* tmap,
                                                   uint32_t bar, int32_t c, int32_t w, int32_t h,
                                                   int32_t n, uint16_t off_w, uint16_t off_h) {
  if constexpr (CG == 2) {
    asm volatile(
        "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h),
        "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h),
        "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
  }
}

// ---- tcgen05 / TMEM ------------------------------------------------------------------
// CG == 2: the same warp of BOTH CTAs of the pair executes these (collective over the pair).
template <int CG = 1>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // whole warp
  if constexpr (CG == 2)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
template <int CG = 1>
__device__ __forceinline__ void tmem_relinquish() {
  if constexpr (CG == 2) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int CG = 1>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  if constexpr (CG == 2)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread.
// CG == 2: M = 256 over the CTA pair; A and half of B come from each CTA's smem at the same offsets.
template <int CG = 1>
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// mbarrier arrives once all previously issued MMAs of this thread have completed
// (CG == 2: on the barrier at this offset in BOTH CTAs of the pair).
template <int CG = 1>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (CG == 2) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
  }
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Each lane of the warp reads 16 consecutive fp32 columns of its own TMEM lane.
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

}  // namespace ptx
}  // namespace y3
