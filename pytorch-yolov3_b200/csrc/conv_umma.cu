// conv_umma.cu — Darknet [convolutional] block as an implicit GEMM on sm_100a.
//
// Replaces torch.nn.Conv2d + BatchNorm2d(eval) + LeakyReLU(0.1) as built at
// yolov3/darknet.py:244-257 and executed at :367-368, with the shortcut add
// (:376-379) and the nearest x2 upsample (:299-305) available as epilogue fusions.
//
//   D[M = N*Ho*Wo, Cout] = im2col(X)[M, K = R*S*Cin] * W[Cout, K]^T      (bf16 x bf16 -> fp32)
//
// Data path per CTA (persistent, one CTA per SM, 6 warps):
//   warp 0  : TMA producer.  A tile = 128 output pixels x BLOCK_K channels of one filter tap,
//             fetched by ONE im2col-mode TMA (zero-filled halo, stride handled by the tensor
//             map) — or a plain 2-D tiled TMA for 1x1/s1.  B tile = BLOCK_N x BLOCK_K weights.
//             Both land in 128B/64B/32B-swizzled K-major smem, STAGES-deep mbarrier ring.
//   warp 1  : allocates TMEM; one lane issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into one
//             of two TMEM accumulator buffers; tcgen05.commit releases smem stages and
//             publishes the finished accumulator.
//   warps2-9: epilogue, two warps per TMEM lane quarter, each taking half of the tile's columns.
//             tcgen05.ld the accumulator (lane = output pixel), + folded-BN bias (staged in smem
//             once per tile), LeakyReLU, + residual, convert to bf16.  STAGED form (the common
//             one): each warp owns 32 rows x its column blocks of a swizzled smem slab; the shortcut
//             operand is TMA-loaded into the slab while the MMAs run, the result overwrites it in
//             place and every finished 64-column block leaves at once through a TMA store
//             (full-line writes, M tail clipped by the tensor map, ld_y pitch = channel slice of a
//             concat buffer).  DIRECT form (float32 head logits, fused 2x upsample): registers ->
//             st.global, one row per thread.
// The double-buffered accumulator lets tile i's epilogue overlap tile i+1's MMAs.  Weights do not
// depend on the previous layer: the producer requests the first ring pass of B tiles BEFORE the
// programmatic-dependent-launch wait, so they stream in while the previous kernel drains.
#include "common.cuh"
#define Y3_FILE_ID 1
#include "ptx.cuh"
#include "decode_math.cuh"

#include <cuda.h>  // CUtensorMap + enums only; entry points are fetched at run time
#include <string.h>

namespace y3 {

static constexpr int BLOCK_M = 128;
static constexpr int UMMA_K = 16;
static constexpr int NUM_THREADS = 320;  // 2 role warps + up to 8 epilogue warps (192 launched by default: 4)
static constexpr int EPI_WARP0 = 2;  // first epilogue warp

struct ConvKernelParams {
  int M;           // output pixels N*Ho*Wo
  int Ho, Wo, HoWo;
  int cin_blocks;  // Cin / BLOCK_K
  int num_kb;      // R*S*cin_blocks
  int S;           // filter width
  int stride, pad;
  int num_m_tiles, num_n_tiles;
  unsigned long long div_ntiles, div_howo, div_wo;  // ceil(2^48 / d): exact x / d for x * d < 2^48 (fast_div)
  int a_tiled;     // 1: A via 2-D tiled map (1x1, stride 1); 0: im2col map
  const float* bias;
  void* out;
  const __nv_bfloat16* res;
  int ld_out, ld_res;
  int leaky, out_f32, upsample;
  // DECODE epilogue (YOLO head: 3 anchors x (5 + 80) channels): candidates instead of logits
  float anchor_w[3], anchor_h[3];
  float train_w, train_h;
  float prob_thresh;
  const y3_thresholds* dyn;  // non-null: thresholds read from device memory at run time
  int box_offset;
  const int* orig_hw;
  uint4* cands;   // y3_cand records, [N][cap]
  int* counts;    // [N]
  int cap;
  int trace;       // diagnostics only (Y3_CONV_TRACE=1): CTA 0 records its pipeline timeline
};

// One anchor of the fused YOLO-head epilogue.  The thread's pixel has its 255 logits in TMEM lane
// `taddr`; anchor A's fields are columns [85 A, 85 A + 85).  Two passes over the same columns
// (TMEM re-reads are cheap): arg-max of the 80 class logits (first maximum wins, like torch.max),
// then the softmax denominator sum(exp(v - max)) in ascending class order.
template <int A>
__device__ __forceinline__ void decode_anchor(uint32_t taddr, const ConvKernelParams& p, float (&t)[5], float& sum,
                                              int& cls) {
  constexpr int BASE = 85 * A, C0 = BASE / 16, C1 = (BASE + 84) / 16;
  float best = -INFINITY;
  int best_idx = 0;
#pragma unroll
  for (int c = C0; c <= C1; ++c) {
    uint32_t v[16];
    ptx::tmem_ld_x16(taddr + 16 * c, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + 16 * c) + q);
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int f = 16 * c + 4 * q + e - BASE;  // compile-time
        if (f < 0 || f >= 85) continue;
        const float x = __uint_as_float(v[4 * q + e]) + bb[e];
        if (f < 5) t[f < 5 ? f : 0] = x;
        else if (x > best) { best = x; best_idx = f - 5; }
      }
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int c = C0; c <= C1; ++c) {
    uint32_t v[16];
    ptx::tmem_ld_x16(taddr + 16 * c, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + 16 * c) + q);
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int f = 16 * c + 4 * q + e - BASE;
        if (f < 5 || f >= 85) continue;
        // softmax denominator: exp(x - max) for x <= max through ex2.approx (2 + 1.16 |x - max| ulp): the terms
        // that matter (x near max) carry <= 4 ulp, the sum <= 3e-7 relative — inside the decode tolerance
        // (2e-6) and a third of this epilogue's instructions less than expf's argument reduction
        acc += __expf((__uint_as_float(v[4 * q + e]) + bb[e]) - best);
      }
    }
  }
  sum = acc;
  cls = best_idx;
}

// Threshold + box math + warp-aggregated append to the per-image candidate lists (lanes of one warp
// may sit in two images when the tile straddles an image boundary).
__device__ __forceinline__ void emit_cand(const ConvKernelParams& p, int a, bool valid, int img, int row, int col,
                                          const float (&t)[5], float sum, int cls, int lane) {
  // softmax value of the arg-max class is exp(0)/sum; then * sigmoid(objectness)  (darknet.py:104-108)
  const float prob = __fmul_rn(__fdiv_rn(1.0f, sum), sigmoidf_ref(t[4]));
  const float thresh = p.dyn ? __ldg(&p.dyn->prob_thresh) : p.prob_thresh;
  const bool pass = valid && prob >= thresh;  // inference.py:342
  uint32_t mask = __ballot_sync(0xffffffffu, pass);
  while (mask) {
    const int leader = __ffs(mask) - 1;
    const int li = __shfl_sync(0xffffffffu, img, leader);
    const uint32_t same = __ballot_sync(0xffffffffu, pass && img == li);
    int base = 0;
    if (lane == leader) base = atomicAdd(p.counts + li, __popc(same));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pass && img == li) {
      const int slot = base + __popc(same & ((1u << lane) - 1u));
      if (slot < p.cap) {
        uint4 lo, hi;
        const int cell = row * p.Wo + col;
        make_cand(t[0], t[1], t[2], t[3], prob, cls, p.box_offset + a * p.HoWo + cell, row, col, p.Ho, p.Wo,
                  p.anchor_w[a], p.anchor_h[a], p.train_w, p.train_h, (float)p.orig_hw[2 * img],
                  (float)p.orig_hw[2 * img + 1], lo, hi);
        uint4* dst = p.cands + 2 * ((long long)img * p.cap + slot);
        dst[0] = lo;
        dst[1] = hi;
      }
    }
    mask &= ~same;
  }
}

// CG = 1: one CTA per tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) shares one
// 256-row tile — each CTA stages its own 128 rows of A and HALF of the weight slab, the leader's
// MMAs read both halves, so per-SM operand traffic and smem per stage drop by a third and the
// pipeline gets deep enough to cover L2/HBM latency at full tensor rate.
// RES_KB > 0: the layer's whole weight slab (RES_KB k-blocks of this CTA's rows; one n tile per layer) is
// loaded ONCE per CTA and stays resident in shared memory; the ring then carries A only.  For layers with
// few k-blocks (1x1 over <= 256 channels, 3x3 over 64) the per-tile weight re-load is a third to a half of
// the L2 -> SM traffic that bounds them.
template <int BLOCK_N, int BLOCK_K, bool STAGED, int CG, int RES_KB = 0>
struct ConvCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_ROWS = BLOCK_N / CG;  // weight rows staged by THIS CTA
  static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
  // keep every stage 1024B-aligned (required for SWIZZLE_128B, harmless otherwise)
  static constexpr int A_STRIDE = (A_BYTES + 1023) / 1024 * 1024;
  static constexpr int B_STRIDE = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = RES_KB > 0 ? A_STRIDE : A_STRIDE + B_STRIDE;
  static constexpr int RES_BYTES = RES_KB * B_STRIDE;  // resident weights, after the ring
  // epilogue staging: 128 rows x BLOCK_N bf16, in column blocks of one swizzle span each
  static constexpr int EPI_SPAN = BLOCK_N * 2 >= 128 ? 128 : BLOCK_N * 2;  // bytes per staged row
  static constexpr int EPI_COLS = EPI_SPAN / 2;                            // columns per block
  static constexpr int EPI_BLOCKS = BLOCK_N / EPI_COLS;
  static constexpr int EPI_BLOCK_BYTES = BLOCK_M * EPI_SPAN;
  // two slabs when they are small, so tile i+1's epilogue does not wait for tile i's TMA store
  static constexpr int SLAB_BYTES = BLOCK_M * BLOCK_N * 2;
  static constexpr int STAGING_BUFS = BLOCK_N <= 128 ? 2 : 1;
  static constexpr int STAGING_BYTES = STAGED ? STAGING_BUFS * SLAB_BYTES : 0;
  // epilogue warps per TMEM lane quarter: 2 (each takes half of the columns) when the tile has at
  // least two staged column blocks (STAGED) or 32 columns (direct / decode forms), otherwise 1
  static constexpr int EPI_SPLIT = STAGED ? (EPI_BLOCKS >= 2 ? 2 : 1) : (BLOCK_N >= 32 ? 2 : 1);
  static constexpr int BIAS_BYTES = 2 * BLOCK_N * 4;  // this tile's bias, double-buffered
  static constexpr int BAR_BYTES = 512;
  // dynamic smem is declared __align__(1024): no alignment slack needed
  static constexpr int SMEM_LIMIT = 232448 - BAR_BYTES - BIAS_BYTES;
  static constexpr int STAGES_RAW = (SMEM_LIMIT - STAGING_BYTES - RES_BYTES) / STAGE_BYTES;
  static constexpr int MAX_STAGES = 16;  // barrier block: (2*STAGES + 14) * 8 + 8 bytes <= 512
  static constexpr int STAGES = STAGES_RAW > MAX_STAGES ? MAX_STAGES : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr int TMEM_COLS_RAW = 2 * BLOCK_N;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64
                                 : TMEM_COLS_RAW <= 128 ? 128 : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RES_BYTES + STAGING_BYTES + BAR_BYTES + BIAS_BYTES;
  // UMMA smem descriptor pieces (K-major, swizzle span = BLOCK_K*2 bytes)
  static constexpr uint64_t LAYOUT_TYPE = BLOCK_K == 64 ? 2 : BLOCK_K == 32 ? 4 : 6;
  static constexpr uint64_t SBO = 8 * BLOCK_K * 2;  // 8 rows of one swizzle atom
  static constexpr uint64_t DESC_HI = ((SBO >> 4) << 32) | (1ull << 46) | (LAYOUT_TYPE << 61);
  // instruction descriptor: D=f32, A=B=bf16, both K-major, N, M=128 per CTA (256 for a pair)
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) |
                                    (uint32_t(BLOCK_N >> 3) << 17) | (uint32_t((BLOCK_M * CG) >> 4) << 24);
};

// Diagnostics (Y3_CONV_TRACE=1): CTA 0 records clock64() at the pipeline events of its first tiles;
// read back with y3_debug_conv_trace (tools/conv_trace.py).  Slots: 0 entry, 1 set-up done, 2 PDL wait
// passed, 3 first TMA issued, 8+2i / 9+2i MMA warp: operands of tile i landed / accumulator committed,
// 32+4i.. epilogue warp 2: start, accumulator ready, drained, store issued; 80 stores drained, 81 exit.
__device__ unsigned long long g_conv_trace[96];
#define Y3_TRACE(slot) do { if (trace_on && (slot) < 96) g_conv_trace[(slot)] = clock64(); } while (0)

// x / d through the precomputed reciprocal m = ceil(2^48 / d) (host: div_magic); exact while x * d < 2^48.
__device__ __forceinline__ int fast_div(int x, unsigned long long m) {
  return (int)__umul64hi((unsigned long long)(unsigned)x << 16, m);
}
static unsigned long long div_magic(int d) { return ((1ull << 48) + (unsigned long long)d - 1) / (unsigned long long)d; }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint64_t desc_hi) {
  return desc_hi | (1ull << 16) | uint64_t((smem_addr >> 4) & 0x3FFFu);
}

template <int BLOCK_N, int BLOCK_K, bool STAGED, int CG, bool DECODE = false, int RES_KB = 0>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_r,
                 const ConvKernelParams p) {
  using Cfg = ConvCfg<BLOCK_N, BLOCK_K, STAGED, CG, RES_KB>;
  constexpr int STAGES = Cfg::STAGES;
  const uint32_t cta_rank = CG == 2 ? ptx::cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
  const int tile_first = blockIdx.x / CG;   // persistent loop over tiles, one CTA (pair) per SM (pair)
  const int tile_step = gridDim.x / CG;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();  // the swizzled operand tiles need 1024-byte alignment
  const uint32_t wres_base = smem_base + STAGES * Cfg::STAGE_BYTES;      // resident weights (RES_KB > 0), 1024B-aligned
  const uint32_t staging_base = wres_base + Cfg::RES_BYTES;              // 1024B-aligned
  const uint32_t bar_base = staging_base + Cfg::STAGING_BYTES;
  // barrier block: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], res[8], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  auto res_bar = [&](int q) { return bar_base + 8u * (2 * STAGES + 4 + q); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 12);
  const uint32_t wres_bar = bar_base + 8u * (2 * STAGES + 13);           // resident weights landed
  uint32_t* tmem_ptr_gen = reinterpret_cast<uint32_t*>(smem_raw + (tmem_ptr_addr - smem_base));
  const uint32_t bias_base = bar_base + Cfg::BAR_BYTES;  // float [2][BLOCK_N]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const bool trace_on = p.trace && blockIdx.x == 0 && lane == 0;
  // epilogue warps per TMEM lane quarter: launched with 10 warps -> the config's split, with 6 -> 1
  const int split = blockDim.x == NUM_THREADS ? Cfg::EPI_SPLIT : 1;
  if (warp == 0) Y3_TRACE(0);

  // PDL: the next kernel's CTAs may take this SM as soon as this CTA leaves; this CTA's own set-up
  // (barriers, TMEM, descriptor prefetch) overlaps the tail of the previous kernel.
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    if (STAGED) {
      ptx::prefetch_tensormap(&tmap_y);
      if (p.res) ptx::prefetch_tensormap(&tmap_r);
    }
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), CG);   // one producer arrival per CTA of the pair (leader's barrier)
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), 4 * split * CG);  // one arrival per working epilogue warp (of both CTAs)
    }
    for (int q = 0; q < 8; ++q) ptx::mbar_init(res_bar(q), 1);
    if (RES_KB > 0) ptx::mbar_init(wres_bar, CG);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<CG>(tmem_ptr_addr, Cfg::TMEM_COLS);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync();  // peer barriers must be initialised before any remote arrive
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  if (warp == 0) Y3_TRACE(1);
  // Programmatic dependent launch: only the threads that touch activations wait for the previous
  // kernel (producer before its first A load, epilogue warps before their first residual load /
  // store); weights and bias are constants and are requested before the wait.

  if (warp == 0) {
    // ===================== TMA producer =====================
    // One thread, and its instruction stream is on the critical path of every ring refill (a few
    // extra instructions per k-block cost 3-6 % on the deep layers — measured): keep the loop minimal.
    if (RES_KB > 0) {
      // ---- resident weights: every k-block of this CTA's rows once (constants: before the PDL wait),
      // then a ring of A tiles only
      if (lane == 0) {
        const uint32_t full_base = CG == 2 ? ptx::mapa(full_bar(0), 0) : full_bar(0);
        if (tile_first < num_tiles) {
          const uint32_t wbar = CG == 2 ? ptx::mapa(wres_bar, 0) : wres_bar;
          if (CG == 1 || cta_rank == 0) ptx::mbar_arrive_expect_tx(wres_bar, CG * RES_KB * Cfg::B_BYTES);
          else ptx::mbar_arrive_cluster(wbar);
          for (int kb = 0; kb < RES_KB; ++kb)
            ptx::tma_load_2d<CG>(wres_base + kb * Cfg::B_STRIDE, &tmap_b, wbar, kb * BLOCK_K, (int)cta_rank * Cfg::B_ROWS);
        }
        pdl_wait();
        Y3_TRACE(2);
        uint32_t phase = 0;
        uint32_t a_dst = smem_base, fbar = full_base, fbar_l = full_bar(0), ebar = empty_bar(0);
        const uint32_t fbar_end = full_bar(STAGES);
        const int cin = p.cin_blocks * BLOCK_K;
        for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
          const int m0 = (tile * CG + (int)cta_rank) * BLOCK_M;  // one n tile: tile == m tile
          const int img = fast_div(m0, p.div_howo);
          const int rem = m0 - img * p.HoWo;
          const int ho0 = fast_div(rem, p.div_wo);
          const int wo0 = rem - ho0 * p.Wo;
          const int w_base = wo0 * p.stride - p.pad;
          const int h_base = ho0 * p.stride - p.pad;
          int r = 0, s = 0, c = 0;
          for (int kb = 0; kb < RES_KB; ++kb) {
            ptx::mbar_wait(ebar, phase ^ 1u);
            if (CG == 1 || cta_rank == 0) ptx::mbar_arrive_expect_tx(fbar_l, CG * Cfg::A_BYTES);
            else ptx::mbar_arrive_cluster(fbar);
            if (p.a_tiled) ptx::tma_load_2d<CG>(a_dst, &tmap_a, fbar, c, m0);
            else ptx::tma_load_im2col_4d<CG>(a_dst, &tmap_a, fbar, c, w_base, h_base, img, (uint16_t)s, (uint16_t)r);
            c += BLOCK_K;
            if (c == cin) { c = 0; if (++s == p.S) { s = 0; ++r; } }
            a_dst += Cfg::STAGE_BYTES; fbar += 8; fbar_l += 8; ebar += 8;
            if (fbar_l == fbar_end) { a_dst = smem_base; fbar = full_base; fbar_l = full_bar(0); ebar = empty_bar(0); phase ^= 1u; }
          }
        }
      }
    } else
    if (lane == 0) {
      uint32_t phase = 0;
      // CG == 2: TMA completions of BOTH CTAs are counted on the leader's full barrier
      const uint32_t full_base = CG == 2 ? ptx::mapa(full_bar(0), 0) : full_bar(0);
      const int b_row0 = (int)cta_rank * Cfg::B_ROWS;
      // weights of the first ring pass do not depend on the previous layer: requested before the
      // programmatic-dependent-launch wait (the ring is empty, no empty-barrier wait needed)
      int pre = 0;
      if (tile_first < num_tiles) {
        pre = p.num_kb < STAGES ? p.num_kb : STAGES;
        const int n_tile = tile_first - fast_div(tile_first, p.div_ntiles) * p.num_n_tiles;
        for (int kb = 0; kb < pre; ++kb) {
          const uint32_t fbar = full_base + 8u * kb;
          if (CG == 1 || cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar(kb), CG * (Cfg::A_BYTES + Cfg::B_BYTES));
          else ptx::mbar_arrive_cluster(fbar);
          ptx::tma_load_2d<CG>(smem_base + kb * Cfg::STAGE_BYTES + Cfg::A_STRIDE, &tmap_b, fbar, kb * BLOCK_K,
                               n_tile * BLOCK_N + b_row0);
        }
      }
      pdl_wait();
      Y3_TRACE(2);
      // ring position as running registers: stage address, full barrier (leader's, cluster address),
      // own full barrier (expect_tx), empty barrier
      uint32_t a_dst = smem_base, fbar = full_base, fbar_l = full_bar(0), ebar = empty_bar(0);
      const uint32_t fbar_end = full_bar(STAGES);
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
        // three divisions per tile, each a 64-bit multiply (a true division costs the ring ~600 idle cycles)
        const int m_tile = fast_div(tile, p.div_ntiles);
        const int n_tile = tile - m_tile * p.num_n_tiles;
        const int m0 = (m_tile * CG + (int)cta_rank) * BLOCK_M;
        const int img = fast_div(m0, p.div_howo);
        const int rem = m0 - img * p.HoWo;
        const int ho0 = fast_div(rem, p.div_wo);
        const int wo0 = rem - ho0 * p.Wo;
        const int w_base = wo0 * p.stride - p.pad;
        const int h_base = ho0 * p.stride - p.pad;
        const int b_row = n_tile * BLOCK_N + b_row0;
        int r = 0, s = 0, c = 0, kb = 0;  // filter tap (r, s), first channel of the K block, K block
        const int cin = p.cin_blocks * BLOCK_K;
        if (tile == tile_first) {  // barriers armed and B in flight (above): only A is missing
          for (; kb < pre; ++kb) {
            if (p.a_tiled) ptx::tma_load_2d<CG>(a_dst, &tmap_a, fbar, c, m0);
            else ptx::tma_load_im2col_4d<CG>(a_dst, &tmap_a, fbar, c, w_base, h_base, img, (uint16_t)s, (uint16_t)r);
            c += BLOCK_K;
            if (c == cin) { c = 0; if (++s == p.S) { s = 0; ++r; } }
            a_dst += Cfg::STAGE_BYTES; fbar += 8; fbar_l += 8; ebar += 8;
            if (fbar_l == fbar_end) { a_dst = smem_base; fbar = full_base; fbar_l = full_bar(0); ebar = empty_bar(0); phase ^= 1u; }
          }
          Y3_TRACE(3);
        }
        int k0 = kb * BLOCK_K;  // K coordinate of the weight slab
        if (p.a_tiled) {
          for (; kb < p.num_kb; ++kb) {
            ptx::mbar_wait(ebar, phase ^ 1u);
            if (CG == 1 || cta_rank == 0) ptx::mbar_arrive_expect_tx(fbar_l, CG * (Cfg::A_BYTES + Cfg::B_BYTES));
            else ptx::mbar_arrive_cluster(fbar);
            ptx::tma_load_2d<CG>(a_dst, &tmap_a, fbar, k0, m0);
            ptx::tma_load_2d<CG>(a_dst + Cfg::A_STRIDE, &tmap_b, fbar, k0, b_row);
            k0 += BLOCK_K;
            a_dst += Cfg::STAGE_BYTES; fbar += 8; fbar_l += 8; ebar += 8;
            if (fbar_l == fbar_end) { a_dst = smem_base; fbar = full_base; fbar_l = full_bar(0); ebar = empty_bar(0); phase ^= 1u; }
          }
        } else {
          for (; kb < p.num_kb; ++kb) {
            ptx::mbar_wait(ebar, phase ^ 1u);
            if (CG == 1 || cta_rank == 0) ptx::mbar_arrive_expect_tx(fbar_l, CG * (Cfg::A_BYTES + Cfg::B_BYTES));
            else ptx::mbar_arrive_cluster(fbar);
            ptx::tma_load_im2col_4d<CG>(a_dst, &tmap_a, fbar, c, w_base, h_base, img, (uint16_t)s, (uint16_t)r);
            ptx::tma_load_2d<CG>(a_dst + Cfg::A_STRIDE, &tmap_b, fbar, k0, b_row);
            k0 += BLOCK_K;
            c += BLOCK_K;
            if (c == cin) { c = 0; if (++s == p.S) { s = 0; ++r; } }
            a_dst += Cfg::STAGE_BYTES; fbar += 8; fbar_l += 8; ebar += 8;
            if (fbar_l == fbar_end) { a_dst = smem_base; fbar = full_base; fbar_l = full_bar(0); ebar = empty_bar(0); phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // One thread; like the producer's, its instruction stream sits on the critical path whenever the
    // ring runs dry (operands land -> MMAs must be issued at once), so everything is a running register.
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t DESC_STEP = Cfg::STAGE_BYTES >> 4;  // smem descriptor address field counts 16-byte units
      const uint64_t desc_a0 = make_smem_desc(smem_base, Cfg::DESC_HI);
      const uint64_t desc_b0 = make_smem_desc(RES_KB > 0 ? wres_base : smem_base + Cfg::A_STRIDE, Cfg::DESC_HI);
      constexpr uint32_t RES_STEP = Cfg::B_STRIDE >> 4;
      uint64_t desc_a = desc_a0, desc_b = desc_b0;
      if (RES_KB > 0 && tile_first < num_tiles) {  // the resident weights (both CTAs' halves) have landed
        ptx::mbar_wait(wres_bar, 0);
        ptx::tc_fence_after();
      }
      uint32_t fbar = full_bar(0), ebar = empty_bar(0);
      const uint32_t fbar_end = full_bar(STAGES);
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this buffer
        Y3_TRACE(8 + 2 * it);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        uint32_t accumulate = 0;  // the tile's first MMA overwrites the accumulator
        if (RES_KB > 0) desc_b = desc_b0;  // k-block 0 of the resident slab
        for (int kb = p.num_kb; kb > 0; --kb) {
          ptx::mbar_wait(fbar, phase);  // TMA bytes have landed
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advancing K by 16 bf16 = 32 bytes inside the swizzle span: +2 in the >>4 address field
            ptx::umma_bf16_ss<CG>(tmem_d, desc_a + 2u * k, desc_b + 2u * k, Cfg::IDESC, accumulate);
            accumulate = 1;
          }
          ptx::umma_commit<CG>(ebar);  // smem stage reusable (in both CTAs) once these MMAs retire
          fbar += 8; ebar += 8; desc_a += DESC_STEP; desc_b += RES_KB > 0 ? RES_STEP : DESC_STEP;
          if (fbar == fbar_end) {
            fbar = full_bar(0); ebar = empty_bar(0); desc_a = desc_a0;
            if (RES_KB == 0) desc_b = desc_b0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit<CG>(tfull_bar(acc));  // accumulator complete (signalled in both CTAs)
        Y3_TRACE(9 + 2 * it);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quarter = warp & 3;               // TMEM lanes [32*quarter, 32*quarter+32) are this warp's
    const int half = (warp - EPI_WARP0) >> 2;   // which half of the tile's columns (warps 2-5: 0, 6-9: 1)
    const int row = quarter * 32 + lane;
    const int e_tid = (warp - EPI_WARP0) * 32 + lane;
    // the accumulator is handed back on the LEADER's barrier (its MMA thread waits there)
    const uint32_t tempty_base = (CG == 2 && cta_rank != 0) ? ptx::mapa(tempty_bar(0), 0) : tempty_bar(0);
    if (half < split) {
    int it = 0;
    bool waited = false;  // griddepcontrol.wait executed (before the first access to activations)
    for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_tile = fast_div(tile, p.div_ntiles);
      const int n_tile = tile - m_tile * p.num_n_tiles;
      const int m0 = (m_tile * CG + (int)cta_rank) * BLOCK_M;
      const int m = m0 + row;
      const int n0 = n_tile * BLOCK_N;
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BLOCK_N;

      if constexpr (STAGED) {
        // ---- staged: smem slab (swizzled like the TMA box) -> TMA store ----
        // slab row r of column block cb lives at staging + cb*EPI_BLOCK_BYTES + r*EPI_SPAN; its
        // 16-byte units are XOR-swizzled exactly as CU_TENSOR_MAP_SWIZZLE_{128,64,32}B does.
        constexpr int SPAN = Cfg::EPI_SPAN;
        const int CB_PER = Cfg::EPI_BLOCKS / split;  // column blocks of this warp
        const int cb_first = half * CB_PER;
        const uint32_t swz = SPAN == 128 ? (row & 7) : SPAN == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1);
        const uint32_t slab = staging_base + (Cfg::STAGING_BUFS == 2 ? (it & 1) * Cfg::SLAB_BYTES : 0);
        const uint32_t row_base = slab + row * SPAN;
        const uint32_t warp_base = slab + quarter * 32 * SPAN;
        const uint32_t bias_s = bias_base + (it & 1) * (BLOCK_N * 4);
        const bool has_res = p.res != nullptr;
        if (warp == EPI_WARP0) Y3_TRACE(32 + 4 * it);
        // this tile's bias: global -> registers now, -> smem once the previous readers are past it
        constexpr int BIAS_PER = (BLOCK_N + 127) / 128;  // values per thread with one warp per quarter
        float bias_v[BIAS_PER];
#pragma unroll
        for (int j = 0; j < BIAS_PER; ++j) {
          const int c = e_tid + j * split * 128;
          bias_v[j] = c < BLOCK_N ? __ldg(p.bias + n0 + c) : 0.f;
        }
        if (!waited) { pdl_wait(); waited = true; }
        if (lane == 0) {
          // the TMA stores that last used this slab must have finished READING it before reuse
          if (Cfg::STAGING_BUFS == 2 && CB_PER == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else if (Cfg::STAGING_BUFS == 2 && CB_PER == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (has_res) {
            ptx::mbar_arrive_expect_tx(res_bar(warp - EPI_WARP0), 32 * CB_PER * Cfg::EPI_COLS * 2);
            for (int cbi = 0; cbi < CB_PER; ++cbi)
              ptx::tma_load_2d(warp_base + (cb_first + cbi) * Cfg::EPI_BLOCK_BYTES, &tmap_r, res_bar(warp - EPI_WARP0),
                               n0 + (cb_first + cbi) * Cfg::EPI_COLS, m0 + quarter * 32);
          }
        }
#pragma unroll
        for (int j = 0; j < BIAS_PER; ++j) {
          const int c = e_tid + j * split * 128;
          if (c < BLOCK_N) asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4u * c), "f"(bias_v[j]) : "memory");
        }
        ptx::named_bar_sync(1, split * 128);  // bias visible to every epilogue warp (double-buffered by tile parity)
        ptx::mbar_wait(tfull_bar(acc), acc_phase);
        if (warp == EPI_WARP0) Y3_TRACE(33 + 4 * it);
        ptx::tc_fence_after();
        if (has_res) ptx::mbar_wait(res_bar(warp - EPI_WARP0), it & 1);

#pragma unroll 1
        for (int cbi = 0; cbi < CB_PER; ++cbi) {
          const int cb = cb_first + cbi;
#pragma unroll 1
          for (int cc = 0; cc < Cfg::EPI_COLS; cc += 16) {
            const int c0 = cb * Cfg::EPI_COLS + cc;
            uint32_t v[16];
            ptx::tmem_ld_x16(taddr + c0, v);
            float bz[16];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bz[4 * q]), "=f"(bz[4 * q + 1]), "=f"(bz[4 * q + 2]),
                           "=f"(bz[4 * q + 3]) : "r"(bias_s + 4u * (c0 + 4 * q)));
            ptx::tmem_ld_wait();
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + bz[j];
            if (p.leaky) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.1f * f[j]);
            }
            const uint32_t u0 = uint32_t(cc >> 3);  // 16-byte unit index in the row
            const uint32_t a0 = row_base + cb * Cfg::EPI_BLOCK_BYTES + ((u0 ^ swz) << 4);
            const uint32_t a1 = row_base + cb * Cfg::EPI_BLOCK_BYTES + (((u0 + 1) ^ swz) << 4);
            if (has_res) {
              uint4 r0, r1;
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(a0));
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(a1));
              const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 t = unpack_bf16x2(rr[j]);
                f[2 * j] += t.x;
                f[2 * j + 1] += t.y;
              }
            }
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(pack_bf16x2(f[0], f[1])),
                         "r"(pack_bf16x2(f[2], f[3])), "r"(pack_bf16x2(f[4], f[5])), "r"(pack_bf16x2(f[6], f[7])) : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a1), "r"(pack_bf16x2(f[8], f[9])),
                         "r"(pack_bf16x2(f[10], f[11])), "r"(pack_bf16x2(f[12], f[13])), "r"(pack_bf16x2(f[14], f[15])) : "memory");
          }
          // this column block is complete: publish it to the async proxy and store it right away
          if (cbi == CB_PER - 1) ptx::tc_fence_before();  // (all TMEM reads of this warp are done)
          ptx::fence_proxy_async();
          __syncwarp();
          if (cbi == CB_PER - 1 && warp == EPI_WARP0) Y3_TRACE(34 + 4 * it);
          if (lane == 0) {
            if (cbi == CB_PER - 1) ptx::mbar_arrive_cluster(tempty_base + 8u * acc);  // hand the accumulator back
            if (m0 + quarter * 32 < p.M)
              ptx::tma_store_2d(&tmap_y, warp_base + cb * Cfg::EPI_BLOCK_BYTES, n0 + cb * Cfg::EPI_COLS, m0 + quarter * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (warp == EPI_WARP0) Y3_TRACE(35 + 4 * it);
      } else if constexpr (DECODE) {
        // ---- YOLO head: decode the pixel's three anchors straight from the accumulator ----
        // The two warps of a lane quarter share a pixel's anchors 2:1, alternating with the tile parity
        // (tile i: warps 2-5 take anchors 0,1 and warps 6-9 anchor 2; tile i+1 the other way round), so
        // with the double-buffered accumulator both halves carry the same load.
        const bool valid = m < p.M;
        const int mm = valid ? m : 0;
        const int img = mm / p.HoWo;
        const int rem = mm - img * p.HoWo;
        const int grow = rem / p.Wo;
        const int gcol = rem - grow * p.Wo;
        if (!waited) { pdl_wait(); waited = true; }
        ptx::mbar_wait(tfull_bar(acc), acc_phase);
        ptx::tc_fence_after();
        float t[5], sum;
        int cls;
        const bool two = split == 1 || ((it & 1) == half);  // this warp decodes two anchors of this tile
        if (half == 0) {
          decode_anchor<0>(taddr, p, t, sum, cls);
          emit_cand(p, 0, valid, img, grow, gcol, t, sum, cls, lane);
          if (two) {
            decode_anchor<1>(taddr, p, t, sum, cls);
            emit_cand(p, 1, valid, img, grow, gcol, t, sum, cls, lane);
          }
          if (split == 1) {
            decode_anchor<2>(taddr, p, t, sum, cls);
            emit_cand(p, 2, valid, img, grow, gcol, t, sum, cls, lane);
          }
        } else {
          if (two) {
            decode_anchor<1>(taddr, p, t, sum, cls);
            emit_cand(p, 1, valid, img, grow, gcol, t, sum, cls, lane);
          }
          decode_anchor<2>(taddr, p, t, sum, cls);
          emit_cand(p, 2, valid, img, grow, gcol, t, sum, cls, lane);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_base + 8u * acc);
      } else {
        // ---- direct: registers -> global, one output row per thread, half of the columns per warp ----
        const int COLS = BLOCK_N / split;
        const int c_first = half * COLS;
        const bool valid = m < p.M;
        long long dst_pix = m;
        int up_w2 = 0;
        if (p.upsample) {
          const int img = m / p.HoWo;
          const int rem = m - img * p.HoWo;
          const int ho = rem / p.Wo;
          const int wo = rem - ho * p.Wo;
          up_w2 = 2 * p.Wo;
          dst_pix = ((long long)img * (2 * p.Ho) + 2 * ho) * up_w2 + 2 * wo;
        }
        const __nv_bfloat16* res_row = p.res ? p.res + (long long)m * p.ld_res + n0 : nullptr;

        if (!waited) { pdl_wait(); waited = true; }
        ptx::mbar_wait(tfull_bar(acc), acc_phase);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c0 = c_first; c0 < c_first + COLS; c0 += 16) {
          uint32_t v[16];
          ptx::tmem_ld_x16(taddr + c0, v);
          ptx::tmem_ld_wait();
          float f[16];
          const float4* bias4 = reinterpret_cast<const float4*>(p.bias + n0 + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b = __ldg(bias4 + q);
            f[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + b.x;
            f[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + b.y;
            f[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + b.z;
            f[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + b.w;
          }
          if (p.leaky) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = f[j] > 0.f ? f[j] : 0.1f * f[j];
          }
          if (valid) {
            if (res_row) {
              const uint4 r0 = ld_nc_16(res_row + c0);
              const uint4 r1 = ld_nc_16(res_row + c0 + 8);
              const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 t = unpack_bf16x2(rr[j]);
                f[2 * j] += t.x;
                f[2 * j + 1] += t.y;
              }
            }
            if (p.out_f32) {
              float* o = reinterpret_cast<float*>(p.out) + dst_pix * p.ld_out + n0 + c0;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                st_16(o + 4 * q, make_uint4(__float_as_uint(f[4 * q]), __float_as_uint(f[4 * q + 1]),
                                            __float_as_uint(f[4 * q + 2]), __float_as_uint(f[4 * q + 3])));
            } else {
              const uint4 o0 = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                          pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
              const uint4 o1 = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]),
                                          pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + dst_pix * p.ld_out + n0 + c0;
              st_16(o, o0);
              st_16(o + 8, o1);
              if (p.upsample) {
                __nv_bfloat16* o01 = o + p.ld_out;
                __nv_bfloat16* o10 = o + (long long)up_w2 * p.ld_out;
                __nv_bfloat16* o11 = o10 + p.ld_out;
                st_16(o01, o0); st_16(o01 + 8, o1);
                st_16(o10, o0); st_16(o10 + 8, o1);
                st_16(o11, o0); st_16(o11 + 8, o1);
              }
            }
          }
        }
        // all TMEM reads of this warp are complete (wait::ld above): hand the buffer back
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_base + 8u * acc);
      }
    }
    // smem must stay valid until the bulk stores have read it; their global writes complete with the grid
    if (STAGED && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (warp == EPI_WARP0) Y3_TRACE(80);
  }

  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync();  // the peer may still be reading this CTA's smem / signalling it
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
  }
  if (warp == 0) Y3_TRACE(81);
}

// ---------------------------------------------------------------------------------------
// Host side: tensor-map encoding (driver entry points resolved through the runtime so the
// library has no link-time dependency on libcuda and loads on a CPU-only box).
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static int resolve_driver_entry_points() {
  if (g_encode_tiled && g_encode_im2col) return Y3_OK;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  Y3_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return Y3_ECUDA;
  }
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  Y3_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeIm2col not available from the driver");
    return Y3_ECUDA;
  }
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  Y3_CUDA_OK(cudaDriverGetVersion(&g_driver_version));
  return Y3_OK;
}

static CUtensorMapSwizzle swizzle_for(int block_k) {
  return block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
       : block_k == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

static int encode_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                     uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer, int block_k) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(block_k),
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): dims=[%llu,%llu] pitch=%llu box=[%u,%u]",
              (int)r, (unsigned long long)inner, (unsigned long long)outer,
              (unsigned long long)pitch_bytes, box_inner, box_outer);
    return Y3_ECUDA;
  }
  return Y3_OK;
}

int encode_tiled_map(void* map, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes) {
  int rc = resolve_driver_entry_points();
  if (rc != Y3_OK) return rc;
  Y3_CHECK_ARG(rank >= 2 && rank <= 5, "tensor map rank %d", rank);
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = g_encode_tiled(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                              const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle_for(swizzle_bytes / 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (rank %d) failed (CUresult %d): dims0=%llu box0=%u stride0=%llu", rank, (int)r,
              (unsigned long long)dims[0], box[0], (unsigned long long)strides_bytes[0]);
    return Y3_ECUDA;
  }
  return Y3_OK;
}

static int encode_im2col(CUtensorMap* map, const y3_conv_desc* d, const void* x, int block_k) {
  cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
  const uint64_t pix = (uint64_t)d->ld_x * 2;
  cuuint64_t strides[3] = {pix, pix * d->w, pix * d->w * d->h};
  // Bounding box of filter-window origins in input coordinates:
  // from -pad to (extent-1) + pad - (ksize-1)  (see SURVEY.md §7 step 3).
  int lower[2] = {-d->pad, -d->pad};
  int upper[2] = {d->pad - (d->ksize - 1), d->pad - (d->ksize - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
  CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims,
                               strides, lower, upper, (cuuint32_t)block_k, (cuuint32_t)BLOCK_M, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(block_k),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed (CUresult %d): dims=[%d,%d,%d,%d] ld=%d k=%d s=%d p=%d",
              (int)r, d->cin, d->w, d->h, d->n, d->ld_x, d->ksize, d->stride, d->pad);
    return Y3_ECUDA;
  }
  // Drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KiB
  // (descriptor word 1, bit 21 must be clear); same fix-up NVIDIA's own templates apply.
  if (g_driver_version <= 13010) {
    const uint64_t extent = pix * d->w * d->h * (uint64_t)d->n;
    if (extent < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  }
  return Y3_OK;
}

struct DecodeArgs {
  const y3_head_desc* head;
  float prob_thresh;
  const y3_thresholds* dyn;
  const int* orig_hw;
  void* cands;
  int* counts;
  int cap;
};

template <int BLOCK_N, int BLOCK_K, bool STAGED, int CG, bool DECODE = false, int RES_KB = 0>
static int launch_conv(const y3_conv_desc* d, const void* x, const void* w, const float* bias,
                       const void* residual, void* y, cudaStream_t stream, int force_im2col,
                       const DecodeArgs* dec = nullptr) {
  using Cfg = ConvCfg<BLOCK_N, BLOCK_K, STAGED, CG, RES_KB>;
  static_assert(RES_KB == 0 || Cfg::STAGES >= 4, "resident weights leave too few A stages");
  const int ho = (d->h + 2 * d->pad - d->ksize) / d->stride + 1;
  const int wo = (d->w + 2 * d->pad - d->ksize) / d->stride + 1;
  const long long M = (long long)d->n * ho * wo;
  Y3_CHECK_ARG(M > 0 && M < (1ll << 31) - BLOCK_M, "conv: M=%lld out of range", M);

  ConvKernelParams p;
  p.M = (int)M;
  p.Ho = ho; p.Wo = wo; p.HoWo = ho * wo;
  p.cin_blocks = d->cin / BLOCK_K;
  p.num_kb = d->ksize * d->ksize * p.cin_blocks;
  p.S = d->ksize;
  p.stride = d->stride; p.pad = d->pad;
  p.num_m_tiles = (int)((M + BLOCK_M * CG - 1) / (BLOCK_M * CG));
  p.num_n_tiles = d->cout / BLOCK_N;
  p.div_ntiles = div_magic(p.num_n_tiles);
  p.div_howo = div_magic(p.HoWo);
  p.div_wo = div_magic(p.Wo);
  p.a_tiled = (d->ksize == 1 && d->stride == 1 && d->pad == 0 && !force_im2col) ? 1 : 0;
  p.bias = bias;
  p.out = y;
  p.res = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.ld_out = d->ld_y; p.ld_res = d->ld_res;
  p.leaky = d->leaky; p.out_f32 = d->out_f32; p.upsample = d->upsample2x;
  p.orig_hw = nullptr; p.cands = nullptr; p.counts = nullptr; p.cap = 0;
  {
    static int trace = -1;
    if (trace < 0) { const char* e = getenv("Y3_CONV_TRACE"); trace = (e && e[0] == '1') ? 1 : 0; }
    p.trace = trace;
  }
  p.train_w = p.train_h = 1.f; p.prob_thresh = 0.f; p.dyn = nullptr; p.box_offset = 0;
  for (int a = 0; a < 3; ++a) p.anchor_w[a] = p.anchor_h[a] = 0.f;
  if (DECODE) {
    for (int a = 0; a < 3; ++a) { p.anchor_w[a] = dec->head->anchor_w[a]; p.anchor_h[a] = dec->head->anchor_h[a]; }
    p.train_w = dec->head->train_w; p.train_h = dec->head->train_h;
    p.prob_thresh = dec->prob_thresh;
    p.dyn = dec->dyn;
    p.box_offset = dec->head->box_offset;
    p.orig_hw = dec->orig_hw;
    p.cands = reinterpret_cast<uint4*>(dec->cands);
    p.counts = dec->counts;
    p.cap = dec->cap;
  }

  int rc = resolve_driver_entry_points();
  if (rc != Y3_OK) return rc;

  alignas(64) CUtensorMap tmap_a, tmap_b, tmap_y, tmap_r;
  memset(&tmap_y, 0, sizeof(tmap_y));
  memset(&tmap_r, 0, sizeof(tmap_r));
  if (p.a_tiled) {
    rc = encode_2d(&tmap_a, x, (uint64_t)d->cin, (uint64_t)d->n * d->h * d->w, (uint64_t)d->ld_x * 2,
                   BLOCK_K, BLOCK_M, BLOCK_K);
  } else {
    rc = encode_im2col(&tmap_a, d, x, BLOCK_K);
  }
  if (rc != Y3_OK) return rc;
  const uint64_t k_total = (uint64_t)d->ksize * d->ksize * d->cin;
  rc = encode_2d(&tmap_b, w, k_total, (uint64_t)d->cout, k_total * 2, BLOCK_K, Cfg::B_ROWS, BLOCK_K);
  if (rc != Y3_OK) return rc;

  if (STAGED) {
    // output / residual tiles: [32 rows x EPI_COLS] boxes, swizzle span = EPI_SPAN bytes
    const int span_k = Cfg::EPI_SPAN / 2;  // "block_k" equivalent that selects the same swizzle mode
    rc = encode_2d(&tmap_y, y, (uint64_t)d->cout, (uint64_t)M, (uint64_t)d->ld_y * 2, Cfg::EPI_COLS, 32, span_k);
    if (rc != Y3_OK) return rc;
    if (residual) {
      rc = encode_2d(&tmap_r, residual, (uint64_t)d->cout, (uint64_t)M, (uint64_t)d->ld_res * 2, Cfg::EPI_COLS, 32,
                     span_k);
      if (rc != Y3_OK) return rc;
    }
  }

  auto kernel = conv_umma_kernel<BLOCK_N, BLOCK_K, STAGED, CG, DECODE, RES_KB>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    Y3_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int slots = num_sms() / CG;  // CTAs (or CTA pairs) resident at once
  const int grid = CG * (num_tiles < slots ? num_tiles : slots);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  // Epilogue warps per TMEM lane quarter (Y3_EPI_WARPS=4|8 forces one or two everywhere).  Measured on
  // yolov3-416 x 64: two warps speed the decode / direct epilogues up by 10-20 % (register-heavy, long
  // drains) and are neutral to slightly negative for the staged one on the load-bound 3x3 layers.
  static int epi_warps = -1;
  if (epi_warps < 0) { const char* e = getenv("Y3_EPI_WARPS"); epi_warps = e ? atoi(e) : 0; }
  // default: the register-heavy direct / decode epilogues run two warps per quarter, the staged one one
  const bool eight = epi_warps == 8 || (epi_warps != 4 && !STAGED);
  cfg.blockDim = dim3(eight ? NUM_THREADS : 192);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  Y3_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, tmap_a, tmap_b, tmap_y, tmap_r, p));
  Y3_LAUNCH_OK("conv_umma_kernel");
  return Y3_OK;
}

template <int BLOCK_N, int BLOCK_K>
static int dispatch_epi(const y3_conv_desc* d, const void* x, const void* w, const float* bias,
                        const void* residual, void* y, cudaStream_t stream, int force_im2col) {
  // float32 head logits and the fused 2x upsample use the direct (register -> global) epilogue
  if (d->out_f32 || d->upsample2x || (d->flags & 2))
    return launch_conv<BLOCK_N, BLOCK_K, false, 1>(d, x, w, bias, residual, y, stream, force_im2col);
  // 128- and 256-channel tiles: CTA pairs (cta_group::2) — each CTA stages only half of the weight slab,
  // which is what the L2 -> SM path limits.  flags bit2 forces single-CTA tiles.
  // layers with ONE n tile and few k-blocks keep their weights resident in shared memory (RES_KB):
  // 1x1 128->64 (2 k-blocks), 1x1 256->128 (4), 3x3 64->128 (9).  Y3_NO_BRES=1 streams them as before.
  static int bres = -1;
  if (bres < 0) { const char* e = getenv("Y3_NO_BRES"); bres = (e && e[0] == '1') ? 0 : 1; }
  const int num_kb = d->ksize * d->ksize * (d->cin / BLOCK_K);
  const bool one_n_tile = bres && d->cout == BLOCK_N && !(d->flags & 8);
  if constexpr ((BLOCK_N == 256 || BLOCK_N == 128) && BLOCK_K == 64) {
    const long long M = (long long)d->n * ((d->h + 2 * d->pad - d->ksize) / d->stride + 1) *
                        ((d->w + 2 * d->pad - d->ksize) / d->stride + 1);
    if (!(d->flags & 4) && M >= 2 * BLOCK_M) {
      if constexpr (BLOCK_N == 128) {
        if (one_n_tile && num_kb == 9)
          return launch_conv<BLOCK_N, BLOCK_K, true, 2, false, 9>(d, x, w, bias, residual, y, stream, force_im2col);
        if (one_n_tile && num_kb == 4)
          return launch_conv<BLOCK_N, BLOCK_K, true, 2, false, 4>(d, x, w, bias, residual, y, stream, force_im2col);
      }
      return launch_conv<BLOCK_N, BLOCK_K, true, 2>(d, x, w, bias, residual, y, stream, force_im2col);
    }
  }
  if constexpr (BLOCK_N == 64 && BLOCK_K == 64) {
    if (one_n_tile && num_kb == 2)
      return launch_conv<BLOCK_N, BLOCK_K, true, 1, false, 2>(d, x, w, bias, residual, y, stream, force_im2col);
  }
  return launch_conv<BLOCK_N, BLOCK_K, true, 1>(d, x, w, bias, residual, y, stream, force_im2col);
}

template <int BLOCK_K>
static int dispatch_n(const y3_conv_desc* d, const void* x, const void* w, const float* bias,
                      const void* residual, void* y, cudaStream_t stream, int force_im2col) {
  const int c = d->cout;
  if (c % 256 == 0) return dispatch_epi<256, BLOCK_K>(d, x, w, bias, residual, y, stream, force_im2col);
  if (c % 128 == 0) return dispatch_epi<128, BLOCK_K>(d, x, w, bias, residual, y, stream, force_im2col);
  if (c % 64 == 0) return dispatch_epi<64, BLOCK_K>(d, x, w, bias, residual, y, stream, force_im2col);
  if (c % 32 == 0) return dispatch_epi<32, BLOCK_K>(d, x, w, bias, residual, y, stream, force_im2col);
  return dispatch_epi<16, BLOCK_K>(d, x, w, bias, residual, y, stream, force_im2col);
}

static int conv2d_impl(const y3_conv_desc* d, const void* x, const void* w, const float* bias,
                       const void* residual, void* y, void* stream, int force_im2col) {
  Y3_CHECK_ARG(d && x && w && bias && y, "conv: null argument");
  Y3_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0, "conv: bad input shape %dx%dx%d", d->n, d->h, d->w);
  Y3_CHECK_ARG(d->cin > 0 && d->cin % 16 == 0, "conv: cin=%d must be a positive multiple of 16", d->cin);
  Y3_CHECK_ARG(d->cout > 0 && d->cout % 16 == 0, "conv: cout=%d must be a positive multiple of 16", d->cout);
  Y3_CHECK_ARG(d->ksize == 1 || d->ksize == 3, "conv: ksize=%d unsupported (1 or 3)", d->ksize);
  Y3_CHECK_ARG(d->stride == 1 || d->stride == 2, "conv: stride=%d unsupported (1 or 2)", d->stride);
  Y3_CHECK_ARG(d->pad == 0 || d->pad == (d->ksize - 1) / 2, "conv: pad=%d unsupported", d->pad);
  Y3_CHECK_ARG(d->h + 2 * d->pad >= d->ksize && d->w + 2 * d->pad >= d->ksize, "conv: input smaller than filter");
  Y3_CHECK_ARG(d->ld_x >= d->cin && d->ld_x % 8 == 0, "conv: ld_x=%d must be >= cin and a multiple of 8", d->ld_x);
  Y3_CHECK_ARG(d->ld_y >= d->cout && d->ld_y % (d->out_f32 ? 4 : 8) == 0, "conv: ld_y=%d invalid", d->ld_y);
  Y3_CHECK_ARG(!residual || (d->ld_res >= d->cout && d->ld_res % 8 == 0), "conv: ld_res=%d invalid", d->ld_res);
  Y3_CHECK_ARG(!(d->upsample2x && d->out_f32), "conv: upsample2x with out_f32 unsupported");
  Y3_CHECK_ARG(!(d->upsample2x && residual), "conv: upsample2x with residual unsupported");
  Y3_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
               "conv: pointers must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (!force_im2col) {  // 3x3 / 1 layers on large feature maps: input patch in smem, nine taps = nine descriptor offsets
    const int rc = conv3x3_patch_try(d, x, w, bias, residual, y, s);
    if (rc >= 0) return rc;
  }
  if (d->cin % 64 == 0) return dispatch_n<64>(d, x, w, bias, residual, y, s, force_im2col);
  if (d->cin % 32 == 0) return dispatch_n<32>(d, x, w, bias, residual, y, s, force_im2col);
  return dispatch_n<16>(d, x, w, bias, residual, y, s, force_im2col);
}

}  // namespace y3

// Diagnostics (declared at the end of include/yolov3_b200.h): copy the trace of the last traced launch to the host.
extern "C" int y3_debug_conv_trace(unsigned long long* out, int n) {
  using namespace y3;
  Y3_CHECK_ARG(out && n > 0 && n <= 96, "debug_conv_trace: n=%d", n);
  Y3_CUDA_OK(cudaDeviceSynchronize());
  Y3_CUDA_OK(cudaMemcpyFromSymbol(out, g_conv_trace, sizeof(unsigned long long) * n));
  return Y3_OK;
}

extern "C" int y3_conv2d_yolo_head(const y3_conv_desc* d, const void* x, const void* w, const float* bias,
                                   const y3_head_desc* head, float prob_thresh,
                                   const y3_thresholds* dev_thresholds, const int32_t* orig_hw,
                                   y3_cand* cands, int32_t* counts, int32_t cap, void* stream) {
  using namespace y3;
  Y3_CHECK_ARG(d && x && w && bias && head && orig_hw && cands && counts && cap > 0, "conv2d_yolo_head: null argument");
  Y3_CHECK_ARG(d->ksize == 1 && d->stride == 1 && d->pad == 0, "conv2d_yolo_head: the head convolution must be 1x1/1");
  Y3_CHECK_ARG(d->cout == 256 && d->cin % 64 == 0 && d->ld_x >= d->cin && d->ld_x % 8 == 0,
               "conv2d_yolo_head: cout=%d (256 stored channels required), cin=%d (multiple of 64 required)", d->cout,
               d->cin);
  Y3_CHECK_ARG(head->num_anchors == 3 && head->num_classes == 80,
               "conv2d_yolo_head: %d anchors x %d classes (3 x 80 required; use y3_conv2d + y3_yolo_decode_cands)",
               head->num_anchors, head->num_classes);
  Y3_CHECK_ARG(head->n == d->n && head->g_h == d->h && head->g_w == d->w, "conv2d_yolo_head: head / conv shapes differ");
  Y3_CHECK_ARG(!d->leaky && !d->upsample2x, "conv2d_yolo_head: the head convolution is linear");
  Y3_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(cands) & 15) == 0,
               "conv2d_yolo_head: pointers must be 16-byte aligned");
  DecodeArgs dec = {head, prob_thresh, dev_thresholds, orig_hw, cands, counts, cap};
  return launch_conv<256, 64, false, 1, true>(d, x, w, bias, nullptr, const_cast<float*>(bias) /*unused*/,
                                              reinterpret_cast<cudaStream_t>(stream), 0, &dec);
}

extern "C" int y3_conv2d(const y3_conv_desc* d, const void* x, const void* w, const float* bias,
                         const void* residual, void* y, void* stream) {
  return y3::conv2d_impl(d, x, w, bias, residual, y, stream, d ? (d->flags & 1) : 0);
}

// y3_debug_set_trap_record (api.cu): this translation unit's copy of the watchdog record pointer
namespace y3 { cudaError_t conv_umma_set_trap_record(unsigned long long* host_mapped) { return ptx::set_trap_record_tu(host_mapped); } }
