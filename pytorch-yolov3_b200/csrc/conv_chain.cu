// conv_chain.cu — two chained Darknet [convolutional] blocks in ONE kernel, the intermediate
// activation never leaving the SM (sm_100a: TMA + tcgen05.mma + TMEM).
//
// The first layers of Darknet-53 are HBM-bound: 5 % of the FLOPs but a third of the forward time
// when every block round-trips its 416^2 / 208^2 activation through HBM.  Two shapes are fused:
//
//   STEM  uint8 BGR image -> [conv 3x3/1 3->32 +BN+leaky] -> [conv 3x3/2 32->64 +BN+leaky]
//         (yolov3/inference.py:332-333 + blocks 0,1 of models/yolov3.cfg; darknet.py:244-257)
//   RES   x(64ch) -> [conv 1x1 64->32 +BN+leaky] -> [conv 3x3/1 32->64 +BN+leaky] + x
//         (one Darknet residual unit incl. its [shortcut], darknet.py:376-379)
//
// Per CTA: TEAMS independent teams of 4 warps, each looping over 8(w) x 16(h) output tiles:
//   stage 1  STEM: one 32-byte A row per image-patch pixel (its BGR bytes and its two right neighbours'),
//            the three filter rows = three MMAs over the same rows shifted by one patch row each;
//            RES: the x patch with halo fetched by one 4-D TMA.  D1[pixel of the patch][32] in TMEM
//   epi 1    tcgen05.ld -> +bias, leaky, ZERO outside the image (stage 2's padding) -> bf16
//            patch in smem, laid out (swizzled) exactly as a K-major UMMA operand
//   stage 2  nine taps = nine descriptor START OFFSETS into that one patch (row stride of the
//            8-pixel groups through SBO; stride 2 by viewing pixel pairs as 128-byte rows)
//   epi 2    +bias, leaky, (+x from the patch already in smem) -> swizzled slab -> 4-D TMA store
// Weights of both layers stay resident in smem for the whole kernel.  Teams are not pipelined
// internally; three teams per SM overlap each other's phases.
#include "common.cuh"
#define Y3_FILE_ID 3
#include "ptx.cuh"

#include <cuda.h>
#include <string.h>

namespace y3 {

enum { CHAIN_STEM = 0, CHAIN_RES = 1 };

constexpr int round_up_c(int x, int m) { return (x + m - 1) / m * m; }

template <int MODE>
struct ChainCfg {
  static constexpr bool STEM = MODE == CHAIN_STEM;
  static constexpr int TEAMS = 3;
  static constexpr int TEAM_THREADS = 128;
  static constexpr int THREADS = TEAMS * TEAM_THREADS;
  static constexpr int TW = 8, TH = 16;             // stage-2 output tile (M = 128 rows = 16 groups of 8)
  static constexpr int S2 = STEM ? 2 : 1;           // stage-2 stride
  static constexpr int CM = 32, C2 = 64;            // intermediate / output channels
  static constexpr int PH = (TH - 1) * S2 + 3;      // patch rows of the intermediate: 33 | 18
  static constexpr int PWM = STEM ? 18 : 10;        // patch pitch in pixels (even for pixel pairs)
  static constexpr int NP = PH * PWM;               // 594 | 180 patch pixels = stage-1 GEMM rows
  static constexpr int MT1 = (NP + 127) / 128;      // 5 | 2 stage-1 M tiles
  // Stage-1 A rows.  STEM: one 32-byte row per pixel of the (PH + 2) x PWM input patch = the 3 x BGR
  // bytes of that pixel and its two right neighbours (9 values, K padded to 16); the three filter
  // rows dy are three MMAs whose A operand starts dy * PWM rows further down.  RES: 128-byte rows of x.
  static constexpr int A1_SPAN = STEM ? 32 : 128;
  static constexpr int K1_STEPS = STEM ? 3 : 4;     // K = 16 MMAs per stage-1 M tile
  static constexpr int W1_BYTES = CM * A1_SPAN * (STEM ? 3 : 1);  // 3072 | 4096
  static constexpr int W2_TAP_BYTES = C2 * CM * 2;  // 4096
  static constexpr int W2_BYTES = 9 * W2_TAP_BYTES;
  static constexpr int BIAS_BYTES = (CM + C2) * 4;  // both bias vectors, read as broadcast LDS.128
  // Team-private regions.  Swizzles are functions of the absolute smem address, so operands only need
  // 128-byte alignment: STEM's intermediate patch overlays the (dead) stage-1 rows; RES's two x-patch
  // buffers are packed back to back and the output slab overlays the (dead) intermediate.
  static constexpr int IMG_ROWS = PH + 2, IMG_PITCH = 64;           // STEM image patch: 35 rows x 64 bytes
  static constexpr int A1_ROWS = STEM ? (MT1 * 128 + 2 * PWM + 4) : NP;  // incl. the last M tile's over-read
  static constexpr int A1_BYTES = STEM ? round_up_c(round_up_c(A1_ROWS * A1_SPAN, 1024), 1024) : NP * A1_SPAN;
  static constexpr int A1_BUFS = STEM ? 1 : 2;      // RES: x patch of the next tile is prefetched
  static constexpr int MID_BYTES = round_up_c(NP * CM * 2, 1024);   // 38912 | 12288
  static constexpr int STG_BYTES = 128 * C2 * 2;    // 16384
  static constexpr int IMG_BYTES = STEM ? round_up_c(IMG_ROWS * IMG_PITCH * 2, 1024) : 0;
  static constexpr int OFF_MID = STEM ? 0 : A1_BUFS * A1_BYTES;     // 0 | 46080
  static constexpr int OFF_STG = STEM ? (A1_BYTES > MID_BYTES ? A1_BYTES : MID_BYTES) : OFF_MID;  // 38912 | 46080
  static constexpr int OFF_IMG = OFF_STG + STG_BYTES;
  static constexpr int TEAM_BYTES = OFF_IMG + IMG_BYTES;            // 60416 | 62464
  static_assert(TEAM_BYTES % 1024 == 0 && OFF_STG % 1024 == 0 && OFF_MID % 1024 == 0, "region alignment");
  static_assert(STEM || NP * CM * 2 <= STG_BYTES, "intermediate patch must fit its overlay");
  // the stage-2 accumulator overlays the stage-1 accumulators (drained by epilogue 1 before stage 2 starts)
  static constexpr int TMEM_COLS_TEAM = MT1 * CM;   // 160 | 64
  static constexpr int TMEM_COLS = TEAMS * TMEM_COLS_TEAM <= 256 ? 256 : 512;
  static_assert(TEAMS * TMEM_COLS_TEAM <= 512 && C2 <= TMEM_COLS_TEAM, "TMEM budget");
  static constexpr int SMEM_BYTES = 1024 + W1_BYTES + W2_BYTES + TEAMS * TEAM_BYTES + BIAS_BYTES + 256;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  // stage-1 over-read of the last M tile (rows NP..MT1*128) must stay inside the team region
  static_assert(STEM || ((A1_BUFS - 1) * A1_BYTES + MT1 * 128 * A1_SPAN <= TEAM_BYTES), "RES stage-1 over-read");
  static constexpr int IMG_WORDS = IMG_ROWS * (IMG_PITCH / 4);  // 560 four-byte loads per image patch
  static constexpr int IMG_PRE = (IMG_WORDS + TEAM_THREADS - 1) / TEAM_THREADS;  // 5 per thread
};

struct ChainParams {
  int B, H2, W2;      // stage-2 output
  int Hm, Wm;         // intermediate (= stage-1 input) height / width
  int tiles_x, tiles_y, num_tiles;
  const uint8_t* img; // STEM: uint8 [B, Hm, Wm, 3] BGR
  const float* bias1;
  const float* bias2;
  int leaky1, leaky2;
};

__device__ __forceinline__ uint32_t swz128(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }
__device__ __forceinline__ uint32_t swz64(uint32_t a) { return a ^ (((a >> 7) & 3u) << 4); }
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
// packed fp32 pairs (FADD2 / FMUL2): same IEEE results as the scalar forms, half the issue slots
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(r))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return r;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(r))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return r;
}
// leaky(v) = max(v, slope * v) for slope <= 1 (slope 1 = identity)
__device__ __forceinline__ float2 leaky2(float2 v, float2 slope) {
  const float2 t = fmul2(v, slope);
  return make_float2(fmaxf(v.x, t.x), fmaxf(v.y, t.y));
}
__device__ __forceinline__ uint32_t swz32(uint32_t a) { return a ^ (((a >> 7) & 1u) << 4); }

// K-major smem descriptor: start address, SBO (bytes between 8-row groups), layout 2/4/6 = SW128/64/32
__device__ __forceinline__ uint64_t chain_desc(uint32_t addr, uint32_t sbo_bytes, uint64_t layout) {
  return (uint64_t(sbo_bytes >> 4) << 32) | (1ull << 46) | (layout << 61) | (1ull << 16) | uint64_t((addr >> 4) & 0x3FFFu);
}
__device__ __forceinline__ uint32_t chain_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(128 >> 4) << 24);
}

template <int MODE>
__global__ void __launch_bounds__(ChainCfg<MODE>::THREADS, 1)
conv_chain_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                  const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_y,
                  const ChainParams p) {
  using Cfg = ChainCfg<MODE>;
  constexpr bool STEM = Cfg::STEM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w1_s = smem_base;
  const uint32_t w2_s = w1_s + Cfg::W1_BYTES;
  const uint32_t teams_s = w2_s + Cfg::W2_BYTES;
  const uint32_t bias_s = teams_s + Cfg::TEAMS * Cfg::TEAM_BYTES;  // bias1[CM] then bias2[C2], fp32
  const uint32_t bar_base = bias_s + Cfg::BIAS_BYTES;
  const uint32_t wbar = bar_base;                            // weights landed
  const uint32_t tmem_slot = bar_base + 8;
  const int team = threadIdx.x >> 7;
  const int tid = threadIdx.x & 127;
  const int warp = tid >> 5;  // warp within the team = TMEM lane quarter
  const int lane = tid & 31;
  const uint32_t mma_bar = bar_base + 16 + 32u * team;       // per team: mma_bar, xbar[2]
  const uint32_t xbar0 = mma_bar + 8;
  const uint32_t team_s = teams_s + team * Cfg::TEAM_BYTES;
  const uint32_t mid_s = team_s + Cfg::OFF_MID;
  const uint32_t stg_s = team_s + Cfg::OFF_STG;
  uint32_t* const tmem_slot_gen = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap_w1);
    ptx::prefetch_tensormap(&tmap_w2);
    ptx::prefetch_tensormap(&tmap_y);
    if (!STEM) ptx::prefetch_tensormap(&tmap_x);
    ptx::mbar_init(wbar, 1);
    for (int t = 0; t < Cfg::TEAMS; ++t) {
      ptx::mbar_init(bar_base + 16 + 32u * t, 1);
      ptx::mbar_init(bar_base + 16 + 32u * t + 8, 1);
      ptx::mbar_init(bar_base + 16 + 32u * t + 16, 1);
    }
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    ptx::tmem_alloc<1>(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish<1>();
  }
  if (threadIdx.x < Cfg::CM + Cfg::C2) {  // biases are constants: no dependency on the previous kernel
    float* const bias_gen = reinterpret_cast<float*>(smem_gen + (bias_s - smem_base));
    bias_gen[threadIdx.x] = threadIdx.x < Cfg::CM ? __ldg(p.bias1 + threadIdx.x) : __ldg(p.bias2 + threadIdx.x - Cfg::CM);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_team = *tmem_slot_gen + team * Cfg::TMEM_COLS_TEAM;
  pdl_wait();

  if (threadIdx.x == 0) {  // resident weights of both layers
    ptx::mbar_arrive_expect_tx(wbar, Cfg::W1_BYTES + Cfg::W2_BYTES);
    if constexpr (STEM) {
      for (int dy = 0; dy < 3; ++dy) ptx::tma_load_2d(w1_s + dy * Cfg::CM * Cfg::A1_SPAN, &tmap_w1, wbar, dy * 16, 0);
    } else {
      ptx::tma_load_2d(w1_s, &tmap_w1, wbar, 0, 0);
    }
    for (int tap = 0; tap < 9; ++tap) ptx::tma_load_2d(w2_s + tap * Cfg::W2_TAP_BYTES, &tmap_w2, wbar, tap * Cfg::CM, 0);
  }

  const int team_gid = blockIdx.x * Cfg::TEAMS + team;
  const int team_step = gridDim.x * Cfg::TEAMS;
  const uint32_t bar_id = 1 + team;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const float2 slope1 = p.leaky1 ? make_float2(0.1f, 0.1f) : make_float2(1.f, 1.f);  // leaky(v) = max(v, 0.1 v)
  const float2 slope2 = p.leaky2 ? make_float2(0.1f, 0.1f) : make_float2(1.f, 1.f);
  float2 bias1[Cfg::CM / 2];  // every thread applies all 32 stage-1 biases to its patch pixel
#pragma unroll
  for (int q = 0; q < Cfg::CM / 2; ++q) bias1[q] = __ldg(reinterpret_cast<const float2*>(p.bias1) + q);
  auto tile_origin = [&](int tile, int& b, int& y0, int& x0) {
    b = tile / tiles_per_img;
    const int rem = tile - b * tiles_per_img;
    const int ty = rem / p.tiles_x;
    y0 = ty * Cfg::TH;
    x0 = (rem - ty * p.tiles_x) * Cfg::TW;
  };

  // ---- STEM: register prefetch of the next tile's uint8 patch: 35 rows x 64 bytes starting one BGR
  // pixel + 2 bytes left of the tile's first needed pixel, which makes every row 8-byte aligned.
  // Word e = i * 128 + tid of the patch is row e / 16, byte 4 (e % 16).
  uint32_t pre[STEM ? Cfg::IMG_PRE : 1];
  auto img_prefetch = [&](int tile) {
    if constexpr (STEM) {
      int b, y0, x0;
      tile_origin(tile, b, y0, x0);
      const int byte0 = 6 * x0 - 8;
      const int row_bytes = 3 * p.Wm;
      const uint8_t* const base = p.img + (long long)b * p.Hm * row_bytes;
#pragma unroll
      for (int i = 0; i < Cfg::IMG_PRE; ++i) {
        const int e = i * Cfg::TEAM_THREADS + tid;
        const int iy = 2 * y0 - 2 + (e >> 4);
        const int byte = byte0 + 4 * (e & 15);
        const bool ok = e < Cfg::IMG_WORDS && iy >= 0 && iy < p.Hm && byte >= 0 && byte < row_bytes;
        uint32_t v = 0;
        if (ok) v = __ldg(reinterpret_cast<const unsigned int*>(base + (long long)iy * row_bytes + byte));
        pre[i] = v;
      }
    }
  };

  if (team_gid < p.num_tiles) {
    if constexpr (STEM) {
      img_prefetch(team_gid);
    } else {
      if (tid == 0) {
        int b, y0, x0;
        tile_origin(team_gid, b, y0, x0);
        ptx::mbar_arrive_expect_tx(xbar0, Cfg::NP * Cfg::A1_SPAN);
        ptx::tma_load_4d(team_s, &tmap_x, xbar0, 0, x0 - 1, y0 - 1, b);
      }
    }
    if (tid == 0) ptx::mbar_wait(wbar, 0);
  }

  int it = 0;
  for (int tile = team_gid; tile < p.num_tiles; tile += team_step, ++it) {
    int b, y0, x0;
    tile_origin(tile, b, y0, x0);
    const int cur = STEM ? 0 : (it & 1);
    const uint32_t a1_s = team_s + cur * Cfg::A1_BYTES;

    if constexpr (STEM) {
      // (1) image bytes -> bf16 /255 patch [35][64], still in BGR byte order (the first layer's weights are
      // laid out to match).  v * fl(1/255) rounds to the same bf16 as the reference's fp32 v / 255 for
      // all 256 byte values (inference.py:332-333).
      const uint32_t patch_s = team_s + Cfg::OFF_IMG;
#pragma unroll
      for (int i = 0; i < Cfg::IMG_PRE; ++i) {
        const int e = i * Cfg::TEAM_THREADS + tid;
        if (e < Cfg::IMG_WORDS) {
          const float k = 0.003921568859368563f;
          const uint32_t lo = pack_bf16x2(__fmul_rn((float)(pre[i] & 0xffu), k), __fmul_rn((float)((pre[i] >> 8) & 0xffu), k));
          const uint32_t hi = pack_bf16x2(__fmul_rn((float)((pre[i] >> 16) & 0xffu), k), __fmul_rn((float)(pre[i] >> 24), k));
          asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(patch_s + 8 * e), "r"(lo), "r"(hi) : "memory");
        }
      }
      ptx::named_bar_sync(bar_id, Cfg::TEAM_THREADS);
      if (tile + team_step < p.num_tiles) img_prefetch(tile + team_step);
      // (2) stage-1 A rows: row q = r * 18 + c holds the 9 values of patch pixels c, c+1, c+2 of row r
      // (pixel c starts at value 2 + 3c), K padded to 16
      const unsigned short* const pu = reinterpret_cast<const unsigned short*>(smem_gen + (patch_s - smem_base));
      for (int q = tid; q < Cfg::IMG_ROWS * Cfg::PWM; q += Cfg::TEAM_THREADS) {
        const int r = q / Cfg::PWM;
        const int c = q - r * Cfg::PWM;
        const unsigned short* src = pu + r * Cfg::IMG_PITCH + 2 + 3 * c;
        uint32_t v[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] = src[j];
        const uint32_t row = a1_s + q * Cfg::A1_SPAN;
        sts128(swz32(row), v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
        sts128(swz32(row + 16), v[8], 0u, 0u, 0u);
      }
      ptx::fence_proxy_async();
    } else {
      if (tid == 0 && tile + team_step < p.num_tiles) {  // prefetch the next tile's x patch
        int nb, ny0, nx0;
        tile_origin(tile + team_step, nb, ny0, nx0);
        const uint32_t xb = xbar0 + 8u * (cur ^ 1);
        ptx::mbar_arrive_expect_tx(xb, Cfg::NP * Cfg::A1_SPAN);
        ptx::tma_load_4d(team_s + (cur ^ 1) * Cfg::A1_BYTES, &tmap_x, xb, 0, nx0 - 1, ny0 - 1, nb);
      }
      ptx::mbar_wait(xbar0 + 8u * cur, (it >> 1) & 1);
    }
    // the previous tile's TMA store must have finished reading the slab before anyone rewrites it
    // (RES: the slab is also the intermediate patch written by epilogue 1)
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    ptx::tc_fence_before();
    ptx::named_bar_sync(bar_id, Cfg::TEAM_THREADS);

    // ---- stage 1 MMAs ----
    if (tid == 0) {
      ptx::tc_fence_after();
      if constexpr (STEM) {
        // filter row dy: same A rows shifted by dy * PWM (next patch row), weights slice dy; SW32 rows
#pragma unroll
        for (int t = 0; t < Cfg::MT1; ++t) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t da = chain_desc(a1_s + (t * 128 + dy * Cfg::PWM) * Cfg::A1_SPAN, 8 * Cfg::A1_SPAN, 6);
            const uint64_t db = chain_desc(w1_s + dy * Cfg::CM * Cfg::A1_SPAN, 8 * Cfg::A1_SPAN, 6);
            ptx::umma_bf16_ss<1>(tmem_team + t * Cfg::CM, da, db, chain_idesc(Cfg::CM), dy != 0);
          }
        }
      } else {
        const uint64_t db = chain_desc(w1_s, 8 * Cfg::A1_SPAN, 2);
#pragma unroll
        for (int t = 0; t < Cfg::MT1; ++t) {
          const uint64_t da = chain_desc(a1_s + t * 128 * Cfg::A1_SPAN, 8 * Cfg::A1_SPAN, 2);
#pragma unroll
          for (int k = 0; k < Cfg::K1_STEPS; ++k)
            ptx::umma_bf16_ss<1>(tmem_team + t * Cfg::CM, da + 2u * k, db + 2u * k, chain_idesc(Cfg::CM), k != 0);
        }
      }
      ptx::umma_commit<1>(mma_bar);
    }
    __syncwarp();
    ptx::mbar_wait(mma_bar, 0);
    ptx::tc_fence_after();

    // ---- epilogue 1: intermediate patch (bf16, zero outside the image) ----
    const uint32_t tlane = uint32_t(warp * 32) << 16;
#pragma unroll 1
    for (int t = 0; t < Cfg::MT1; ++t) {
      uint32_t v[32];
      ptx::tmem_ld_x16(tmem_team + tlane + t * Cfg::CM, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      ptx::tmem_ld_x16(tmem_team + tlane + t * Cfg::CM + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
      ptx::tmem_ld_wait();
      const int pi = t * 128 + warp * 32 + lane;
      if (pi < Cfg::NP) {
        const int mr = pi / Cfg::PWM;
        const int mc = pi - mr * Cfg::PWM;
        const int my = y0 * Cfg::S2 - 1 + mr;
        const int mx = x0 * Cfg::S2 - 1 + mc;
        const bool inside = my >= 0 && my < p.Hm && mx >= 0 && mx < p.Wm;
        const uint32_t row = mid_s + pi * (Cfg::CM * 2);
        uint32_t w[16];
        if (inside) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float2 a = leaky2(fadd2(make_float2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])), bias1[q]), slope1);
            w[q] = pack_bf16x2(a.x, a.y);
          }
        } else {  // stage 2's zero padding applies to the intermediate, not to stage 1's input
#pragma unroll
          for (int q = 0; q < 16; ++q) w[q] = 0u;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t a = STEM ? swz128(row + c * 16) : swz64(row + c * 16);
          sts128(a, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        }
      }
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    ptx::named_bar_sync(bar_id, Cfg::TEAM_THREADS);

    // ---- stage 2 MMAs: nine taps = nine start offsets into the patch ----
    if (tid == 0) {
      ptx::tc_fence_after();
      constexpr uint64_t L2 = STEM ? 2 : 4;                       // pixel pairs as SW128 rows | SW64 rows
      constexpr uint32_t SBO2 = Cfg::S2 * Cfg::PWM * Cfg::CM * 2;  // next output row = S2 patch rows down
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int r = tap / 3, s = tap - 3 * (tap / 3);
        const uint64_t da = chain_desc(mid_s + (r * Cfg::PWM + s) * (Cfg::CM * 2), SBO2, L2);
        const uint64_t db = chain_desc(w2_s + tap * Cfg::W2_TAP_BYTES, 8 * Cfg::CM * 2, 4);
#pragma unroll
        for (int k = 0; k < Cfg::CM / 16; ++k)
          ptx::umma_bf16_ss<1>(tmem_team, da + 2u * k, db + 2u * k, chain_idesc(Cfg::C2), (tap | k) != 0);
      }
      ptx::umma_commit<1>(mma_bar);
    }
    __syncwarp();
    ptx::mbar_wait(mma_bar, 1);
    ptx::tc_fence_after();

    // ---- epilogue 2: +bias, leaky, (+x), bf16 -> swizzled slab -> TMA store ----
    {
      const int m = warp * 32 + lane;
      const uint32_t srow = stg_s + m * 128;
      const uint32_t rrow = a1_s + (((m >> 3) + 1) * Cfg::PWM + (m & 7) + 1) * 128;  // RES: x at the output pixel
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[16];
        ptx::tmem_ld_x16(tmem_team + tlane + cc * 16, v);
        ptx::tmem_ld_wait();
        float2 f[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 bq = lds128(bias_s + Cfg::CM * 4 + cc * 64 + q * 16);
          f[2 * q] = fadd2(make_float2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])),
                           make_float2(__uint_as_float(bq.x), __uint_as_float(bq.y)));
          f[2 * q + 1] = fadd2(make_float2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])),
                               make_float2(__uint_as_float(bq.z), __uint_as_float(bq.w)));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = leaky2(f[j], slope2);
        if constexpr (!STEM) {
          const uint4 r0 = lds128(swz128(rrow + (2 * cc) * 16));
          const uint4 r1 = lds128(swz128(rrow + (2 * cc + 1) * 16));
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fadd2(f[j], unpack_bf16x2(rr[j]));
        }
        sts128(swz128(srow + (2 * cc) * 16), pack_bf16x2(f[0].x, f[0].y), pack_bf16x2(f[1].x, f[1].y),
               pack_bf16x2(f[2].x, f[2].y), pack_bf16x2(f[3].x, f[3].y));
        sts128(swz128(srow + (2 * cc + 1) * 16), pack_bf16x2(f[4].x, f[4].y), pack_bf16x2(f[5].x, f[5].y),
               pack_bf16x2(f[6].x, f[6].y), pack_bf16x2(f[7].x, f[7].y));
      }
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async();
    ptx::named_bar_sync(bar_id, Cfg::TEAM_THREADS);
    if (tid == 0) {
      ptx::tma_store_4d(&tmap_y, stg_s, 0, x0, y0, b);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");

  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(*tmem_slot_gen, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
template <int MODE>
static int launch_chain(const y3_chain_desc* d, const void* x, const void* w1, const float* b1, const void* w2,
                        const float* b2, void* y, cudaStream_t stream) {
  using Cfg = ChainCfg<MODE>;
  constexpr bool STEM = Cfg::STEM;
  const int h2 = STEM ? (d->h - 1) / 2 + 1 : d->h;  // 3x3 / stride S2 / pad 1
  const int w2o = STEM ? (d->w - 1) / 2 + 1 : d->w;
  Y3_CHECK_ARG(h2 % Cfg::TH == 0 && w2o % Cfg::TW == 0,
               "conv chain: output %dx%d must tile by %dx%d", h2, w2o, Cfg::TH, Cfg::TW);
  Y3_CHECK_ARG(!STEM || (d->w % 8 == 0 && d->h % 2 == 0), "conv chain stem: image width must be a multiple of 8");

  ChainParams p;
  p.B = d->n; p.H2 = h2; p.W2 = w2o;
  p.Hm = d->h; p.Wm = d->w;
  p.tiles_x = w2o / Cfg::TW; p.tiles_y = h2 / Cfg::TH;
  const long long nt = (long long)d->n * p.tiles_x * p.tiles_y;
  Y3_CHECK_ARG(nt > 0 && nt < (1ll << 31), "conv chain: tile count out of range");
  p.num_tiles = (int)nt;
  p.img = STEM ? reinterpret_cast<const uint8_t*>(x) : nullptr;
  p.bias1 = b1; p.bias2 = b2;
  p.leaky1 = d->leaky1; p.leaky2 = d->leaky2;

  alignas(64) CUtensorMap tx, tw1, tw2, ty;
  memset(&tx, 0, sizeof(tx));
  int rc;
  {  // stage-1 weights: STEM [32][3 dy][16] read as three [16 x 32] boxes; RES [32][64]
    const uint64_t kk = STEM ? 48 : 64;
    const uint64_t dims[2] = {kk, (uint64_t)Cfg::CM};
    const uint64_t str[1] = {kk * 2};
    const uint32_t box[2] = {(uint32_t)(STEM ? 16 : 64), (uint32_t)Cfg::CM};
    if ((rc = encode_tiled_map(&tw1, 2, w1, dims, str, box, Cfg::A1_SPAN)) != Y3_OK) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)9 * Cfg::CM, (uint64_t)Cfg::C2};
    const uint64_t str[1] = {(uint64_t)9 * Cfg::CM * 2};
    const uint32_t box[2] = {(uint32_t)Cfg::CM, (uint32_t)Cfg::C2};
    if ((rc = encode_tiled_map(&tw2, 2, w2, dims, str, box, Cfg::CM * 2)) != Y3_OK) return rc;
  }
  {
    const uint64_t pix = (uint64_t)d->ld_y * 2;
    const uint64_t dims[4] = {(uint64_t)Cfg::C2, (uint64_t)w2o, (uint64_t)h2, (uint64_t)d->n};
    const uint64_t str[3] = {pix, pix * w2o, pix * w2o * h2};
    const uint32_t box[4] = {(uint32_t)Cfg::C2, (uint32_t)Cfg::TW, (uint32_t)Cfg::TH, 1};
    if ((rc = encode_tiled_map(&ty, 4, y, dims, str, box, 128)) != Y3_OK) return rc;
  }
  if (!STEM) {
    const uint64_t pix = (uint64_t)d->ld_x * 2;
    const uint64_t dims[4] = {64, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n};
    const uint64_t str[3] = {pix, pix * d->w, pix * d->w * d->h};
    const uint32_t box[4] = {64, (uint32_t)Cfg::PWM, (uint32_t)Cfg::PH, 1};
    if ((rc = encode_tiled_map(&tx, 4, x, dims, str, box, 128)) != Y3_OK) return rc;
  }

  auto kernel = conv_chain_kernel<MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    Y3_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int want = (p.num_tiles + Cfg::TEAMS - 1) / Cfg::TEAMS;
  const int grid = want < num_sms() ? want : num_sms();
  Y3_CUDA_OK(launch_kernel(kernel, dim3(grid), dim3(Cfg::THREADS), (size_t)Cfg::SMEM_BYTES, stream, tx, tw1, tw2, ty, p));
  Y3_LAUNCH_OK("conv_chain_kernel");
  return Y3_OK;
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

}  // namespace y3

using namespace y3;

extern "C" {

int y3_conv_chain_stem_u8(const y3_chain_desc* d, const uint8_t* img, const void* w1, const float* b1, const void* w2,
                          const float* b2, void* y, void* stream) {
  Y3_CHECK_ARG(d && img && w1 && b1 && w2 && b2 && y, "conv chain stem: null argument");
  Y3_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0, "conv chain stem: bad shape");
  Y3_CHECK_ARG(d->ld_y >= 64 && d->ld_y % 8 == 0, "conv chain stem: ld_y=%d invalid", d->ld_y);
  Y3_CHECK_ARG(aligned16(w1) && aligned16(w2) && aligned16(y) && aligned16(b1) && aligned16(b2) &&
               (reinterpret_cast<uintptr_t>(img) & 7) == 0, "conv chain stem: alignment");
  return launch_chain<CHAIN_STEM>(d, img, w1, b1, w2, b2, y, reinterpret_cast<cudaStream_t>(stream));
}

int y3_conv_chain_res64(const y3_chain_desc* d, const void* x, const void* w1, const float* b1, const void* w2,
                        const float* b2, void* y, void* stream) {
  Y3_CHECK_ARG(d && x && w1 && b1 && w2 && b2 && y, "conv chain res: null argument");
  Y3_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0, "conv chain res: bad shape");
  Y3_CHECK_ARG(d->ld_x >= 64 && d->ld_x % 8 == 0 && d->ld_y >= 64 && d->ld_y % 8 == 0, "conv chain res: bad pitch");
  Y3_CHECK_ARG(aligned16(x) && aligned16(w1) && aligned16(w2) && aligned16(y) && aligned16(b1) && aligned16(b2),
               "conv chain res: alignment");
  Y3_CHECK_ARG(x != y, "conv chain res: in-place operation is not supported (tiles read their neighbours' halo)");
  return launch_chain<CHAIN_RES>(d, x, w1, b1, w2, b2, y, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"

// y3_debug_set_trap_record (api.cu): this translation unit's copy of the watchdog record pointer
namespace y3 { cudaError_t conv_chain_set_trap_record(unsigned long long* host_mapped) { return ptx::set_trap_record_tu(host_mapped); } }
