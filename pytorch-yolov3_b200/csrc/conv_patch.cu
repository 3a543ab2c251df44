// conv_patch.cu — 3x3 / stride 1 / pad 1 Darknet [convolutional] block with the input window kept
// in shared memory: the nine filter taps read ONE patch of the input instead of nine im2col tiles.
//
// Same arithmetic as conv_umma.cu (yolov3/darknet.py:244-257 executed at :367-368, shortcut add of
// :376-379 in the epilogue) — only the A-operand path differs.  conv_umma.cu's 3x3 layers are bound
// by L2 -> SM operand traffic (profiles/: stopping the A re-loads alone is worth 16 %), and 8 of
// every 9 A bytes are re-reads of pixels the SM already holds.  Here:
//
//   * Output pixels are addressed in a VIRTUAL row-padded space: image row h owns Wp = W + 2
//     positions, the last two are dummies.  Position v = h * Wp + wp needs, for tap (r, s), the
//     padded-input pixel at flat index v + r * Wp + s — a pure shift, no row-boundary cases.
//   * One 4-D tiled TMA per 64-channel block fetches the padded rows the tile touches
//     (box = 64 ch x Wp x NR rows, starting at w = -1: the left / right / top / bottom zero padding is
//     the tensor map's out-of-bounds fill) into a dense, 128B-swizzled [row][Wp][64] patch.
//   * A tile is 128 consecutive virtual positions per CTA; the tap (r, s) A operand is the patch
//     viewed from START OFFSET (off + r * Wp + s) * 128 bytes (tcgen05 swizzles are functions of the
//     absolute smem address, so a row-shifted descriptor start is legal — same property
//     conv_chain.cu relies on).  In a CTA pair (cta_group::2) both CTAs must present their rows at
//     the SAME smem offset, so the peer lands its patch shifted by (off_leader - off_peer) rows.
//   * Weights stream through their own ring, one 64-channel tap slab per stage; patches and weights
//     have a producer thread each, so a patch is requested one unit ahead of the MMAs whatever the
//     state of the weight ring.
//   * Epilogue: one thread per virtual position (its TMEM lane); dummies (wp >= W, v >= H * Wp) are
//     dropped.  A row's pixels are not consecutive in the dense output any more, so TMA boxes do not
//     apply; every warp transposes 32 rows x 64 columns through a 4 KB swizzled smem tile instead, so
//     that global loads (shortcut operand, prefetched into registers while the MMAs run) and stores
//     move whole 128-byte lines: 8 lanes per pixel, the pixel's address by shuffle from its owner.
//
// Cost: (W + 2) / W more MMA rows plus the ragged last tile of every image; the dispatcher
// (conv_umma.cu) only sends layers here whose tiles are >= 93 % full (W >= 38 at these sizes).
#include "common.cuh"
#define Y3_FILE_ID 2
#include "ptx.cuh"

#include <cuda.h>
#include <string.h>

namespace y3 {

static constexpr int PT_BLOCK_M = 128;
static constexpr int PT_BLOCK_K = 64;
// warps: patch producer, MMA issuer, 4 * EW epilogue warps (EW per TMEM lane quarter), weight producer
static constexpr int PT_MAX_B_STAGES = 12;

struct PatchParams {
  int H, W, Wp, V;                // V = H * Wp virtual positions per image
  int tiles_img;                  // tiles (CTA pairs' worth) per image
  int num_tiles, num_n_tiles;
  int cin, cin_blocks;
  uint32_t patch_bytes;           // bytes one patch TMA delivers (= box)
  uint32_t patch_stride;          // bytes between the two patch buffers (box + shift slack, 1024-multiple)
  int q128;                       // 128 % Wp: the peer's rows start q128 positions further into their padded row
  int b_stages;
  unsigned long long div_ntiles, div_tiles_img, div_wp;  // ceil(2^48 / d)
  const float* bias;
  __nv_bfloat16* out;
  const __nv_bfloat16* res;
  int ld_out, ld_res;
  int leaky;
};

__device__ __forceinline__ int pt_div(int x, unsigned long long m) {
  return (int)__umul64hi((unsigned long long)(unsigned)x << 16, m);
}
static unsigned long long pt_magic(int d) { return ((1ull << 48) + (unsigned long long)d - 1) / (unsigned long long)d; }

template <int BLOCK_N, int CG>
struct PatchCfg {
  static constexpr int B_ROWS = BLOCK_N / CG;
  static constexpr int B_BYTES = B_ROWS * PT_BLOCK_K * 2;
  static constexpr int BAR_BYTES = 512;   // (2 * 12 + 2 + 2 + 2 + 2) barriers * 8 + tmem pointer
  static constexpr int BIAS_BYTES = 2 * BLOCK_N * 4;
  // two epilogue warps per TMEM lane quarter for 256-channel tiles; one for 128-channel tiles, whose
  // 8 KB weight stages need every byte the transpose tiles can spare to keep the ring deep enough
  static constexpr int EW = BLOCK_N >= 256 ? 2 : 1;
  static constexpr int EPI_THREADS = 128 * EW;
  static constexpr int THREADS = 64 + EPI_THREADS + 32;
  static constexpr int WEIGHT_WARP = 2 + 4 * EW;
  static constexpr int STG_BYTES = 4 * EW * 4096;  // per epilogue warp: 32 rows x 64 columns bf16, swizzled
  static constexpr int TMEM_COLS = 2 * BLOCK_N <= 256 ? 256 : 512;
  static constexpr int COLS = BLOCK_N / (BLOCK_N >= 256 ? 2 : 1);  // columns per epilogue warp
  // K-major, 128B swizzle: SBO = 8 rows * 128 B, layout type 2
  static constexpr uint64_t DESC_HI = ((uint64_t(1024) >> 4) << 32) | (1ull << 46) | (2ull << 61);
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BLOCK_N >> 3) << 17) |
                                    (uint32_t((PT_BLOCK_M * CG) >> 4) << 24);
};

__device__ __forceinline__ uint64_t pt_desc(uint32_t smem_addr, uint64_t hi) {
  return hi | (1ull << 16) | uint64_t((smem_addr >> 4) & 0x3FFFu);
}

template <int BLOCK_N, int CG>
__global__ void __launch_bounds__(PatchCfg<BLOCK_N, CG>::THREADS, 1)
conv_patch_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_b,
                  const PatchParams p) {
  using Cfg = PatchCfg<BLOCK_N, CG>;
  const uint32_t cta_rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const int tile_first = blockIdx.x / CG;
  const int tile_step = gridDim.x / CG;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int B_STAGES = p.b_stages;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t patch_base = smem_base;                          // two buffers of patch_stride bytes
  const uint32_t b_base = smem_base + 2u * p.patch_stride;        // B ring
  const uint32_t stg_base = b_base + (uint32_t)B_STAGES * Cfg::B_BYTES;  // epilogue transpose tiles
  const uint32_t bar_base = stg_base + Cfg::STG_BYTES;
  // barriers: b_full[12] b_empty[12] p_full[2] p_empty[2] tfull[2] tempty[2] | tmem pointer
  auto bfull_bar = [&](int s) { return bar_base + 8u * s; };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (PT_MAX_B_STAGES + s); };
  auto pfull_bar = [&](int b) { return bar_base + 8u * (2 * PT_MAX_B_STAGES + b); };
  auto pempty_bar = [&](int b) { return bar_base + 8u * (2 * PT_MAX_B_STAGES + 2 + b); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * PT_MAX_B_STAGES + 4 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * PT_MAX_B_STAGES + 6 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * PT_MAX_B_STAGES + 8);
  uint32_t* tmem_ptr_gen = reinterpret_cast<uint32_t*>(smem_raw + (tmem_ptr_addr - smem_base));
  const uint32_t bias_base = bar_base + Cfg::BAR_BYTES;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_b);
    for (int s = 0; s < B_STAGES; ++s) {
      ptx::mbar_init(bfull_bar(s), CG);
      ptx::mbar_init(bempty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(pfull_bar(b), CG);
      ptx::mbar_init(pempty_bar(b), 1);
      ptx::mbar_init(tfull_bar(b), 1);
      ptx::mbar_init(tempty_bar(b), 4 * Cfg::EW * CG);  // one arrival per epilogue warp (of both CTAs)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc<CG>(tmem_ptr_addr, Cfg::TMEM_COLS);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync();
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // tile -> (image, first virtual position of the LEADER's 128 rows, n tile)
  auto tile_coords = [&](int tile, int& img, int& v0_leader, int& n_tile) {
    const int m_tile = pt_div(tile, p.div_ntiles);
    n_tile = tile - m_tile * p.num_n_tiles;
    img = pt_div(m_tile, p.div_tiles_img);
    v0_leader = (m_tile - img * p.tiles_img) * (PT_BLOCK_M * CG);
  };

  // Where a CTA lands its patch and where the tile's first row sits in it.  Both CTAs of a pair must
  // present row i of their 128 at the same smem offset; the peer's first position lies (off_l + 128)
  // % Wp into its padded row, so whichever of the two starts later inside its row lands its box
  // further down by the difference (one-sided slack of max(q, Wp - q) rows, q = 128 % Wp).
  auto placement = [&](int v0_leader, int& hp_own, int& dst_rows_own, int& row0) {
    const int hp_l = pt_div(v0_leader, p.div_wp);
    const int off_l = v0_leader - hp_l * p.Wp;
    hp_own = hp_l;
    dst_rows_own = 0;
    row0 = off_l;
    if (CG == 2) {
      int off_p = off_l + p.q128;
      int hp_p = hp_l + (PT_BLOCK_M - p.q128) / p.Wp;   // exact: 128 = k * Wp + q128
      if (off_p >= p.Wp) { off_p -= p.Wp; ++hp_p; }
      const int delta = off_l - off_p;                  // > 0: the peer's box goes delta rows down
      const int d_l = delta < 0 ? -delta : 0, d_p = delta > 0 ? delta : 0;
      row0 = d_l + off_l;
      if (cta_rank != 0) { hp_own = hp_p; dst_rows_own = d_p; }
      else dst_rows_own = d_l;
    }
  };

  if (warp == 0) {
    // ===================== TMA producer: input patches (one thread) =====================
    // Patches and weights have a producer thread each: a patch is requested as soon as its buffer is
    // free (one unit ahead of the MMAs), whatever the state of the weight ring.
    if (lane == 0) {
      const uint32_t pfull0 = CG == 2 ? ptx::mapa(pfull_bar(0), 0) : pfull_bar(0);  // completions count on the leader
      int pb = 0;
      uint32_t p_phase = 0;
      pdl_wait();
      for (int tile = tile_first; tile < p.num_tiles; tile += tile_step) {
        int img, v0l, n_tile;
        tile_coords(tile, img, v0l, n_tile);
        int hp, dst_rows, row0;
        placement(v0l, hp, dst_rows, row0);
        const uint32_t dst_off = (uint32_t)(dst_rows * 128);
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          ptx::mbar_wait(pempty_bar(pb), p_phase ^ 1u);
          if (CG == 1 || cta_rank == 0) ptx::mbar_arrive_expect_tx(pfull_bar(pb), CG * p.patch_bytes);
          else ptx::mbar_arrive_cluster(pfull0 + 8u * pb);
          // padded row hp is input row hp - 1; padded column 0 is input column -1
          ptx::tma_load_4d<CG>(patch_base + pb * p.patch_stride + dst_off, &tmap_x, pfull0 + 8u * pb, cb * PT_BLOCK_K, -1,
                               hp - 1, img);
          if (++pb == 2) { pb = 0; p_phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == Cfg::WEIGHT_WARP) {
    // ===================== TMA producer: weights (one thread) =====================
    // nine 64-channel tap slabs per channel block; constants, so no wait for the previous kernel
    if (lane == 0) {
      const uint32_t bfull0 = CG == 2 ? ptx::mapa(bfull_bar(0), 0) : bfull_bar(0);
      const int b_row0 = (int)cta_rank * Cfg::B_ROWS;
      uint32_t b_dst = b_base, bf = bfull0, bf_l = bfull_bar(0), be = bempty_bar(0);
      const uint32_t bf_end = bfull_bar(B_STAGES);
      uint32_t b_phase = 0;
      for (int tile = tile_first; tile < p.num_tiles; tile += tile_step) {
        const int n_tile = tile - pt_div(tile, p.div_ntiles) * p.num_n_tiles;
        const int b_row = n_tile * BLOCK_N + b_row0;
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          int k = cb * PT_BLOCK_K;
          for (int tap = 0; tap < 9; ++tap) {
            ptx::mbar_wait(be, b_phase ^ 1u);
            if (CG == 1 || cta_rank == 0) ptx::mbar_arrive_expect_tx(bf_l, CG * Cfg::B_BYTES);
            else ptx::mbar_arrive_cluster(bf);
            ptx::tma_load_2d<CG>(b_dst, &tmap_b, bf, k, b_row);
            k += p.cin;
            b_dst += Cfg::B_BYTES; bf += 8; bf_l += 8; be += 8;
            if (bf_l == bf_end) { b_dst = b_base; bf = bfull0; bf_l = bfull_bar(0); be = bempty_bar(0); b_phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t B_STEP = Cfg::B_BYTES >> 4;
      const uint64_t desc_b0 = pt_desc(b_base, Cfg::DESC_HI);
      uint64_t desc_b = desc_b0;
      uint32_t bf = bfull_bar(0), be = bempty_bar(0);
      const uint32_t bf_end = bfull_bar(B_STAGES);
      uint32_t b_phase = 0;
      int pb = 0;
      uint32_t p_phase = 0;
      int it = 0;
      for (int tile = tile_first; tile < p.num_tiles; tile += tile_step, ++it) {
        int img, v0l, n_tile;
        tile_coords(tile, img, v0l, n_tile);
        int hp_unused, dst_unused, row0;
        placement(v0l, hp_unused, dst_unused, row0);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        uint32_t accumulate = 0;
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          ptx::mbar_wait(pfull_bar(pb), p_phase);
          ptx::tc_fence_after();
          // row 0 of the tile inside the patch buffer
          uint64_t desc_row = pt_desc(patch_base + pb * p.patch_stride + (uint32_t)(row0 * 128), Cfg::DESC_HI);
          for (int r = 0; r < 3; ++r) {
            uint64_t desc_a = desc_row;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              ptx::mbar_wait(bf, b_phase);
              ptx::tc_fence_after();
#pragma unroll
              for (int k = 0; k < PT_BLOCK_K / 16; ++k) {
                ptx::umma_bf16_ss<CG>(tmem_d, desc_a + 2u * k, desc_b + 2u * k, Cfg::IDESC, accumulate);
                accumulate = 1;
              }
              ptx::umma_commit<CG>(be);
              desc_a += 8;  // next filter column: one pixel = 128 bytes further
              bf += 8; be += 8; desc_b += B_STEP;
              if (bf == bf_end) { bf = bfull_bar(0); be = bempty_bar(0); desc_b = desc_b0; b_phase ^= 1u; }
            }
            desc_row += (uint64_t)(p.Wp * 8);  // next filter row: Wp pixels further
          }
          ptx::umma_commit<CG>(pempty_bar(pb));  // patch reusable (in both CTAs) once these MMAs retire
          if (++pb == 2) { pb = 0; p_phase ^= 1u; }
        }
        ptx::umma_commit<CG>(tfull_bar(acc));
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2 .. 2 + 4 * EW - 1) =====================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int e_tid = (warp - 2) * 32 + lane;  // 0 .. EPI_THREADS - 1
    const int c_first = half * Cfg::COLS;
    constexpr int NBLK = Cfg::COLS / 64;       // 64-column blocks of this warp
    const uint32_t tempty0 = (CG == 2 && cta_rank != 0) ? ptx::mapa(tempty_bar(0), 0) : tempty_bar(0);
    const bool has_res = p.res != nullptr;
    // this warp's transpose tile: 32 rows x 128 bytes, 16-byte units XOR-swizzled by (row & 7)
    const uint32_t stg = stg_base + (uint32_t)(warp - 2) * 4096u;
    const uint32_t own_row = stg + (uint32_t)lane * 128u;                 // this thread's row (it owns TMEM lane `row`)
    const uint32_t own_swz = (uint32_t)(lane & 7);
    const int t_sub = lane >> 3, t_unit = lane & 7;                        // transposed view: 4 rows x 8 units per instruction
    bool waited = false;
    int it = 0;
    for (int tile = tile_first; tile < p.num_tiles; tile += tile_step, ++it) {
      int img, v0l, n_tile;
      tile_coords(tile, img, v0l, n_tile);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n0 = n_tile * BLOCK_N;
      const int v = v0l + (int)cta_rank * PT_BLOCK_M + row;
      const int h = pt_div(v, p.div_wp);
      const int wp = v - h * p.Wp;
      // dense output pixel of this thread's row, -1 for a dummy position
      const int m_own = (v < p.V && wp < p.W) ? (img * p.H + h) * p.W + wp : -1;
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BLOCK_N + c_first;
      const uint32_t bias_s = bias_base + (it & 1) * (BLOCK_N * 4);

      float bias_v = 0.f;
      if (e_tid < BLOCK_N) bias_v = __ldg(p.bias + n0 + e_tid);
      if (!waited) { pdl_wait(); waited = true; }
      // pixel index of the 8 rows this lane serves in the transposed view
      int m_t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) m_t[i] = __shfl_sync(0xffffffffu, m_own, 4 * i + t_sub);
      // the shortcut operand, whole 128-byte lines, in flight while the MMAs run
      uint4 rres[NBLK][8];
      if (has_res) {
#pragma unroll
        for (int b = 0; b < NBLK; ++b)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            rres[b][i] = m_t[i] >= 0 ? ld_nc_16(p.res + (long long)m_t[i] * p.ld_res + n0 + c_first + 64 * b + 8 * t_unit)
                                     : make_uint4(0u, 0u, 0u, 0u);
      }
      if (e_tid < BLOCK_N) asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4u * e_tid), "f"(bias_v) : "memory");
      ptx::named_bar_sync(1, Cfg::EPI_THREADS);
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
#pragma unroll
      for (int b = 0; b < NBLK; ++b) {
        if (has_res) {  // transposed registers -> tile
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t r_ = 4 * i + t_sub;
            const uint32_t ad = stg + r_ * 128u + (((uint32_t)t_unit ^ (r_ & 7u)) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ad), "r"(rres[b][i].x), "r"(rres[b][i].y),
                         "r"(rres[b][i].z), "r"(rres[b][i].w) : "memory");
          }
          __syncwarp();
        }
#pragma unroll
        for (int cc = 0; cc < 64; cc += 16) {
          uint32_t a[16];
          ptx::tmem_ld_x16(taddr + 64 * b + cc, a);
          float bz[16];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bz[4 * q]), "=f"(bz[4 * q + 1]), "=f"(bz[4 * q + 2]),
                         "=f"(bz[4 * q + 3]) : "r"(bias_s + 4u * (c_first + 64 * b + cc + 4 * q)));
          ptx::tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(a[j]) + bz[j];
          if (p.leaky) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.1f * f[j]);
          }
          const uint32_t u0 = (uint32_t)(cc >> 3);
          const uint32_t a0 = own_row + ((u0 ^ own_swz) << 4), a1 = own_row + (((u0 + 1) ^ own_swz) << 4);
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
        if (b == NBLK - 1) ptx::tc_fence_before();  // all TMEM reads of this warp are done
        __syncwarp();
        if (b == NBLK - 1 && lane == 0) ptx::mbar_arrive_cluster(tempty0 + 8u * acc);  // hand the accumulator back
        // tile -> global, 8 lanes per pixel: whole 128-byte lines
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t r_ = 4 * i + t_sub;
          const uint32_t ad = stg + r_ * 128u + (((uint32_t)t_unit ^ (r_ & 7u)) << 4);
          uint4 o;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"(ad));
          if (m_t[i] >= 0)
            st_16(p.out + (long long)m_t[i] * p.ld_out + n0 + c_first + 64 * b + 8 * t_unit, o);
        }
        __syncwarp();  // the tile is overwritten by the next block / tile
      }
    }
  }

  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync();
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*PtEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct PatchPlan {
  int Wp, V, tiles_img, NR, b_stages;
  uint32_t patch_bytes, patch_stride;
  size_t smem_bytes;
  double fill;  // useful rows / MMA rows
};

template <int BLOCK_N, int CG>
static bool patch_plan(const y3_conv_desc* d, PatchPlan* pl) {
  using Cfg = PatchCfg<BLOCK_N, CG>;
  pl->Wp = d->w + 2;
  pl->V = d->h * pl->Wp;
  const int rows_tile = PT_BLOCK_M * CG;
  pl->tiles_img = (pl->V + rows_tile - 1) / rows_tile;
  pl->NR = 3 + (129 + pl->Wp - 1) / pl->Wp;  // NR * Wp >= 3 * Wp + 129 rows are read
  if (pl->Wp > 256 || pl->NR > 256) return false;
  pl->patch_bytes = (uint32_t)pl->NR * pl->Wp * 128u;
  const int q = PT_BLOCK_M % pl->Wp;
  const int slack = CG == 2 ? (q > pl->Wp - q ? q : pl->Wp - q) : 0;  // rows one CTA of a pair may have to shift its box down
  pl->patch_stride = ((uint32_t)(pl->NR * pl->Wp + slack) * 128u + 1023u) & ~1023u;
  const long long left = 232448ll - 2ll * pl->patch_stride - Cfg::STG_BYTES - Cfg::BAR_BYTES - Cfg::BIAS_BYTES;
  if (left < 4ll * Cfg::B_BYTES) return false;
  long long st = left / Cfg::B_BYTES;
  pl->b_stages = (int)(st > PT_MAX_B_STAGES ? PT_MAX_B_STAGES : st);
  pl->smem_bytes = 2ull * pl->patch_stride + (size_t)pl->b_stages * Cfg::B_BYTES + Cfg::STG_BYTES + Cfg::BAR_BYTES + Cfg::BIAS_BYTES;
  pl->fill = (double)d->h * d->w / ((double)pl->tiles_img * rows_tile);
  return true;
}

template <int BLOCK_N, int CG>
static int launch_patch(const y3_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual,
                        void* y, cudaStream_t stream, const PatchPlan& pl) {
  using Cfg = PatchCfg<BLOCK_N, CG>;
  static PtEncodeTiledFn encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    Y3_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return Y3_ECUDA;
    }
    encode = reinterpret_cast<PtEncodeTiledFn>(fn);
  }
  PatchParams p;
  p.H = d->h; p.W = d->w; p.Wp = pl.Wp; p.V = pl.V;
  p.tiles_img = pl.tiles_img;
  p.num_n_tiles = d->cout / BLOCK_N;
  p.num_tiles = d->n * pl.tiles_img * p.num_n_tiles;
  p.cin = d->cin; p.cin_blocks = d->cin / PT_BLOCK_K;
  p.patch_bytes = pl.patch_bytes; p.patch_stride = pl.patch_stride;
  p.q128 = PT_BLOCK_M % pl.Wp;
  p.b_stages = pl.b_stages;
  p.div_ntiles = pt_magic(p.num_n_tiles);
  p.div_tiles_img = pt_magic(pl.tiles_img);
  p.div_wp = pt_magic(pl.Wp);
  p.bias = bias;
  p.out = reinterpret_cast<__nv_bfloat16*>(y);
  p.res = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.ld_out = d->ld_y; p.ld_res = d->ld_res;
  p.leaky = d->leaky;

  alignas(64) CUtensorMap tmap_x, tmap_b;
  {
    const uint64_t pix = (uint64_t)d->ld_x * 2;
    cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {pix, pix * d->w, pix * d->w * d->h};
    cuuint32_t box[4] = {PT_BLOCK_K, (cuuint32_t)pl.Wp, (cuuint32_t)pl.NR, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_patch: cuTensorMapEncodeTiled(x) failed (CUresult %d): box=[64,%d,%d,1]", (int)r, pl.Wp, pl.NR);
      return Y3_ECUDA;
    }
  }
  {
    const uint64_t k_total = 9ull * d->cin;
    cuuint64_t dims[2] = {k_total, (cuuint64_t)d->cout};
    cuuint64_t strides[1] = {k_total * 2};
    cuuint32_t box[2] = {PT_BLOCK_K, (cuuint32_t)Cfg::B_ROWS};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&tmap_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_patch: cuTensorMapEncodeTiled(w) failed (CUresult %d)", (int)r);
      return Y3_ECUDA;
    }
  }
  auto kernel = conv_patch_kernel<BLOCK_N, CG>;
  static bool attr_set = false;
  if (!attr_set) {
    Y3_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const int slots = num_sms() / CG;
  const int grid = CG * (p.num_tiles < slots ? p.num_tiles : slots);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = pl.smem_bytes;
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
  Y3_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, tmap_x, tmap_b, p));
  Y3_LAUNCH_OK("conv_patch_kernel");
  return Y3_OK;
}

// Called by conv2d_impl (conv_umma.cu) for 3x3 / stride 1 / pad 1 bf16 layers.  Returns -1 when the
// layer should stay on the im2col kernel (shape not covered, tiles too ragged, or Y3_NO_PATCH=1).
int conv3x3_patch_try(const y3_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual,
                      void* y, cudaStream_t stream) {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("Y3_NO_PATCH"); disabled = (e && e[0] == '1') ? 1 : 0; }
  if (disabled || d->flags) return -1;
  if (d->ksize != 3 || d->stride != 1 || d->pad != 1 || d->out_f32 || d->upsample2x) return -1;
  if (d->cin % PT_BLOCK_K || d->ld_x % 8 || d->ld_y % 8 || (residual && d->ld_res % 8)) return -1;
  if ((long long)d->n * d->h * d->w >= (1ll << 31)) return -1;
  static double min_fill = -1.0;
  if (min_fill < 0) { const char* e = getenv("Y3_PATCH_MIN_FILL"); min_fill = e ? atof(e) : 0.93; }
  PatchPlan pl;
  if (d->cout % 256 == 0) {
    if (!patch_plan<256, 2>(d, &pl) || pl.fill < min_fill) return -1;
    return launch_patch<256, 2>(d, x, w, bias, residual, y, stream, pl);
  }
  // 128-channel tiles: measured slower than the im2col kernel (8 KB weight stages are too short for the
  // ring the patches leave room for) — kept for experiments behind Y3_PATCH_N128=1
  static int n128 = -1;
  if (n128 < 0) { const char* e = getenv("Y3_PATCH_N128"); n128 = (e && e[0] == '1') ? 1 : 0; }
  if (n128 && d->cout % 128 == 0) {
    if (!patch_plan<128, 2>(d, &pl) || pl.fill < min_fill) return -1;
    return launch_patch<128, 2>(d, x, w, bias, residual, y, stream, pl);
  }
  return -1;
}

}  // namespace y3

// y3_debug_set_trap_record (api.cu): this translation unit's copy of the watchdog record pointer
namespace y3 { cudaError_t conv_patch_set_trap_record(unsigned long long* host_mapped) { return ptx::set_trap_record_tu(host_mapped); } }
