// nms.cu — greedy IoU suppression, bit-exact with the reference's NumPy implementation
// (_non_max_suppression, yolov3/inference.py:161-217; per-class driver, :220-266).
//
// Three kernels, all integer / fp64-compare work on a few hundred KB per image:
//   1. nms_bucket_kernel   one CTA per image: class histogram in shared memory, exclusive
//                          scan, scatter of the candidates into per-class segments.
//   2. nms_bitmask_kernel  (image, class) segments of <= 128 / 256 / 512 boxes: rank sort in shared
//                          memory, the full "i suppresses j" bit matrix (fp32 test with an exact
//                          fallback inside a guard band), one warp scans it in score order.
//      nms_segment_kernel  larger segments: bitonic sort in shared memory, then 32 pivots at a time
//                          (32x32 mask by shuffles, resolved with ballots, every thread tests its
//                          later boxes against the chunk's KEPT pivots only), boxes in shared memory.
//   3. nms_compact_kernel  ordered compaction of the kept records (block scan).
// IoU follows the reference exactly: "+1" pixel areas in int64, iou = inter/union as an IEEE
// float64 divide, suppressed iff iou > thresh.
#include "common.cuh"

namespace y3 {

static constexpr int MAX_CLASSES = 1024;

struct Box4 { int x1, y1, x2, y2; };

__device__ __forceinline__ bool iou_gt(const Box4& a, const Box4& b, double thr) {
  // disjoint boxes (iw <= 0 or ih <= 0): the intersection is exactly 0, no 64-bit work needed
  if (min(a.x2, b.x2) < max(a.x1, b.x1) || min(a.y2, b.y2) < max(a.y1, b.y1)) return 0.0 > thr;
  const long long area_a = ((long long)a.x2 - a.x1 + 1) * ((long long)a.y2 - a.y1 + 1);
  const long long area_b = ((long long)b.x2 - b.x1 + 1) * ((long long)b.y2 - b.y1 + 1);
  long long iw = (long long)min(a.x2, b.x2) - (long long)max(a.x1, b.x1) + 1;
  long long ih = (long long)min(a.y2, b.y2) - (long long)max(a.y1, b.y1) + 1;
  iw = iw > 0 ? iw : 0;
  ih = ih > 0 ? ih : 0;
  const long long inter = iw * ih;
  if (inter == 0) return 0.0 > thr;  // disjoint boxes: iou is exactly 0
  const long long uni = area_a + area_b - inter;
  // iou > thr  <=>  fl(inter/uni) > thr.  Away from the boundary the comparison is decided by one
  // fp64 multiply (inter, uni < 2^53 are exact; the product carries <= 2^-53 relative error); only
  // within 2^-49 of it the exact IEEE divide the reference performs is evaluated.
  const double di = (double)inter, du = (double)uni;
  const double t = thr * du;
  const double slack = fabs(t) * 1.8e-15;
  if (di > t + slack) return true;
  if (di < t - slack) return false;
  return __ddiv_rn(di, du) > thr;
}

// workspace layout: bucketed records [N][cap] | seg_off int32 [N][C+1] | four work lists int32 [4][N*C] |
// four list counters.  class-agnostic mode uses C = 1.
//
// Work lists: the bucket kernel files every non-empty (image, class) segment under its size class
// (<= 128, <= 256, <= 512 boxes, larger) and the suppression kernels run PERSISTENT grids over "their" list.
// (Launching one CTA per (image, class, size class) instead — 3 x 5120 CTAs for 64 images x 80 classes,
// most of which exit at once — cost 0.1 ms per launch in block scheduling alone.)
static constexpr int NUM_LISTS = 4;
__device__ __forceinline__ int size_class(int n) { return n <= 128 ? 0 : n <= 256 ? 1 : n <= 512 ? 2 : 3; }

__global__ void __launch_bounds__(1024)
nms_bucket_kernel(const y3_cand* __restrict__ cands, const int* __restrict__ counts, int cap,
                  int num_classes, int per_class, y3_cand* __restrict__ bucketed,
                  int* __restrict__ seg_off, int* __restrict__ class_first_box,
                  int* __restrict__ class_start, int* __restrict__ class_kept, int* __restrict__ lists,
                  int* __restrict__ list_counts, int list_stride) {
  pdl_enter();
  __shared__ int hist[MAX_CLASSES];
  __shared__ int first[MAX_CLASSES];
  __shared__ int offs[MAX_CLASSES + 1];
  const int img = blockIdx.x;
  const int C = per_class ? num_classes : 1;
  int n = counts[img];
  n = n < cap ? n : cap;
  const y3_cand* src = cands + (long long)img * cap;
  y3_cand* dst = bucketed + (long long)img * cap;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { hist[c] = 0; first[c] = 0x7fffffff; }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int c = per_class ? src[i].cls : 0;
    c = min(max(c, 0), C - 1);
    atomicAdd(&hist[c], 1);
    atomicMin(&first[c], src[i].box);
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // C <= 1024: a serial scan is a few hundred cycles
    int run = 0;
    for (int c = 0; c < C; ++c) { offs[c] = run; run += hist[c]; }
    offs[C] = run;
  }
  __syncthreads();
  for (int c = threadIdx.x; c <= C; c += blockDim.x) {
    seg_off[(long long)img * (C + 1) + c] = offs[c];
    if (class_start) class_start[(long long)img * (C + 1) + c] = offs[c];
  }
  if (class_first_box)
    for (int c = threadIdx.x; c < C; c += blockDim.x) class_first_box[(long long)img * C + c] = first[c];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {  // file the segment under its size class
    const int len = hist[c];
    if (len == 0) {
      if (class_kept) class_kept[(long long)img * C + c] = 0;
    } else {
      const int k = size_class(len);
      lists[k * list_stride + atomicAdd(list_counts + k, 1)] = img * C + c;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) hist[c] = 0;  // reuse as cursors
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint4 lo = reinterpret_cast<const uint4*>(src + i)[0];
    const uint4 hi = reinterpret_cast<const uint4*>(src + i)[1];
    int c = per_class ? (int)hi.y : 0;
    c = min(max(c, 0), C - 1);
    const int pos = offs[c] + atomicAdd(&hist[c], 1);
    reinterpret_cast<uint4*>(dst + pos)[0] = lo;
    reinterpret_cast<uint4*>(dst + pos)[1] = hi;
  }
}

// key order: higher prob first; ties by lower box index (the reference's tie order is
// unspecified — np.argsort is not stable — so any deterministic rule is admissible).
__device__ __forceinline__ bool before(float pa, int ba, float pb, int bb) {
  return pa > pb || (pa == pb && ba < bb);
}

// ---- large segments (> 512 boxes: class-agnostic NMS, or a class that dominates an image) --------------
// One CTA of 1024 threads per segment, shared memory instead of global memory for everything hot:
//   1. sort: the 64-bit keys (prob descending, box ascending) are bitonic-sorted in shared memory
//      (up to SEG_SORT_MAX boxes, 128 KB); every source record then finds its rank by binary search of its
//      own key (keys are unique: box indices are) and is copied to its sorted position.  Larger segments
//      fall back to the O(n^2) rank count straight from global memory.
//   2. greedy suppression, 32 pivots at a time: warp 0 resolves the 32 x 32 block among the pivots with
//      shuffles / ballots, then every thread tests its later, still alive boxes against the chunk's KEPT
//      pivots only.  Boxes (and alive flags) live in shared memory when the segment fits (SEG_SMEM_BOXES).
static constexpr int SEG_SORT_MAX = 16384;    // keys: 8 B each
static constexpr int SEG_SMEM_BOXES = 6144;   // fp32 box 16 B + area 4 B + alive 1 B each: 126 KB
static constexpr int SEG_SMEM_BYTES = SEG_SORT_MAX * 8;

__device__ __forceinline__ float fast_area(int x1, int y1, int x2, int y2);

// monotone map: larger prob -> smaller key; equal prob -> smaller box first
__device__ __forceinline__ unsigned long long seg_key(uint32_t prob_bits, uint32_t box) {
  if (prob_bits == 0x80000000u) prob_bits = 0u;  // -0 == +0, as in before()
  const uint32_t asc = prob_bits ^ ((prob_bits >> 31) ? 0xffffffffu : 0x80000000u);  // float order as unsigned
  return ((unsigned long long)(~asc) << 32) | box;
}

__global__ void __launch_bounds__(1024)
nms_segment_kernel(const y3_cand* __restrict__ bucketed, const int* __restrict__ seg_off,
                   int cap, int C, double thr, const y3_thresholds* __restrict__ dyn,
                   y3_cand* __restrict__ sorted, uint8_t* __restrict__ keep, int* __restrict__ class_kept,
                   const int* __restrict__ list, const int* __restrict__ list_count) {
  pdl_enter();
  if (dyn) thr = dyn->iou_thresh;  // device-resident thresholds: one graph, any setting
  extern __shared__ __align__(16) uint8_t seg_smem[];
  __shared__ uint32_t kept_mask_s;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int items = *list_count;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    __syncthreads();  // the previous segment's shared state is dead
    const int id = list[item];
    const int img = id / C;
    const int seg = id - img * C;
    const int off = seg_off[(long long)img * (C + 1) + seg];
    const int n = seg_off[(long long)img * (C + 1) + seg + 1] - off;
    const y3_cand* src = bucketed + (long long)img * cap + off;
    y3_cand* out = sorted + (long long)img * cap + off;
    uint8_t* keep_out = keep + (long long)img * cap + off;

    // ---- 1. sort by (prob desc, box asc) -------------------------------------------------------------
    if (n <= SEG_SORT_MAX) {
      unsigned long long* skeys = reinterpret_cast<unsigned long long*>(seg_smem);
      int np = 1024;
      while (np < n) np <<= 1;
      for (int i = threadIdx.x; i < np; i += blockDim.x) {
        unsigned long long k = ~0ull;  // padding sorts last
        if (i < n) {
          const uint4 hi = reinterpret_cast<const uint4*>(src + i)[1];
          k = seg_key(hi.x, hi.z);
        }
        skeys[i] = k;
      }
      __syncthreads();
      for (int k = 2; k <= np; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = threadIdx.x; i < np; i += blockDim.x) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const unsigned long long a = skeys[i], b = skeys[ixj];
              const bool up = (i & k) == 0;
              if ((a > b) == up) { skeys[i] = b; skeys[ixj] = a; }
            }
          }
          __syncthreads();
        }
      }
      for (int i = threadIdx.x; i < n; i += blockDim.x) {  // rank of record i = position of its key
        const uint4 lo = reinterpret_cast<const uint4*>(src + i)[0];
        const uint4 hi = reinterpret_cast<const uint4*>(src + i)[1];
        const unsigned long long key = seg_key(hi.x, hi.z);
        int lo_i = 0, hi_i = n - 1;
        while (lo_i < hi_i) {
          const int mid = (lo_i + hi_i) >> 1;
          if (skeys[mid] < key) lo_i = mid + 1; else hi_i = mid;
        }
        reinterpret_cast<uint4*>(out + lo_i)[0] = lo;
        reinterpret_cast<uint4*>(out + lo_i)[1] = hi;
      }
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint4 lo = reinterpret_cast<const uint4*>(src + i)[0];
        const uint4 hi = reinterpret_cast<const uint4*>(src + i)[1];
        const float p = __uint_as_float(hi.x);
        const int b = (int)hi.z;
        int rank = 0;
        for (int j = 0; j < n; ++j) {
          const uint4 hj = reinterpret_cast<const uint4*>(src + j)[1];
          rank += before(__uint_as_float(hj.x), (int)hj.z, p, b) ? 1 : 0;
        }
        reinterpret_cast<uint4*>(out + rank)[0] = lo;
        reinterpret_cast<uint4*>(out + rank)[1] = hi;
      }
    }
    __syncthreads();  // global writes to `out` by this CTA are visible to it from here on; the keys are dead

    // ---- 2. greedy suppression, 32 pivots at a time --------------------------------------------------
    // in_smem: boxes as fp32 (x1, y1, x2+1, y2+1) + fp32 area + alive flags in shared memory; the IoU test is
    // the bitmask kernel's: fp32 with a 2^-18 guard band around the threshold, the reference's exact int64 /
    // fp64 expression (iou_gt on the int records in `out`) only inside the band or for huge / degenerate boxes.
    const bool in_smem = n <= SEG_SMEM_BOXES;
    float4* sbf = reinterpret_cast<float4*>(seg_smem);
    float* sar = reinterpret_cast<float*>(seg_smem + (size_t)SEG_SMEM_BOXES * 16);
    volatile uint8_t* alive = in_smem ? reinterpret_cast<volatile uint8_t*>(seg_smem + (size_t)SEG_SMEM_BOXES * 20)
                                      : reinterpret_cast<volatile uint8_t*>(keep_out);
    const double cd = thr / (1.0 + thr);
    const bool band_ok = thr > 1e-6 && thr < 1e6;  // the guard band assumes a positive finite threshold
    const float c_hi = band_ok ? (float)cd * (1.0f + 3.814697265625e-06f) : __int_as_float(0x7fc00000);  // 1 + 2^-18
    const float c_lo = band_ok ? (float)cd * (1.0f - 3.814697265625e-06f) : __int_as_float(0x7fc00000);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (in_smem) {
        const int4 v = reinterpret_cast<const int4*>(out + i)[0];
        sbf[i] = make_float4((float)v.x, (float)v.y, (float)v.z + 1.0f, (float)v.w + 1.0f);
        sar[i] = fast_area(v.x, v.y, v.z, v.w);
      }
      alive[i] = 1;
    }
    __syncthreads();
    // does box i (earlier in score order) suppress box j?
    auto suppresses = [&](int i, int j) -> bool {
      if (in_smem) {
        const float4 pv = sbf[i], mb = sbf[j];
        const float s2 = sar[i] + sar[j];
        const float iw = fminf(pv.z, mb.z) - fmaxf(pv.x, mb.x);
        const float ih = fminf(pv.w, mb.w) - fmaxf(pv.y, mb.y);
        const float fi = fmaxf(iw, 0.f) * fmaxf(ih, 0.f);
        if (fi > c_hi * s2) return true;
        if (fi <= c_lo * s2) return false;
      }
      const int4 a = reinterpret_cast<const int4*>(out + i)[0];
      const int4 b = reinterpret_cast<const int4*>(out + j)[0];
      return iou_gt(Box4{a.x, a.y, a.z, a.w}, Box4{b.x, b.y, b.z, b.w}, thr);
    };
    __shared__ uint32_t s_sup[32];  // s_sup[w]: which boxes of the chunk pivot w would suppress
    for (int c0 = 0; c0 < n; c0 += 32) {
      {  // 32 x 32 block among the chunk's boxes: warp w = pivot c0 + w, lane l = box c0 + l (blockDim = 1024)
        const int i = c0 + warp, j = c0 + lane;
        const bool s = warp < lane && j < n && suppresses(i, j);
        const uint32_t word = __ballot_sync(0xffffffffu, s);
        if (lane == 0) s_sup[warp] = word;
      }
      __syncthreads();
      if (warp == 0) {
        const int j = c0 + lane;
        const bool valid = j < n;
        const bool was_alive = valid && alive[valid ? j : 0];
        uint32_t sup = 0;  // bit w: pivot w (earlier in order) would suppress me
#pragma unroll
        for (int w = 0; w < 32; ++w) sup |= ((s_sup[w] >> lane) & 1u) << w;
        uint32_t kept = 0;
        for (int i = 0; i < 32; ++i) {
          const bool k = (lane == i) && was_alive && ((sup & kept) == 0);
          kept |= __ballot_sync(0xffffffffu, k);
        }
        if (valid) alive[j] = (kept >> lane) & 1u;
        if (lane == 0) kept_mask_s = kept;
      }
      __syncthreads();
      const uint32_t kept = kept_mask_s;
      if (kept != 0) {
        for (int j = c0 + 32 + threadIdx.x; j < n; j += blockDim.x) {
          if (!alive[j]) continue;
          uint32_t m = kept;
          while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1;
            if (suppresses(c0 + i, j)) { alive[j] = 0; break; }
          }
        }
      }
      __syncthreads();
    }
    // ---- 3. flags out, kept count -----------------------------------------------------------------------
    if (threadIdx.x == 0) kept_mask_s = 0u;
    __syncthreads();
    int local = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const uint8_t k = alive[j];
      if (in_smem) keep_out[j] = k;
      local += k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0 && local) atomicAdd(&kept_mask_s, (uint32_t)local);
    __syncthreads();
    if (class_kept && threadIdx.x == 0) class_kept[(long long)img * C + seg] = (int)kept_mask_s;
  }  // work list
}

// Segments of up to 512 boxes — every per-class segment of a real detector output — in three
// size classes (<= 128, <= 256, <= 512 boxes), one CTA of MAXN threads per (image, class):
//   1. rank sort by (prob desc, box asc) through shared memory; the sorted boxes are kept there as
//      fp32 (x1, y1, x2+1, y2+1) + area;
//   2. warp w owns the 32 boxes of column block w in registers and walks the pivots i < 32(w+1): the
//      pivot is a shared-memory broadcast, 32 lanes test at once, __ballot_sync makes the word
//      "pivot i suppresses boxes 32w..32w+31";
//   3. one warp resolves the greedy order from the bit matrix, 32 boxes per step.
// The IoU decision is made in fp32 with a 2^-18 guard band around the threshold (coordinates below
// 2^22 are exact in fp32; the accumulated rounding of inter / union stays under 2^-20); only pairs
// inside the band, or involving a box with huge or degenerate coordinates (area stored as NaN, which
// fails both band comparisons), evaluate the reference's int64 / float64 expression (iou_gt) — the
// kept set stays bit-exact.  Larger segments are left to nms_segment_kernel.

__device__ __forceinline__ float fast_area(int x1, int y1, int x2, int y2) {
  const bool small = x1 > -4194304 && x1 <= x2 && x2 < 4194304 && y1 > -4194304 && y1 <= y2 && y2 < 4194304;
  return small ? __fmul_rn((float)(x2 - x1 + 1), (float)(y2 - y1 + 1)) : __int_as_float(0x7fc00000);
}

template <int MAXN>
__global__ void __launch_bounds__(MAXN)
nms_bitmask_kernel(const y3_cand* __restrict__ bucketed, const int* __restrict__ seg_off, int cap, int C,
                   double thr, const y3_thresholds* __restrict__ dyn, y3_cand* __restrict__ sorted,
                   uint8_t* __restrict__ keep, int* __restrict__ class_kept, const int* __restrict__ list,
                   const int* __restrict__ list_count) {
  pdl_enter();
  if (dyn) thr = dyn->iou_thresh;
  constexpr int WORDS = MAXN / 32;
  __shared__ float4 sboxf[MAXN];
  __shared__ float sarea[MAXN];
  __shared__ uint2 skey[MAXN];
  __shared__ uint32_t smask[WORDS][MAXN + 1];  // [w][i]; +1: the scan reads one column across w
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, t = threadIdx.x;
  const int items = *list_count;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
  __syncthreads();  // the previous segment's shared state is dead
  const int id = list[item];
  const int img = id / C;
  const int seg = id - img * C;
  const int off = seg_off[(long long)img * (C + 1) + seg];
  const int n = seg_off[(long long)img * (C + 1) + seg + 1] - off;  // 1 .. MAXN by construction of the list
  const y3_cand* src = bucketed + (long long)img * cap + off;
  y3_cand* out = sorted + (long long)img * cap + off;
  uint8_t* keep_out = keep + (long long)img * cap + off;

  // ---- 1. rank sort -------------------------------------------------------------------------------
  uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
  if (t < n) {
    lo = reinterpret_cast<const uint4*>(src + t)[0];
    hi = reinterpret_cast<const uint4*>(src + t)[1];
    skey[t] = make_uint2(hi.x, hi.z);
  }
  __syncthreads();
  if (t < n) {
    const float p = __uint_as_float(hi.x);
    const int b = (int)hi.z;
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      const uint2 kj = skey[j];
      rank += before(__uint_as_float(kj.x), (int)kj.y, p, b) ? 1 : 0;
    }
    reinterpret_cast<uint4*>(out + rank)[0] = lo;
    reinterpret_cast<uint4*>(out + rank)[1] = hi;
    const int x1 = (int)lo.x, y1 = (int)lo.y, x2 = (int)lo.z, y2 = (int)lo.w;
    sboxf[rank] = make_float4((float)x1, (float)y1, (float)x2 + 1.0f, (float)y2 + 1.0f);
    sarea[rank] = fast_area(x1, y1, x2, y2);
  }
  __syncthreads();

  // ---- 2. bit matrix ------------------------------------------------------------------------------
  // Work items (w, c), c <= w: the 32 boxes of column block w against the 32 pivots of chunk c, dealt
  // round-robin to ALL warps of the CTA (a warp-per-column split would leave the first warps idle:
  // column w meets 32 (w + 1) pivots).  iou > thr  <=>  inter > thr / (1 + thr) * (area_i + area_j).
  const int words = (n + 31) >> 5;
  {
    const double cd = thr / (1.0 + thr);
    const bool band_ok = thr > 1e-6 && thr < 1e6;  // the guard band assumes a positive finite threshold
    const float c_hi = band_ok ? (float)cd * (1.0f + 3.814697265625e-06f) : __int_as_float(0x7fc00000);  // 1 + 2^-18
    const float c_lo = band_ok ? (float)cd * (1.0f - 3.814697265625e-06f) : __int_as_float(0x7fc00000);
    const int items = words * (words + 1) / 2;
    for (int item = warp; item < items; item += MAXN / 32) {
      int w = 0, c = item;
      while (c > w) { c -= w + 1; ++w; }
      const int j = 32 * w + lane;
      const bool jv = j < n;
      const float4 mb = jv ? sboxf[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float ma = jv ? sarea[j] : 1.0f;
      const int i0 = 32 * c;
      const int rn = min(32, n - i0);
      const int jrel = jv ? j - i0 : 0;  // pivot r precedes my box iff r < jrel
      uint32_t myword = 0;               // lane r keeps the word of pivot i0 + r
#pragma unroll 4
      for (int r = 0; r < rn; ++r) {
        const float4 pv = sboxf[i0 + r];
        const float s2 = sarea[i0 + r] + ma;
        const float iw = fminf(pv.z, mb.z) - fmaxf(pv.x, mb.x);
        const float ih = fminf(pv.w, mb.w) - fmaxf(pv.y, mb.y);
        const float fi = fmaxf(iw, 0.f) * fmaxf(ih, 0.f);
        const bool live = r < jrel;
        bool sup = live && fi > c_hi * s2;
        // inside the guard band, NaN area, or odd threshold: the reference's exact expression
        const bool band = live && !sup && !(fi <= c_lo * s2);
        if (__any_sync(0xffffffffu, band)) {
          if (band) {
            const int4 a = reinterpret_cast<const int4*>(out + i0 + r)[0];
            const int4 b = reinterpret_cast<const int4*>(out + j)[0];
            sup = iou_gt(Box4{a.x, a.y, a.z, a.w}, Box4{b.x, b.y, b.z, b.w}, thr);
          }
        }
        const uint32_t word = __ballot_sync(0xffffffffu, sup);
        if (lane == r) myword = word;
      }
      if (lane < rn) smask[w][i0 + lane] = myword;
    }
  }
  __syncthreads();

  // ---- 3. greedy scan, 32 boxes per step ------------------------------------------------------------
  if (warp == 0) {
    uint32_t removed = 0;  // lane w: bits [32w, 32w+32) of the removed set
    for (int c = 0; c < words; ++c) {
      const int i0 = 32 * c;
      const int cnt = min(32, n - i0);
      // diagonal block: lane r holds which boxes of this chunk box i0+r suppresses
      const uint32_t diag = (lane < cnt) ? smask[c][i0 + lane] : 0u;
      uint32_t rem = __shfl_sync(0xffffffffu, removed, c);
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const uint32_t dt = __shfl_sync(0xffffffffu, diag, r);  // independent of the chain on `rem`
        if (!((rem >> r) & 1u)) rem |= dt;
      }
      if (lane == c) removed = rem;
      uint32_t kept = ~rem & (cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u));  // warp-uniform
      if (lane > c && lane < words) {
        while (kept) {
          const int r = __ffs(kept) - 1;
          kept &= kept - 1;
          removed |= smask[lane][i0 + r];
        }
      }
    }
    int nkept = 0;
    for (int b = 0; b < 32; ++b) {
      const int i = 32 * lane + b;
      if (i < n) {
        const int k = ((removed >> b) & 1u) ? 0 : 1;
        keep_out[i] = (uint8_t)k;
        nkept += k;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nkept += __shfl_xor_sync(0xffffffffu, nkept, o);
    if (class_kept && lane == 0) class_kept[(long long)img * C + seg] = nkept;
  }
  }  // work list
}

// a17: the three arrays `inference` returns per image (yolov3/inference.py:360-366), written
// straight in their final dtypes and final order: one warp per (image, class) segment compacts
// the segment's kept records (already prob-descending) to dst_off[image, class], the position the
// host assigned to that class group (class groups follow the reference's set() visiting order).
__global__ void __launch_bounds__(128)
emit_detections_kernel(const y3_cand* __restrict__ sorted, const uint8_t* __restrict__ keep,
                       const int* __restrict__ class_start, const int* __restrict__ dst_off, int n_images,
                       int cap, int C, long long* __restrict__ tlbr, float* __restrict__ prob,
                       long long* __restrict__ cls) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)n_images * C) return;
  const int img = (int)(wid / C);
  const int c = (int)(wid - (long long)img * C);
  const int off = class_start[(long long)img * (C + 1) + c];
  const int n = class_start[(long long)img * (C + 1) + c + 1] - off;
  long long dst = dst_off[wid];
  if (n <= 0 || dst < 0) return;
  const y3_cand* src = sorted + (long long)img * cap + off;
  const uint8_t* kp = keep + (long long)img * cap + off;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const bool k = i < n && kp[i] != 0;
    const uint32_t ball = __ballot_sync(0xffffffffu, k);
    if (k) {
      const long long o = dst + __popc(ball & ((1u << lane) - 1u));
      const uint4 lo = reinterpret_cast<const uint4*>(src + i)[0];
      const uint4 hi = reinterpret_cast<const uint4*>(src + i)[1];
      longlong2* t = reinterpret_cast<longlong2*>(tlbr + 4 * o);
      t[0] = make_longlong2((long long)(int)lo.x, (long long)(int)lo.y);
      t[1] = make_longlong2((long long)(int)lo.z, (long long)(int)lo.w);
      prob[o] = __uint_as_float(hi.x);
      cls[o] = (long long)(int)hi.y;
    }
    dst += __popc(ball);
  }
}

// Destinations of the (image, class) groups for emit_detections_kernel with classes ascending inside an
// image: exclusive scan of class_kept in (image, class) order by ONE CTA (n*C <= 2^20 values: each thread
// sums a contiguous run, the run totals are scanned through shared memory), per-image totals on the side.
__global__ void __launch_bounds__(1024)
plan_destinations_kernel(const int* __restrict__ class_kept, int n_images, int C, int* __restrict__ dst_off,
                         int* __restrict__ det_counts) {
  pdl_enter();
  __shared__ int warp_tot[32];
  const int total = n_images * C;
  const int per = (total + blockDim.x - 1) / blockDim.x;
  const int lo = min(total, (int)threadIdx.x * per), hi = min(total, lo + per);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += class_kept[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    warp_tot[lane] = w;  // inclusive totals of warps 0..lane
  }
  __syncthreads();
  int run = incl - sum + (warp ? warp_tot[warp - 1] : 0);
  for (int i = lo; i < hi; ++i) { dst_off[i] = run; run += class_kept[i]; }
  if (threadIdx.x == 0) det_counts[n_images] = warp_tot[(blockDim.x >> 5) - 1];
  for (int img = threadIdx.x; img < n_images; img += blockDim.x) {
    int k = 0;
    for (int c = 0; c < C; ++c) k += class_kept[img * C + c];
    det_counts[img] = k;
  }
}

// Ordered compaction of kept records per image; dets are written image after image into one
// flat array, det_offsets[img] .. det_offsets[img+1] (exclusive scan done by nms_offsets_kernel).
__global__ void __launch_bounds__(1024)
nms_count_kernel(const uint8_t* __restrict__ keep, const int* __restrict__ counts, int cap,
                 int* __restrict__ det_counts) {
  pdl_enter();
  __shared__ int total;
  const int img = blockIdx.x;
  int n = counts[img];
  n = n < cap ? n : cap;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) local += keep[(long long)img * cap + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&total, local);
  __syncthreads();
  if (threadIdx.x == 0) det_counts[img] = total;
}

__global__ void __launch_bounds__(1024)
nms_compact_kernel(const y3_cand* __restrict__ sorted, const uint8_t* __restrict__ keep,
                   const int* __restrict__ counts, int cap, y3_cand* __restrict__ dets,
                   const int* __restrict__ det_counts, int n_images, int flat) {
  pdl_enter();
  __shared__ int warp_sums[32];
  __shared__ int base_s;
  const int img = blockIdx.x;
  int n = counts[img];
  n = n < cap ? n : cap;
  // destination base: flat -> exclusive sum of det_counts of earlier images; else img*cap
  if (threadIdx.x == 0) {
    long long b = 0;
    if (flat) { for (int i = 0; i < img; ++i) b += det_counts[i]; }
    else b = (long long)img * cap;
    base_s = (int)b;
  }
  __syncthreads();
  int running = base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const int k = (i < n) ? keep[(long long)img * cap + i] : 0;
    const uint32_t ball = __ballot_sync(0xffffffffu, k);
    const int prefix = __popc(ball & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[warp] = __popc(ball);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < nwarps; ++w) { const int s = warp_sums[w]; if (w < warp) woff += s; tot += s; }
    if (k) {
      const y3_cand* s = sorted + (long long)img * cap + i;
      y3_cand* d = dets + running + woff + prefix;
      reinterpret_cast<uint4*>(d)[0] = reinterpret_cast<const uint4*>(s)[0];
      reinterpret_cast<uint4*>(d)[1] = reinterpret_cast<const uint4*>(s)[1];
    }
    running += tot;
    __syncthreads();
  }
}

}  // namespace y3

using namespace y3;

extern "C" {

size_t y3_nms_workspace_bytes(int32_t n, int32_t cap, int32_t num_classes) {
  if (n <= 0 || cap <= 0 || num_classes <= 0) return 0;
  const size_t bucketed = (size_t)n * cap * sizeof(y3_cand);
  const size_t seg = (size_t)n * ((size_t)num_classes + 1) * sizeof(int32_t);
  const size_t lists = (size_t)NUM_LISTS * n * num_classes * sizeof(int32_t);
  return bucketed + ((seg + 255) / 256) * 256 + ((lists + 255) / 256) * 256 + 512;
}

int y3_nms(const y3_cand* cands, const int32_t* counts, int32_t n, int32_t cap, int32_t num_classes,
           double iou_thresh, const y3_thresholds* dev_thresholds, int32_t per_class, y3_cand* sorted,
           uint8_t* keep, int32_t* class_first_box,
           int32_t* class_start, int32_t* class_kept, void* workspace, size_t workspace_bytes, void* stream) {
  Y3_CHECK_ARG(cands && counts && sorted && keep && workspace, "nms: null argument");
  Y3_CHECK_ARG(n > 0 && cap > 0, "nms: bad n=%d cap=%d", n, cap);
  Y3_CHECK_ARG(num_classes > 0 && num_classes <= MAX_CLASSES, "nms: num_classes=%d out of range (1..%d)",
               num_classes, MAX_CLASSES);
  Y3_CHECK_ARG((reinterpret_cast<uintptr_t>(cands) & 15) == 0 && (reinterpret_cast<uintptr_t>(sorted) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "nms: buffers must be 16-byte aligned");
  if (workspace_bytes < y3_nms_workspace_bytes(n, cap, num_classes)) {
    y3::set_error("nms: workspace %zu bytes < required %zu", workspace_bytes,
                  y3_nms_workspace_bytes(n, cap, num_classes));
    return Y3_EWORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int C = per_class ? num_classes : 1;
  y3_cand* bucketed = reinterpret_cast<y3_cand*>(workspace);
  uint8_t* const ws8 = reinterpret_cast<uint8_t*>(workspace);
  int* seg_off = reinterpret_cast<int*>(ws8 + (size_t)n * cap * sizeof(y3_cand));
  const size_t seg_bytes = (((size_t)n * ((size_t)C + 1) * sizeof(int32_t)) + 255) / 256 * 256;
  int* lists = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(seg_off) + seg_bytes);
  const int list_stride = n * C;
  int* list_counts = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(lists) +
                                            (((size_t)NUM_LISTS * list_stride * sizeof(int32_t)) + 255) / 256 * 256);
  Y3_CUDA_OK(cudaMemsetAsync(list_counts, 0, NUM_LISTS * sizeof(int), s));

  Y3_CUDA_OK(launch_kernel(nms_bucket_kernel, dim3(n), dim3(1024), 0, s, cands, counts, cap, num_classes, per_class, bucketed, seg_off,
                           class_first_box, class_start, class_kept, lists, list_counts, list_stride));
  Y3_LAUNCH_OK("nms_bucket_kernel");
  const long long segs = (long long)n * C;
  auto grid_for = [&](int per_sm) { const long long g = (long long)num_sms() * per_sm; return (int)(segs < g ? segs : g); };

  // segments of <= 512 boxes: three size classes of the fp32 bit-matrix kernel; the rest (if any):
  // nms_segment_kernel.  Each launch is a persistent grid over the work list of its size class.
  Y3_CUDA_OK(launch_kernel(nms_bitmask_kernel<128>, dim3(grid_for(16)), dim3(128), 0, s, bucketed, seg_off, cap, C, iou_thresh,
                           dev_thresholds, sorted, keep, class_kept, lists, list_counts));
  Y3_LAUNCH_OK("nms_bitmask_kernel<128>");
  Y3_CUDA_OK(launch_kernel(nms_bitmask_kernel<256>, dim3(grid_for(8)), dim3(256), 0, s, bucketed, seg_off, cap, C, iou_thresh,
                           dev_thresholds, sorted, keep, class_kept, lists + list_stride, list_counts + 1));
  Y3_LAUNCH_OK("nms_bitmask_kernel<256>");
  Y3_CUDA_OK(launch_kernel(nms_bitmask_kernel<512>, dim3(grid_for(4)), dim3(512), 0, s, bucketed, seg_off, cap, C, iou_thresh,
                           dev_thresholds, sorted, keep, class_kept, lists + 2 * list_stride, list_counts + 2));
  Y3_LAUNCH_OK("nms_bitmask_kernel<512>");
  static bool seg_attr = false;
  if (!seg_attr) {
    Y3_CUDA_OK(cudaFuncSetAttribute(nms_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SEG_SMEM_BYTES));
    seg_attr = true;
  }
  Y3_CUDA_OK(launch_kernel(nms_segment_kernel, dim3(grid_for(1)), dim3(1024), (size_t)SEG_SMEM_BYTES, s, bucketed, seg_off, cap,
                           C, iou_thresh, dev_thresholds, sorted, keep, class_kept, lists + 3 * list_stride,
                           list_counts + 3));
  Y3_LAUNCH_OK("nms_segment_kernel");
  return Y3_OK;
}

int y3_plan_destinations(const int32_t* class_kept, int32_t n, int32_t num_segments, int32_t* dst_off,
                         int32_t* det_counts, void* stream) {
  Y3_CHECK_ARG(class_kept && dst_off && det_counts, "plan_destinations: null argument");
  Y3_CHECK_ARG(n > 0 && num_segments > 0 && (long long)n * num_segments <= (1ll << 20),
               "plan_destinations: n=%d x segments=%d out of range", n, num_segments);
  Y3_CUDA_OK(launch_kernel(plan_destinations_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, class_kept, n,
                           num_segments, dst_off, det_counts));
  Y3_LAUNCH_OK("plan_destinations_kernel");
  return Y3_OK;
}

int y3_emit_detections(const y3_cand* sorted, const uint8_t* keep, const int32_t* class_start,
                       const int32_t* dst_off, int32_t n, int32_t cap, int32_t num_segments, int64_t* tlbr,
                       float* prob, int64_t* cls, void* stream) {
  Y3_CHECK_ARG(sorted && keep && class_start && dst_off && tlbr && prob && cls, "emit_detections: null argument");
  Y3_CHECK_ARG(n > 0 && cap > 0 && num_segments > 0, "emit_detections: bad n=%d cap=%d segments=%d", n, cap,
               num_segments);
  Y3_CHECK_ARG((reinterpret_cast<uintptr_t>(tlbr) & 15) == 0, "emit_detections: tlbr must be 16-byte aligned");
  const long long warps = (long long)n * num_segments;
  const int grid = (int)((warps + 3) / 4);
  Y3_CUDA_OK(launch_kernel(emit_detections_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, sorted, keep,
                           class_start, dst_off, n, cap, num_segments, reinterpret_cast<long long*>(tlbr), prob,
                           reinterpret_cast<long long*>(cls)));
  Y3_LAUNCH_OK("emit_detections_kernel");
  return Y3_OK;
}

int y3_compact_kept(const y3_cand* sorted, const uint8_t* keep, const int32_t* counts, int32_t n, int32_t cap,
                    y3_cand* dets, int32_t* det_counts, int32_t flat, void* stream) {
  Y3_CHECK_ARG(sorted && keep && counts && dets && det_counts, "compact_kept: null argument");
  Y3_CHECK_ARG(n > 0 && cap > 0, "compact_kept: bad n=%d cap=%d", n, cap);
  cudaStream_t s = (cudaStream_t)stream;
  Y3_CUDA_OK(launch_kernel(nms_count_kernel, dim3(n), dim3(1024), 0, s, keep, counts, cap, det_counts));
  Y3_LAUNCH_OK("nms_count_kernel");
  Y3_CUDA_OK(launch_kernel(nms_compact_kernel, dim3(n), dim3(1024), 0, s, sorted, keep, counts, cap, dets, det_counts, n, flat));
  Y3_LAUNCH_OK("nms_compact_kernel");
  return Y3_OK;
}

}  // extern "C"
