// decode.cu — YOLO head decode (yolov3/darknet.py:48-122, :390-399) and the per-image
// post-processing of yolov3/inference.py:342-353 + cxywh_to_tlbr (:269-283), fused.
//
// A warp decodes 32 consecutive boxes (image, anchor, row, col): a box's 5+classes logits are
// contiguous in the NHWC float32 head tensor; eight lanes reduce one box (max/argmax and the softmax
// denominator with shuffles), then lane b finishes box b.
// All fp32 steps keep the reference's operation order (sigmoid, +offset, /grid; exp, *anchor,
// /train size, *image size) with explicit round-to-nearest intrinsics so no FMA contraction
// changes a rounding; only sigmoid/exp themselves may differ from torch's by an ulp.
#include "common.cuh"
#include "decode_math.cuh"

namespace y3 {

static constexpr int CAND_BOXES = 256;  // boxes per CTA of both decode kernels: 32 per warp

// Phase 1 of both decode kernels (cooperative): a warp owns the 32 consecutive boxes m_warp .. m_warp+31
// of image `img` (nb of them exist).  EIGHT lanes share a box, so the warp reduces four boxes at a time:
// lane `sub` of a group loads fields sub, sub+8, ... (all loads of a box are issued before the first use —
// up to 11 per lane are kept in registers, enough for 5+80 fields; wider heads re-read the remainder),
// then max / argmax / softmax denominator are reduced over the 8 lanes with three shuffle steps and the
// raw fields of box b are parked in lane b.
static constexpr int DEC_CACHED = 11;  // fields cached per lane: 8 * 11 = 88 >= 5 + 80

__device__ __forceinline__ void decode_phase1(const y3_head_desc& d, const float* __restrict__ logits, int img,
                                              int m_warp, int nb, int lane, float& tx, float& ty, float& tw,
                                              float& th, float& to, float& sum, int& cls) {
  const int sub = lane & 7, grp = lane >> 3;
  const int cells = d.g_h * d.g_w;
  const int fields = 5 + d.num_classes;
  tx = ty = tw = th = to = 0.f;
  sum = 1.f;
  cls = 0;
  for (int it = 0; it * 4 < nb; ++it) {
    const int b = it * 4 + grp;
    const bool live = b < nb;
    const int m = m_warp + (live ? b : 0);
    const int a = m / cells;
    const int cell = m - a * cells;  // row * g_w + col: pixels are contiguous in NHWC
    const float* px = logits + ((long long)img * cells + cell) * d.ld + a * fields;
    float v[DEC_CACHED];
#pragma unroll
    for (int i = 0; i < DEC_CACHED; ++i) {
      const int f = sub + 8 * i;
      v[i] = (live && f < fields) ? __ldg(px + f) : -INFINITY;
    }
    float best = -INFINITY;
    int best_idx = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < DEC_CACHED; ++i) {
      const int f = sub + 8 * i;
      if (f >= 5 && f < fields && v[i] > best) { best = v[i]; best_idx = f - 5; }
    }
    for (int f = sub + 8 * DEC_CACHED; f < fields; f += 8) {
      const float x = live ? __ldg(px + f) : -INFINITY;
      if (x > best) { best = x; best_idx = f - 5; }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {  // first index wins ties, like torch.max
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
      if (ob > best || (ob == best && oi < best_idx)) { best = ob; best_idx = oi; }
    }
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < DEC_CACHED; ++i) {
      const int f = sub + 8 * i;
      if (f >= 5 && f < fields) part += expf(v[i] - best);
    }
    for (int f = sub + 8 * DEC_CACHED; f < fields; f += 8) part += live ? expf(__ldg(px + f) - best) : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    // lane b (b = 4*it + g) takes box b's raw fields: tx..to sit in lanes 8g .. 8g+4 of v[0]
    const int g8 = (lane & 3) * 8;
    const float h0 = __shfl_sync(0xffffffffu, v[0], g8), h1 = __shfl_sync(0xffffffffu, v[0], g8 + 1);
    const float h2 = __shfl_sync(0xffffffffu, v[0], g8 + 2), h3 = __shfl_sync(0xffffffffu, v[0], g8 + 3);
    const float h4 = __shfl_sync(0xffffffffu, v[0], g8 + 4);
    const float ps = __shfl_sync(0xffffffffu, part, g8);
    const int pc = __shfl_sync(0xffffffffu, best_idx, g8);
    if ((lane >> 2) == it) { tx = h0; ty = h1; tw = h2; th = h3; to = h4; sum = ps; cls = pc; }
  }

}

// Dense outputs of Darknet.forward (darknet.py:401-405).  Same two phases as decode_cands_kernel: the
// warp reduces its 32 boxes cooperatively (all loads of four boxes in flight at once), then lane b
// finishes box b, so the three output arrays are written with coalesced 512 / 128 / 256-byte requests.
__global__ void __launch_bounds__(256)
decode_dense_kernel(const y3_head_desc d, const float* __restrict__ logits, float* __restrict__ bbox,
                    float* __restrict__ prob, long long* __restrict__ cls_out) {
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int img = blockIdx.y;
  const int cells = d.g_h * d.g_w;
  const int per_img = d.num_anchors * cells;
  const int m_warp = blockIdx.x * CAND_BOXES + warp * 32;
  const int nb = min(32, per_img - m_warp);
  if (nb <= 0) return;
  float tx, ty, tw, th, to, sum;
  int cls;
  decode_phase1(d, logits, img, m_warp, nb, lane, tx, ty, tw, th, to, sum, cls);
  if (lane < nb) {
    const int m = m_warp + lane;  // a*cells + row*g_w + col  (darknet.py:118-120)
    const int a = m / cells;
    const int cell = m - a * cells;
    const int row = cell / d.g_w;
    const int col = cell - row * d.g_w;
    const float x = __fdiv_rn(__fadd_rn(sigmoidf_ref(tx), (float)col), (float)d.g_w);
    const float y = __fdiv_rn(__fadd_rn(sigmoidf_ref(ty), (float)row), (float)d.g_h);
    const float w = __fdiv_rn(__fmul_rn(expf(tw), d.anchor_w[a]), d.train_w);
    const float h = __fdiv_rn(__fmul_rn(expf(th), d.anchor_h[a]), d.train_h);
    // softmax value of the arg-max class is exp(0)/sum; then * sigmoid(objectness)  (darknet.py:104-108)
    const float pr = __fmul_rn(__fdiv_rn(1.0f, sum), sigmoidf_ref(to));
    const long long g = (long long)img * d.boxes_per_image + d.box_offset + m;
    reinterpret_cast<float4*>(bbox)[g] = make_float4(x, y, w, h);
    prob[g] = pr;
    cls_out[g] = (long long)cls;
  }
}

// Fused decode + threshold + pixel scaling + truncation + tl/br + compaction.
// One CTA owns CAND_BOXES consecutive boxes of ONE image, 32 per warp.  Phase 1 (cooperative):
// EIGHT lanes share a box, so a warp reduces four boxes at a time: lane `sub` of a group loads
// fields sub, sub+8, ... (all loads of a box are issued before the first use — up to 11 per lane
// are kept in registers, enough for 5+80 fields; wider heads re-read the remainder), then max /
// argmax / softmax denominator are reduced over the 8 lanes with three shuffle steps and the raw
// fields of box b are parked in lane b.  Phase 2 (lane-parallel): lane b finishes box b (sigmoid,
// exp, scaling, truncation) — 32 boxes per instruction.  Passing boxes are collected in shared
// memory; the CTA reserves its output range with a single global atomic (per-box atomics on the
// per-image counters serialise in L2).

__global__ void __launch_bounds__(256)
decode_cands_kernel(const y3_head_desc d, const float* __restrict__ logits, float prob_thresh,
                    const y3_thresholds* __restrict__ dyn, const int* __restrict__ orig_hw,
                    y3_cand* __restrict__ cands, int* __restrict__ counts, int cap) {
  pdl_enter();
  if (dyn) prob_thresh = dyn->prob_thresh;  // device-resident thresholds: one graph, any setting
  __shared__ uint4 s_rec[CAND_BOXES][2];
  __shared__ int s_count, s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int img = blockIdx.y;
  const int cells = d.g_h * d.g_w;
  const int per_img = d.num_anchors * cells;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();

  // ---- phase 1 (cooperative): box b of this warp's 32 ends up in lane b --------------------------
  const int m_warp = blockIdx.x * CAND_BOXES + warp * 32;
  float tx, ty, tw, th, to, sum;
  int cls;
  const int nb = min(32, per_img - m_warp);
  decode_phase1(d, logits, img, m_warp, nb, lane, tx, ty, tw, th, to, sum, cls);

  // ---- phase 2: lane b finishes box b ---------------------------------------------------------
  if (lane < nb) {
    const int m = m_warp + lane;
    const int a = m / cells;
    const int cell = m - a * cells;
    const int row = cell / d.g_w;
    const int col = cell - row * d.g_w;
    // softmax value of the arg-max class is exp(0)/sum; then * sigmoid(objectness)  (darknet.py:104-108)
    const float prob = __fmul_rn(__fdiv_rn(1.0f, sum), sigmoidf_ref(to));
    if (prob >= prob_thresh) {  // inference.py:342
      uint4 rlo, rhi;
      make_cand(tx, ty, tw, th, prob, cls, d.box_offset + m, row, col, d.g_h, d.g_w, d.anchor_w[a], d.anchor_h[a],
                d.train_w, d.train_h, (float)orig_hw[2 * img], (float)orig_hw[2 * img + 1], rlo, rhi);
      const int slot = atomicAdd(&s_count, 1);
      s_rec[slot][0] = rlo;
      s_rec[slot][1] = rhi;
    }
  }
  __syncthreads();
  const int n = s_count;
  if (n == 0) return;
  if (threadIdx.x == 0) s_base = atomicAdd(counts + img, n);
  __syncthreads();
  const int base = s_base;
  uint4* dst = reinterpret_cast<uint4*>(cands + (long long)img * cap);
  for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) {
    const int slot = base + (i >> 1);
    if (slot < cap) dst[2 * (long long)slot + (i & 1)] = s_rec[i >> 1][i & 1];
  }
}

static int check_head(const y3_head_desc* d, const float* logits) {
  Y3_CHECK_ARG(d && logits, "decode: null argument");
  Y3_CHECK_ARG(d->n > 0 && d->g_h > 0 && d->g_w > 0, "decode: bad grid");
  Y3_CHECK_ARG(d->num_anchors > 0 && d->num_anchors <= 8, "decode: num_anchors=%d out of range", d->num_anchors);
  Y3_CHECK_ARG(d->num_classes > 0 && d->num_classes <= 1024, "decode: num_classes=%d out of range", d->num_classes);
  Y3_CHECK_ARG(d->ld >= d->num_anchors * (5 + d->num_classes), "decode: ld=%d too small", d->ld);
  Y3_CHECK_ARG(d->box_offset >= 0 && d->boxes_per_image >= d->box_offset + d->num_anchors * d->g_h * d->g_w,
               "decode: box_offset/boxes_per_image inconsistent");
  Y3_CHECK_ARG(d->train_w > 0 && d->train_h > 0, "decode: bad train size");
  return Y3_OK;
}

}  // namespace y3

using namespace y3;

extern "C" {

int y3_yolo_decode_dense(const y3_head_desc* d, const float* logits, float* bbox_xywh, float* class_prob,
                         int64_t* class_idx, void* stream) {
  int rc = check_head(d, logits);
  if (rc != Y3_OK) return rc;
  Y3_CHECK_ARG(bbox_xywh && class_prob && class_idx, "decode_dense: null output");
  Y3_CHECK_ARG((reinterpret_cast<uintptr_t>(bbox_xywh) & 15) == 0, "decode_dense: bbox must be 16-byte aligned");
  const int per_img = d->num_anchors * d->g_h * d->g_w;
  Y3_CUDA_OK(launch_kernel(decode_dense_kernel, dim3((per_img + CAND_BOXES - 1) / CAND_BOXES, d->n), dim3(256), 0,
                           (cudaStream_t)stream, *d, logits, bbox_xywh, class_prob,
                           reinterpret_cast<long long*>(class_idx)));
  Y3_LAUNCH_OK("decode_dense_kernel");
  return Y3_OK;
}

int y3_yolo_decode_cands(const y3_head_desc* d, const float* logits, float prob_thresh,
                         const y3_thresholds* dev_thresholds, const int32_t* orig_hw, y3_cand* cands,
                         int32_t* counts, int32_t cap, void* stream) {
  int rc = check_head(d, logits);
  if (rc != Y3_OK) return rc;
  Y3_CHECK_ARG(orig_hw && cands && counts && cap > 0, "decode_cands: bad output arguments");
  Y3_CHECK_ARG((reinterpret_cast<uintptr_t>(cands) & 15) == 0, "decode_cands: cands must be 16-byte aligned");
  const int per_img = d->num_anchors * d->g_h * d->g_w;
  const dim3 grid((per_img + CAND_BOXES - 1) / CAND_BOXES, d->n);
  Y3_CUDA_OK(launch_kernel(decode_cands_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, *d, logits, prob_thresh,
                           dev_thresholds, orig_hw, cands, counts, cap));
  Y3_LAUNCH_OK("decode_cands_kernel");
  return Y3_OK;
}

}  // extern "C"
