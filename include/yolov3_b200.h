/*
 * yolov3_b200.h — C ABI of libyolov3_b200.so (hand-written sm_100a CUDA).
 *
 * The reference (nrsyed/pytorch-yolov3) has no FFI / plugin interface of its
 * own: its hot path is pure Python calling PyTorch and NumPy.
 * This header is therefore the NEW boundary that sits where those library
 * calls sit today.  Every entry point below names the reference call site
 * (file:line under the reference tree) whose arithmetic it replaces.
 *
 * Conventions
 *   - plain C types only; device buffers are raw pointers owned by the caller
 *     (PyTorch allocates them); the library never
 *     allocates or frees device memory; its only mutable state is per calling
 *     thread (error message, launch counter, the y3_set_pdl choice);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as
 *     void*), does no device synchronisation and is CUDA-graph capturable;
 *   - return 0 on success, a Y3_E* code otherwise; y3_last_error() returns a
 *     thread-local message.  There is no CPU fallback: a device that is not
 *     sm_100 is an error.
 *   - activations are NHWC bf16 "views": (pointer to channel 0 of pixel 0,
 *     channel count C, pixel pitch ld in ELEMENTS).  ld > C lets a producer
 *     write straight into a channel slice of a route/concat buffer
 *     (yolov3/darknet.py:369-375 becomes zero-copy).
 */
#ifndef YOLOV3_B200_H_
#define YOLOV3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Y3_ABI_VERSION 6

enum {
  Y3_OK = 0,
  Y3_EINVAL = 1,      /* bad argument / unsupported shape          */
  Y3_ECUDA = 2,       /* CUDA runtime or driver error               */
  Y3_EARCH = 3,       /* device is not sm_100 (no fallback exists)  */
  Y3_EWORKSPACE = 4   /* caller-provided workspace too small        */
};

/* ---- library ---------------------------------------------------------- */
int y3_abi_version(void);
const char* y3_last_error(void);
/* 0 if device `dev` is usable (compute capability 10.x), else Y3_EARCH. */
int y3_check_device(int dev);
/* Number of kernel launches issued by this library on the calling thread
 * since the last y3_reset_launch_count() (bench.py's "gpu_launches"). */
long long y3_launch_count(void);
void y3_reset_launch_count(void);
/* Programmatic dependent launch (every kernel of the library is launched with the
 * programmatic-stream-serialization attribute so that kernel i+1's prologue overlaps kernel i's
 * tail).  on = 0 launches plain kernels instead: the better choice when several independent plans
 * run on different streams at once, because a dependent CTA parked in griddepcontrol.wait holds an
 * SM another plan's kernel could use.  Applies to subsequent launches (a captured CUDA graph keeps
 * what was set at capture time).  Per calling thread (threads that build plans concurrently do
 * not see each other's choice); returns the previous setting.  Default: on (off when the
 * environment has Y3_NO_PDL=1). */
int y3_set_pdl(int on);

/* ---- a12: host-side staging of the image batch ---------------------------------------- */
/* np.stack(images) of yolov3/inference.py:332 as n plain memcpy's into one (pinned) staging buffer,
 * spread over up to `threads` host threads; the caller holds no interpreter lock meanwhile.
 * Host memory only; no CUDA call. */
int y3_stage_images(void* dst, const void* const* srcs, int32_t n, int64_t bytes_each,
                    int32_t threads);

/* ---- a5/a8/a9: convolutional block ------------------------------------- */
/*
 * One Darknet [convolutional] block: Conv2d -> BatchNorm2d(eval) -> LeakyReLU
 * (yolov3/darknet.py:244-257, executed at :367-368) as an implicit GEMM on
 * the 5th-gen tensor cores (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue).
 * BatchNorm is folded into w/bias by the caller at weight-load time.
 * Optional epilogue fusions:
 *   residual != NULL : y = act(conv) + residual   (shortcut, darknet.py:376-379)
 *   upsample2x != 0  : every output pixel is stored to the 2x2 block it
 *                      expands to (nn.Upsample nearest x2, darknet.py:299-305)
 *   out_f32 != 0     : y is float32 (YOLO head logits feeding y3_yolo_decode)
 * Layouts: x NHWC bf16 [N,H,W,Cin] pitch ld_x; w bf16 [Cout_pad][R][S][Cin]
 * (K contiguous, rows beyond Cout zero); bias fp32 [Cout_pad];
 * y NHWC [N,Ho,Wo,*] pitch ld_y (or [N,2Ho,2Wo,*] when upsample2x).
 * Cin must be a multiple of 16; Cout_pad a multiple of 16; pad = (R-1)/2 or 0.
 */
typedef struct y3_conv_desc {
  int32_t n, h, w;          /* input batch / height / width            */
  int32_t cin;              /* input channels (multiple of 16)          */
  int32_t cout;             /* stored output channels (multiple of 16)  */
  int32_t ksize;            /* R = S = 1 or 3                           */
  int32_t stride;           /* 1 or 2                                   */
  int32_t pad;              /* 0 or (ksize-1)/2                         */
  int32_t ld_x, ld_y, ld_res; /* pixel pitches in elements              */
  int32_t leaky;            /* 1: LeakyReLU(0.1); 0: linear (identity)  */
  int32_t out_f32;          /* 1: y is float32; 0: bf16                 */
  int32_t upsample2x;       /* 1: fused nearest x2 store                */
  int32_t flags;            /* validation knobs, normally 0.  bit0: load A
                               through the im2col tensor map even for 1x1/s1;
                               bit1: use the direct (register->global)
                               epilogue instead of the staged TMA-store one;
                               bit2: never use CTA-pair (cta_group::2) tiles;
                               bit3: stream the weights through the operand
                               ring even where the layer's whole slab could
                               stay resident in shared memory */
} y3_conv_desc;

int y3_conv2d(const y3_conv_desc* d, const void* x, const void* w,
              const float* bias, const void* residual, void* y, void* stream);

/* ---- a5 (+a8, a12): two chained convolutional blocks, intermediate kept on chip ---------- */
/*
 * The first layers of Darknet-53 are memory-bound, so two block shapes run as ONE
 * kernel each (conv_chain.cu) and their intermediate activation never reaches HBM:
 *
 *  y3_conv_chain_stem_u8: uint8 BGR HWC images [N,H,W,3] -> RGB /255 (yolov3/inference.py:332-333)
 *    -> conv 3x3 / stride 1 / pad 1, 3 -> 32 (+BN folded, leaky)  -> conv 3x3 / stride 2 / pad 1,
 *    32 -> 64 (+BN folded, leaky): blocks 0 and 1 of models/yolov3.cfg / yolov3-spp.cfg
 *    (built at yolov3/darknet.py:244-257).  w1: bf16 [32][3][16] = per filter row r the 9 taps in
 *    (s, BGR byte) order — the image's own byte order — padded to 16; w2: bf16 [64][3][3][32].  y: NHWC bf16 [N,H/2,W/2,64] pitch ld_y.
 *    H/2 must be a multiple of 16 and W/2 a multiple of 8; img 8-byte aligned.
 *
 *  y3_conv_chain_res64: one residual unit x -> conv 1x1 64 -> 32 -> conv 3x3 / 1 / pad 1 32 -> 64,
 *    + x (the [shortcut] of yolov3/darknet.py:376-379).  x: NHWC bf16 [N,H,W,64] pitch ld_x;
 *    w1: bf16 [32][64]; w2: bf16 [64][3][3][32]; y: NHWC bf16 [N,H,W,64] pitch ld_y, y != x.
 *    H must be a multiple of 16 and W a multiple of 8.
 *
 * The intermediate is rounded to bf16 exactly as the unfused y3_conv2d sequence stores it.
 */
typedef struct y3_chain_desc {
  int32_t n, h, w;          /* input batch / height / width                */
  int32_t ld_x, ld_y;       /* pixel pitches in elements (ld_x: RES only)  */
  int32_t leaky1, leaky2;   /* LeakyReLU(0.1) after the first / second conv */
} y3_chain_desc;

int y3_conv_chain_stem_u8(const y3_chain_desc* d, const uint8_t* img, const void* w1,
                          const float* b1, const void* w2, const float* b2, void* y,
                          void* stream);
int y3_conv_chain_res64(const y3_chain_desc* d, const void* x, const void* w1,
                        const float* b1, const void* w2, const float* b2, void* y,
                        void* stream);

/* ---- a6: max-pool -------------------------------------------------------- */
/*
 * MaxPool2d as the reference patches it (yolov3/darknet.py:16-29): stride
 * `stride`, no symmetric padding; when ksize > 1 and stride == 1 the input is
 * first ZERO-padded on the right/bottom by ksize-1 (so zeros compete with
 * negative activations).  Otherwise floor mode.  NHWC bf16 views.
 */
int y3_maxpool(const void* x, void* y, int32_t n, int32_t h, int32_t w,
               int32_t c, int32_t ld_x, int32_t ld_y, int32_t ksize,
               int32_t stride, void* stream);

/*
 * SPP block of yolov3-spp.cfg (models/yolov3-spp.cfg:575-606): three
 * stride-1 max-pools (k5, k9, k13, zero right/bottom padding as above) of the
 * same input, written to three channel slices in one pass.
 */
int y3_spp3(const void* x, void* y5, void* y9, void* y13, int32_t n, int32_t h,
            int32_t w, int32_t c, int32_t ld_x, int32_t ld_y, void* stream);

/* ---- a7/a8/a9 unfused forms (used when a fusion rule does not apply) ---- */
/* y = a + b (shortcut, yolov3/darknet.py:376-379). */
int y3_add(const void* a, const void* b, void* y, int64_t pixels, int32_t c,
           int32_t ld_a, int32_t ld_b, int32_t ld_y, void* stream);
/* channel-slice copy (route / torch.cat, yolov3/darknet.py:369-375). */
int y3_copy_channels(const void* x, void* y, int64_t pixels, int32_t c,
                     int32_t ld_x, int32_t ld_y, void* stream);
/* nearest-neighbour x2 upsample (yolov3/darknet.py:299-305). */
int y3_upsample2x(const void* x, void* y, int32_t n, int32_t h, int32_t w,
                  int32_t c, int32_t ld_x, int32_t ld_y, void* stream);

/* ---- a12: input packing --------------------------------------------------- */
/* float32 NCHW [N,3,H,W] (Darknet.forward's argument, darknet.py:351) ->
 * NHWC bf16 with channels zero-padded to c_pad. */
int y3_pack_nchw_f32(const float* x, void* y, int32_t n, int32_t c, int32_t h,
                     int32_t w, int32_t c_pad, void* stream);
/* uint8 BGR HWC images [N,H,W,3] -> RGB /255 -> NHWC bf16 padded to c_pad
 * (the flip / transpose / astype / 255 of yolov3/inference.py:332-333). */
int y3_pack_bgr_u8(const uint8_t* x, void* y, int32_t n, int32_t h, int32_t w,
                   int32_t c_pad, void* stream);

/* Same two conversions fused with the im2col of a FIRST layer that is a 3x3 /
 * stride 1 / pad 1 convolution over c <= 3 channels: y is [N*H*W, k_pad] bf16,
 * row = the 9*c taps in (r, s, c) order (the weight layout) then zeros.  The
 * first convolution then runs as y3_conv2d with ksize 1, cin = k_pad. */
int y3_im2col3x3_nchw_f32(const float* x, void* y, int32_t n, int32_t c,
                          int32_t h, int32_t w, int32_t k_pad, void* stream);
int y3_im2col3x3_bgr_u8(const uint8_t* x, void* y, int32_t n, int32_t h,
                        int32_t w, int32_t k_pad, void* stream);

/* ---- a10/a11/a13/a14: YOLO decode ----------------------------------------- */
/*
 * One YOLO head.  logits: float32 NHWC [N,g_h,g_w,ld] with channel
 * a*(5+classes)+f.  anchors: this head's masked anchors (w,h) in pixels.
 * Box m = a*g_h*g_w + row*g_w + col, placed at box_offset+m of each image
 * (yolov3/darknet.py:48-122, :390-399).
 */
typedef struct y3_head_desc {
  int32_t n, g_h, g_w;
  int32_t num_anchors;      /* <= 8                                     */
  int32_t num_classes;      /* <= 1024                                  */
  int32_t ld;               /* pixel pitch of logits in floats          */
  int32_t box_offset;       /* first box index of this head per image   */
  int32_t boxes_per_image;  /* M: total boxes over all heads            */
  float anchor_w[8], anchor_h[8];
  float train_w, train_h;   /* net_info width / height (darknet.py:395) */
} y3_head_desc;

/* Dense outputs exactly as Darknet.forward returns them:
 * bbox_xywh float32 [N,M,4], class_prob float32 [N,M], class_idx int64 [N,M]. */
int y3_yolo_decode_dense(const y3_head_desc* d, const float* logits,
                         float* bbox_xywh, float* class_prob,
                         int64_t* class_idx, void* stream);

/*
 * Candidate record produced by the fused decode and consumed by y3_nms:
 * the per-image post-processing of yolov3/inference.py:342-353 and
 * cxywh_to_tlbr (:269-283) applied on device — threshold prob >= prob_thresh,
 * scale by the original image size in fp32, truncate to integer, tl/br.
 */
typedef struct y3_cand {
  int32_t x1, y1, x2, y2;   /* pixels of the ORIGINAL image, unclipped   */
  float prob;
  int32_t cls;
  int32_t box;              /* box index m within the image              */
  int32_t pad_;
} y3_cand;

/*
 * The two thresholds of `inference` (prob_thresh, nms_iou_thresh;
 * yolov3/inference.py:287) as a DEVICE-resident record.  Entry points that
 * take thresholds by value also take `const y3_thresholds* dev_thresholds`
 * (nullable): when non-NULL the kernels read the thresholds from it at RUN
 * time and ignore the by-value arguments, so one captured CUDA graph serves
 * every threshold setting (the caller updates the 16-byte record with a
 * stream-ordered copy before replaying the graph).
 */
typedef struct y3_thresholds {
  float prob_thresh;        /* keep prob >= prob_thresh                   */
  float reserved_;
  double iou_thresh;        /* suppress iou > iou_thresh                  */
} y3_thresholds;

/* Fused decode + threshold + compaction.  orig_hw: int32 [N,2] (H,W) on the
 * device.  cands: [N,cap]; counts: int32 [N] (caller zeroes before the first
 * head; heads accumulate).  Candidates beyond cap are dropped and counted. */
int y3_yolo_decode_cands(const y3_head_desc* d, const float* logits,
                         float prob_thresh, const y3_thresholds* dev_thresholds,
                         const int32_t* orig_hw, y3_cand* cands,
                         int32_t* counts, int32_t cap, void* stream);

/*
 * YOLO head convolution with the decode fused into its epilogue: the 1x1 convolution feeding a
 * [yolo] block (yolov3/darknet.py:244-261) followed by y3_yolo_decode_cands' arithmetic, applied to
 * the fp32 accumulator (+ bias) of each pixel while it is still in tensor memory — the logits never
 * go to HBM.  Requires 3 anchors x 80 classes (255 channels stored as 256; every shipped cfg);
 * otherwise use y3_conv2d (out_f32) + y3_yolo_decode_cands.  The softmax denominator is summed in
 * ascending class order with ex2.approx-based exponentials (terms near the maximum carry <= 4 ulp), so
 * a probability may differ from y3_yolo_decode_cands' by up to ~3e-7 relative.
 */
int y3_conv2d_yolo_head(const y3_conv_desc* d, const void* x, const void* w,
                        const float* bias, const y3_head_desc* head,
                        float prob_thresh, const y3_thresholds* dev_thresholds,
                        const int32_t* orig_hw, y3_cand* cands,
                        int32_t* counts, int32_t cap, void* stream);

/* ---- a15/a16: non-max suppression ------------------------------------------ */
/*
 * Greedy IoU suppression, bit-exact with _non_max_suppression
 * (yolov3/inference.py:161-217): area (x2-x1+1)(y2-y1+1) in int64,
 * iou = inter/union as an IEEE float64 divide, suppress iff iou > iou_thresh,
 * visiting boxes by descending prob (ties: ascending box index; the reference
 * leaves tie order unspecified).  per_class != 0 runs it independently per
 * class (non_max_suppression, :220-266); classes must lie in [0,num_classes).
 *
 * cands [N,cap] / counts [N] as written by y3_yolo_decode_cands (or by the
 * caller; counts above cap are clamped).  Outputs: `sorted` [N,cap] = each
 * image's candidates ordered by (class asc [per_class only], prob desc, box
 * asc); keep [N,cap] uint8 flags aligned with `sorted`; class_first_box
 * (nullable) int32 [N,num_classes] (or [N,1] when !per_class) = smallest box
 * index among the candidates of each class, INT32_MAX if none — what the
 * host needs to replay the reference's set(class_idx) visiting order.
 * class_start (nullable) int32 [N,C+1] (C = num_classes, or 1 when
 * !per_class): first index of each class segment inside the image's `sorted`
 * row, last entry = number of candidates; class_kept (nullable) int32 [N,C]:
 * records kept per class segment.
 * workspace: at least y3_nms_workspace_bytes(N, cap, num_classes) bytes.
 */
size_t y3_nms_workspace_bytes(int32_t n, int32_t cap, int32_t num_classes);
int y3_nms(const y3_cand* cands, const int32_t* counts, int32_t n, int32_t cap,
           int32_t num_classes, double iou_thresh,
           const y3_thresholds* dev_thresholds, int32_t per_class,
           y3_cand* sorted, uint8_t* keep, int32_t* class_first_box,
           int32_t* class_start, int32_t* class_kept,
           void* workspace, size_t workspace_bytes, void* stream);

/* Destinations for y3_emit_detections with class groups in ASCENDING class
 * order (what set(class_idx) yields whenever an image holds >= 19 distinct
 * classes < 128; the host re-orders the few images where it does not):
 * exclusive scan of class_kept [N,num_segments] in (image, segment) order ->
 * dst_off [N,num_segments]; det_counts [N] = detections kept per image;
 * det_counts[N] (one extra entry) = total.  n*num_segments <= 2^20. */
int y3_plan_destinations(const int32_t* class_kept, int32_t n, int32_t num_segments,
                         int32_t* dst_off, int32_t* det_counts, void* stream);

/* a17: the arrays `inference` returns (yolov3/inference.py:360-366), in their
 * final dtypes and order, built on the device.  For every (image, segment)
 * pair — segments as delimited by class_start [N,num_segments+1] from y3_nms
 * — the kept records (prob descending) are written from position
 * dst_off[image*num_segments + segment] of the flat outputs (negative: skip):
 * tlbr int64 [K,4], prob float32 [K], cls int64 [K].  The host chooses dst_off
 * so that class groups follow the reference's set(class_idx) visiting order. */
int y3_emit_detections(const y3_cand* sorted, const uint8_t* keep,
                       const int32_t* class_start, const int32_t* dst_off,
                       int32_t n, int32_t cap, int32_t num_segments,
                       int64_t* tlbr, float* prob, int64_t* cls, void* stream);

/* Ordered compaction of the kept records.  det_counts [N] receives the number
 * kept per image.  flat == 0: image i's records start at dets[i*cap];
 * flat != 0: records of all images are packed back to back (image i starts at
 * the sum of det_counts[0..i)), so one D2H copy moves exactly the detections. */
int y3_compact_kept(const y3_cand* sorted, const uint8_t* keep,
                    const int32_t* counts, int32_t n, int32_t cap,
                    y3_cand* dets, int32_t* det_counts, int32_t flat,
                    void* stream);

/* ---- diagnostics ------------------------------------------------------------ */
/* With Y3_CONV_TRACE=1 in the environment, CTA 0 of every convolution launch
 * records clock64() stamps of its pipeline roles; this copies the first n
 * (<= 96) words of the last traced launch to the host (synchronises the
 * device; tools/conv_trace.py).  Not used by the hot path. */
int y3_debug_conv_trace(unsigned long long* out, int n);
/* Every blocking mbarrier wait of the convolution kernels has a watchdog (~8 s) that traps instead of
 * hanging the GPU.  Register a zeroed HOST-MAPPED (pinned, device-accessible) buffer of 12 x 8 bytes on
 * the current device and every thread that times out first leaves there: [0] magic 0x59335452415021,
 * [5] number of timed-out threads, and for the LAST of them [1] source file id << 32 | line of the wait,
 * [2] threadIdx.x << 32 | blockIdx.x, [3] parity << 32 | barrier address, [4] blockDim.x << 32 |
 * gridDim.x; [8..11] the same four words for the FIRST one.  NULL unregisters.  The record survives the
 * dead context. */
int y3_debug_set_trap_record(void* host_mapped);

#ifdef __cplusplus
}
#endif
#endif /* YOLOV3_B200_H_ */
