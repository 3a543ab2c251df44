/*
 * nms_oracle.c — CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * Plain-C restatement of the greedy NMS of nrsyed/pytorch-yolov3
 * (yolov3/inference.py:161-217 `_non_max_suppression`, :220-266 per-class driver), fast enough
 * for the 256-image x 10,647-candidate stress configuration.  Same arithmetic as NumPy
 * evaluates there: int64 "+1" areas, iou = (double)inter / (double)union, drop iff iou > thresh.
 * Visiting order inside a class: descending prob, ties by ascending index (the reference's
 * np.argsort tie order is unspecified; test vectors are tie-free and assert it).
 *
 * Built by oracle/Makefile into oracle/_build/libnms_oracle.so and checked against the NumPy
 * restatement (oracle/postprocess_oracle.py), itself pinned to the live reference by
 * tests/golden/make_golden.py.
 */
#include <stdint.h>
#include <stdlib.h>

typedef struct { float prob; int32_t idx; } item_t;

static int cmp_desc(const void* a, const void* b) {
  const item_t* x = (const item_t*)a;
  const item_t* y = (const item_t*)b;
  if (x->prob > y->prob) return -1;
  if (x->prob < y->prob) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* tlbr: int64 [n,4]; prob: float32 [n]; members: indices of the boxes taking part (m of them);
 * keep_out: receives kept ORIGINAL indices in visiting order; returns their number. */
static int64_t greedy(const int64_t* tlbr, const float* prob, const int32_t* members, int64_t m,
                      double thresh, int32_t* keep_out, item_t* order, uint8_t* dead) {
  for (int64_t i = 0; i < m; ++i) { order[i].prob = prob[members[i]]; order[i].idx = members[i]; dead[i] = 0; }
  qsort(order, (size_t)m, sizeof(item_t), cmp_desc);
  int64_t kept = 0;
  for (int64_t i = 0; i < m; ++i) {
    if (dead[i]) continue;
    const int64_t* a = tlbr + 4 * (int64_t)order[i].idx;
    keep_out[kept++] = order[i].idx;
    const int64_t area_a = ((a[2] - a[0]) + 1) * ((a[3] - a[1]) + 1);
    for (int64_t j = i + 1; j < m; ++j) {
      if (dead[j]) continue;
      const int64_t* b = tlbr + 4 * (int64_t)order[j].idx;
      int64_t iw = ((a[2] < b[2] ? a[2] : b[2]) - (a[0] > b[0] ? a[0] : b[0])) + 1;
      int64_t ih = ((a[3] < b[3] ? a[3] : b[3]) - (a[1] > b[1] ? a[1] : b[1])) + 1;
      if (iw < 0) iw = 0;
      if (ih < 0) ih = 0;
      const int64_t inter = iw * ih;
      const int64_t area_b = ((b[2] - b[0]) + 1) * ((b[3] - b[1]) + 1);
      const int64_t uni = area_a + area_b - inter;
      const double iou = (double)inter / (double)uni;
      if (iou > thresh) dead[j] = 1;
    }
  }
  return kept;
}

/* Per-class NMS (class_idx != NULL) visiting classes in the order given by class_order
 * (n_order entries — the caller supplies Python's set() order), or class-agnostic NMS
 * (class_idx == NULL).  keep_out must hold n entries.  Returns the number kept, -1 on OOM. */
int64_t nms_oracle(const int64_t* tlbr, const float* prob, const int64_t* class_idx, int64_t n,
                   const int64_t* class_order, int64_t n_order, double thresh, int32_t* keep_out) {
  item_t* order = (item_t*)malloc(sizeof(item_t) * (size_t)(n > 0 ? n : 1));
  uint8_t* dead = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
  int32_t* members = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  if (!order || !dead || !members) { free(order); free(dead); free(members); return -1; }
  int64_t total = 0;
  if (!class_idx) {
    for (int64_t i = 0; i < n; ++i) members[i] = (int32_t)i;
    total = greedy(tlbr, prob, members, n, thresh, keep_out, order, dead);
  } else {
    for (int64_t c = 0; c < n_order; ++c) {
      int64_t m = 0;
      for (int64_t i = 0; i < n; ++i)
        if (class_idx[i] == class_order[c]) members[m++] = (int32_t)i;
      total += greedy(tlbr, prob, members, m, thresh, keep_out + total, order, dead);
    }
  }
  free(order); free(dead); free(members);
  return total;
}
