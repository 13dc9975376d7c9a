/*
 * CPU restatement of tf.image.non_max_suppression as TensorFlow 1.13 runs it
 * (NonMaxSuppressionV3 with score_threshold = -inf) -- TEST ORACLE, not product.
 *
 * The algorithm lives in the un-vendored TensorFlow 1.13 binary
 * (tensorflow/core/kernels/non_max_suppression_op.cc); it is restated here from
 * its published behaviour, anchored on the reference's
 * call sites SSD300.py:179, RetinaNet.py:246, YOLOv3.py:358, FCOS.py:255:
 *
 *   - all boxes are candidates, visited in descending score order
 *     (a max-heap keyed on score; ties -> lower index first, our defined rule:
 *      TF 1.13's tie order is whatever libstdc++'s heap yields, so the harness
 *      keeps scores tie-free);
 *   - a candidate is selected unless IoU(candidate, s) > iou_threshold (strict)
 *     for some already selected s, checked most-recent first;
 *   - stops after max_output_size selections;
 *   - IoU in float32: corners normalised with min/max, area <= 0 -> IoU 0.
 *
 * boxes: [n,4] (y1,x1,y2,x2); returns the number selected, indices into the
 * input arrays in selection order.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC nms_ref.c -o libodt_oracle.so
 */
#include <stdlib.h>

static float iou(const float* b, int i, int j) {
  const float* a = b + 4 * i;
  const float* c = b + 4 * j;
  float ymin_i = a[0] < a[2] ? a[0] : a[2], xmin_i = a[1] < a[3] ? a[1] : a[3];
  float ymax_i = a[0] > a[2] ? a[0] : a[2], xmax_i = a[1] > a[3] ? a[1] : a[3];
  float ymin_j = c[0] < c[2] ? c[0] : c[2], xmin_j = c[1] < c[3] ? c[1] : c[3];
  float ymax_j = c[0] > c[2] ? c[0] : c[2], xmax_j = c[1] > c[3] ? c[1] : c[3];
  float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
  float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
  if (area_i <= 0.0f || area_j <= 0.0f) return 0.0f;
  float iy1 = ymin_i > ymin_j ? ymin_i : ymin_j, ix1 = xmin_i > xmin_j ? xmin_i : xmin_j;
  float iy2 = ymax_i < ymax_j ? ymax_i : ymax_j, ix2 = xmax_i < xmax_j ? xmax_i : xmax_j;
  float h = iy2 - iy1, w = ix2 - ix1;
  if (h < 0.0f) h = 0.0f;
  if (w < 0.0f) w = 0.0f;
  float inter = h * w;
  return inter / (area_i + area_j - inter);
}

typedef struct {
  float score;
  int index;
} cand_t;

/* heap order: higher score first; equal scores -> lower index first */
static int before(const cand_t* a, const cand_t* b) {
  if (a->score != b->score) return a->score > b->score;
  return a->index < b->index;
}

static void sift_down(cand_t* h, int n, int i) {
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    if (l < n && before(&h[l], &h[m])) m = l;
    if (r < n && before(&h[r], &h[m])) m = r;
    if (m == i) return;
    cand_t t = h[i];
    h[i] = h[m];
    h[m] = t;
    i = m;
  }
}

int odt_oracle_nms(const float* boxes, const float* scores, int n, int max_output_size,
                   float iou_threshold, int* selected) {
  if (n <= 0 || max_output_size <= 0) return 0;
  cand_t* heap = (cand_t*)malloc(sizeof(cand_t) * (size_t)n);
  for (int i = 0; i < n; ++i) {
    heap[i].score = scores[i];
    heap[i].index = i;
  }
  for (int i = n / 2 - 1; i >= 0; --i) sift_down(heap, n, i);
  int hn = n, nsel = 0;
  while (nsel < max_output_size && hn > 0) {
    cand_t c = heap[0];
    heap[0] = heap[--hn];
    sift_down(heap, hn, 0);
    int keep = 1;
    for (int j = nsel - 1; j >= 0; --j) {
      if (iou(boxes, c.index, selected[j]) > iou_threshold) {
        keep = 0;
        break;
      }
    }
    if (keep) selected[nsel++] = c.index;
  }
  free(heap);
  return nsel;
}
