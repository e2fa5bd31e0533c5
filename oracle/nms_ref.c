/* CPU restatement of greedy NMS as used by the reference post-processing.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference calls torchvision.ops.nms (os2d/structures/bounding_box.py:4, :367), a third-party dependency
 * that is not vendored under /root/reference (pinned torchvision 0.14.0 in Docker/requirements.txt:249-252,
 * 0.26.0 installed in this image).  This file restates its published CPU algorithm:
 *   order = stable argsort of scores, descending; areas = (x2-x1)*(y2-y1);
 *   visit boxes in that order, keep a box if not suppressed, then suppress every later box whose
 *   IoU = inter / (area_i + area_j - inter) is > threshold (float IoU compared against a double threshold).
 * tests/test_oracle_vs_reference.py pins it against the installed torchvision.ops.nms on random and tie-heavy
 * inputs; tests/golden/ holds the committed vectors.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>

/* boxes [n][4] xyxy, order [n] visiting order (indices into boxes); keep_out [n] receives the kept indices
 * (in visiting order); returns the number kept. */
int64_t os2d_oracle_nms(const float* boxes, const int64_t* order, int64_t n, double iou_threshold, int64_t* keep_out) {
  unsigned char* suppressed = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  int64_t num_keep = 0;
  for (int64_t _i = 0; _i < n; ++_i) {
    const int64_t i = order[_i];
    if (suppressed[i]) continue;
    keep_out[num_keep++] = i;
    const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    const float iarea = (ix2 - ix1) * (iy2 - iy1);
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      const int64_t j = order[_j];
      if (suppressed[j]) continue;
      const float jx1 = boxes[4 * j], jy1 = boxes[4 * j + 1], jx2 = boxes[4 * j + 2], jy2 = boxes[4 * j + 3];
      const float xx1 = ix1 > jx1 ? ix1 : jx1, yy1 = iy1 > jy1 ? iy1 : jy1;
      const float xx2 = ix2 < jx2 ? ix2 : jx2, yy2 = iy2 < jy2 ? iy2 : jy2;
      float w = xx2 - xx1, h = yy2 - yy1;
      w = w > 0.0f ? w : 0.0f;
      h = h > 0.0f ? h : 0.0f;
      const float inter = w * h;
      const float jarea = (jx2 - jx1) * (jy2 - jy1);
      const float ovr = inter / (iarea + jarea - inter);
      if ((double)ovr > iou_threshold) suppressed[j] = 1;
    }
  }
  free(suppressed);
  return num_keep;
}
