/* TEST INFRASTRUCTURE - CPU oracle of the rotated-box overlap / IoU / NMS path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product never does.
 *
 * Plain-C restatement of the reference's algorithm, each function citing the lines it follows
 * (paths relative to /root/reference/pcdet/ops/iou3d_nms/src):
 *   overlap of two rotated rectangles   iou3d_cpu.cpp:86-210  (= iou3d_nms_kernel.cu:104-236)
 *   iou_bev                             iou3d_cpu.cpp:212-220
 *   boxes_iou_bev_cpu                   iou3d_cpu.cpp:222-252
 *   nms mask + sequential sweep         iou3d_nms_kernel.cu:267-313, iou3d_nms.cpp:88-131
 *   axis-aligned iou / nms_normal       iou3d_nms_kernel.cu:316-366, iou3d_nms.cpp:134-187
 * Pinned: tests/golden/iou3d_kat.npz is written by the reference's own iou3d_cpu.cpp compiled here from where it lies
 * (oracle/build_oracle.py -> oracle/_ref/), and tests/test_oracle_golden.py holds this file to it bit for bit.
 * Boxes are 7 floats [x, y, z, dx, dy, dz, heading].  All arithmetic is float, as in the reference. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EPS 1e-8f

static float cross_o(const float* p1, const float* p2, const float* p0) {
  return (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]);
}

static float fmin2(float a, float b) { return a > b ? b : a; }
static float fmax2(float a, float b) { return a > b ? a : b; }

/* iou3d_cpu.cpp:52-58 fast exclusion + :64-98 intersection of two segments */
static int seg_cross(const float* p1, const float* p0, const float* q1, const float* q0, float* ans) {
  int boxes_meet = fmin2(p0[0], p1[0]) <= fmax2(q0[0], q1[0]) && fmin2(q0[0], q1[0]) <= fmax2(p0[0], p1[0]) &&
                   fmin2(p0[1], p1[1]) <= fmax2(q0[1], q1[1]) && fmin2(q0[1], q1[1]) <= fmax2(p0[1], p1[1]);
  if (!boxes_meet) return 0;
  float s1 = cross_o(q0, p1, p0), s2 = cross_o(p1, q1, p0), s3 = cross_o(p0, q1, q0), s4 = cross_o(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  float s5 = cross_o(q1, p1, p0);
  if (fabsf(s5 - s1) > EPS) {
    ans[0] = (s5 * q0[0] - s1 * q1[0]) / (s5 - s1);
    ans[1] = (s5 * q0[1] - s1 * q1[1]) / (s5 - s1);
  } else {
    float a0 = p0[1] - p1[1], b0 = p1[0] - p0[0], c0 = p0[0] * p1[1] - p1[0] * p0[1];
    float a1 = q0[1] - q1[1], b1 = q1[0] - q0[0], c1 = q0[0] * q1[1] - q1[0] * q0[1];
    float D = a0 * b1 - a1 * b0;
    ans[0] = (b0 * c1 - b1 * c0) / D;
    ans[1] = (a1 * c0 - a0 * c1) / D;
  }
  return 1;
}

/* iou3d_cpu.cpp:60-72: point inside the rotated box, margin 1e-2 */
static int inside(const float* box, const float* p) {
  const float MARGIN = 1e-2f;
  float c = cosf(-box[6]), s = sinf(-box[6]);
  float rx = (p[0] - box[0]) * c + (p[1] - box[1]) * (-s);
  float ry = (p[0] - box[0]) * s + (p[1] - box[1]) * c;
  return fabsf(rx) < box[3] / 2 + MARGIN && fabsf(ry) < box[4] / 2 + MARGIN;
}

/* iou3d_cpu.cpp:112-140: axis-aligned corners rotated around the centre; corner 4 repeats corner 0 */
static void corners(const float* box, float c[5][2]) {
  float hx = box[3] / 2, hy = box[4] / 2;
  float x1 = box[0] - hx, y1 = box[1] - hy, x2 = box[0] + hx, y2 = box[1] + hy;
  float raw[4][2] = {{x1, y1}, {x2, y1}, {x2, y2}, {x1, y2}};
  float ca = cosf(box[6]), sa = sinf(box[6]);
  for (int k = 0; k < 4; ++k) {
    c[k][0] = (raw[k][0] - box[0]) * ca + (raw[k][1] - box[1]) * (-sa) + box[0];
    c[k][1] = (raw[k][0] - box[0]) * sa + (raw[k][1] - box[1]) * ca + box[1];
  }
  c[4][0] = c[0][0];
  c[4][1] = c[0][1];
}

float oracle_box_overlap(const float* a, const float* b) {
  float ca[5][2], cb[5][2], pts[16][2], centre[2] = {0.f, 0.f};
  int cnt = 0;
  corners(a, ca);
  corners(b, cb);
  for (int i = 0; i < 4; ++i)          /* :146-160 */
    for (int j = 0; j < 4; ++j)
      if (seg_cross(ca[i + 1], ca[i], cb[j + 1], cb[j], pts[cnt])) {
        centre[0] = centre[0] + pts[cnt][0];
        centre[1] = centre[1] + pts[cnt][1];
        cnt++;
      }
  for (int k = 0; k < 4; ++k) {        /* :163-180 */
    if (inside(a, cb[k])) {
      centre[0] = centre[0] + cb[k][0];
      centre[1] = centre[1] + cb[k][1];
      pts[cnt][0] = cb[k][0];
      pts[cnt][1] = cb[k][1];
      cnt++;
    }
    if (inside(b, ca[k])) {
      centre[0] = centre[0] + ca[k][0];
      centre[1] = centre[1] + ca[k][1];
      pts[cnt][0] = ca[k][0];
      pts[cnt][1] = ca[k][1];
      cnt++;
    }
  }
  centre[0] /= cnt;
  centre[1] /= cnt;
  for (int j = 0; j < cnt - 1; ++j)    /* :186-196 bubble sort by angle */
    for (int i = 0; i < cnt - j - 1; ++i)
      if (atan2f(pts[i][1] - centre[1], pts[i][0] - centre[0]) > atan2f(pts[i + 1][1] - centre[1], pts[i + 1][0] - centre[0])) {
        float tx = pts[i][0], ty = pts[i][1];
        pts[i][0] = pts[i + 1][0];
        pts[i][1] = pts[i + 1][1];
        pts[i + 1][0] = tx;
        pts[i + 1][1] = ty;
      }
  float area = 0;
  for (int k = 0; k < cnt - 1; ++k) {  /* :199-203 triangle fan */
    float ux = pts[k][0] - pts[0][0], uy = pts[k][1] - pts[0][1], vx = pts[k + 1][0] - pts[0][0], vy = pts[k + 1][1] - pts[0][1];
    area += ux * vy - uy * vx;
  }
  return (float)(fabs(area) / 2.0);
}

float oracle_iou_bev(const float* a, const float* b) {
  float sa = a[3] * a[4], sb = b[3] * b[4];
  float so = oracle_box_overlap(a, b);
  return so / fmaxf(sa + sb - so, EPS);
}

static float iou_normal(const float* a, const float* b) {
  float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
  float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
  float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
  float inter = w * h;
  return inter / fmaxf(a[3] * a[4] + b[3] * b[4] - inter, EPS);
}

void oracle_boxes_overlap_bev(int na, const float* a, int nb, const float* b, float* out) {
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) out[(long)i * nb + j] = oracle_box_overlap(a + 7 * i, b + 7 * j);
}

void oracle_boxes_iou_bev(int na, const float* a, int nb, const float* b, float* out) {
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) out[(long)i * nb + j] = oracle_iou_bev(a + 7 * i, b + 7 * j);
}

/* boxes sorted by descending score; keep[] receives the kept positions in order; returns their number.
 * Box i suppresses every later box j > i with IoU(i, j) > thresh unless i itself was suppressed - the outcome of the
 * reference's mask words + host sweep (iou3d_nms_kernel.cu:267-313, iou3d_nms.cpp:113-128). */
int oracle_nms(const float* boxes, int n, float thresh, int rotated, long long* keep) {
  unsigned char* removed = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  int num = 0;
  for (int i = 0; i < n; ++i) {
    if (removed[i]) continue;
    keep[num++] = i;
    for (int j = i + 1; j < n; ++j) {
      float v = rotated ? oracle_iou_bev(boxes + 7 * i, boxes + 7 * j) : iou_normal(boxes + 7 * i, boxes + 7 * j);
      if (v > thresh) removed[j] = 1;
    }
  }
  free(removed);
  return num;
}
