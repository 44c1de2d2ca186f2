// YOLO_LOSS.build_targets on the GPU (reference loss.py:101-192): the YOLOv3-style matcher that train.py uses when
// --ultralytics_loss is absent (train.py:102-106).  The loss itself (loss.py:195-246) has the same form as ComputeLoss
// (GIoU box term, dense objectness BCE with tobj = GIoU.clamp(0), class BCE, per-level balance) and runs on the kernels of
// loss.cu; this file only produces their row lists.
//
// Reference semantics, per image and per box in label order (sequential: a box sees the cells taken by earlier boxes):
//   iou_width_height(box wh, anchors)  (utils/bboxes_utils.py:6-29) -- float32 by torch's type promotion although the labels
//   are np.loadtxt float64 -- anchors visited in descending IoU (stable: ties -> lower index, torch CPU argsort);
//   scale = a // 3, anchor_on_scale = a % 3, cell (i, j) = (int(H*y), int(W*x));
//   free cell and no anchor yet on this scale -> object: [x*W - j, y*H - i, w*W, h*H], class, objectness 1;
//   free cell and IoU > 0.5                   -> objectness -1 ("ignore"; it still enters the BCE as a target of -1).
// The reference divides its anchor tensor by 640 IN PLACE on every iou_width_height call (bboxes_utils.py:18), so the anchors
// the k-th call ever made sees are anchors0 / 640^k: `anchor_table` holds those (host-computed with the same fp32 divisions;
// all zero from k = 17 on) and `decay_base` is the number of calls made before this batch.  Passing T = 2 and
// decay_base = 0 with stride 0 between boxes (fix mode, see yolo_loss.py) gives every box the correctly normalised anchors.
//
// One thread per image (boxes of an image are inherently sequential; nine anchors per box; the work is microseconds).
#include "../../include/yolov5m_b200.h"
#include "common.cuh"

#include <algorithm>
#include <cstring>

namespace yb {

struct YBTParams {
  yb_loss_level lv[3];
  const double* labels;   // [nt][5] class, x, y, w, h (normalised)
  const int* offsets;     // [B+1] first label row of each image
  const float* table;     // [T][9][2] anchors after k in-place divisions by 640, times the level stride
  const float* head_anchors;  // [3][3][2] stride-divided anchors handed to the loss rows (loss.py:41 anchors_d)
  int8_t* state;          // per level (B,3,H,W): 0 free, 1 object, -1 ignore
  float* dense;           // optional: per level (B,3,H,W,6) target tensors like the reference returns
  int* counts;            // [3] row slots to scan (3 per label row)
  int* nobj;              // [3] object rows (denominator of the box / class means)
  long cap;
  long decay_base;
  int B, nt, T, decay_stride;
  float ignore_thr;
};

__global__ void yolo_build_targets_kernel(const __grid_constant__ YBTParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  if (b == 0)
    for (int l = 0; l < 3; ++l) P.counts[l] = 3 * P.nt;  // row slots the loss kernels scan (unused ones hold image -1)
  const int r0 = P.offsets[b], r1 = P.offsets[b + 1];
  long state_off[3], dense_off[3];
  {
    long so = 0, dn = 0;
    for (int l = 0; l < 3; ++l) {
      state_off[l] = so;
      dense_off[l] = dn;
      so += (long)P.B * 3 * P.lv[l].H * P.lv[l].W;
      dn += (long)P.B * 3 * P.lv[l].H * P.lv[l].W * 6;
    }
  }
  int cursor[3] = {0, 0, 0}, objs[3] = {0, 0, 0};
  for (int r = r0; r < r1; ++r) {
    const double* lb = P.labels + (long)r * 5;
    const double cls = lb[0], x = lb[1], y = lb[2], w = lb[3], h = lb[4];
    long e = P.decay_base + (long)(r + 1) * P.decay_stride + (P.decay_stride == 0 ? 1 : 0);
    if (e > P.T - 1) e = P.T - 1;
    const float* an = P.table + e * 18;
    // torch type promotion (bboxes_utils.py:22-29): the float64 box values are 0-dim tensors and the anchors a dimensioned
    // float32 tensor, so every mixed operation runs in float32; only w*h (0-dim x 0-dim) is a float64 product
    float iou[9];
    const float wf = (float)w, hf = (float)h, whf = (float)__dmul_rn(w, h);
    for (int a = 0; a < 9; ++a) {
      const float aw = an[2 * a], ah = an[2 * a + 1];
      const float inter = __fmul_rn(fminf(wf, aw), fminf(hf, ah));
      const float uni = __fsub_rn(__fadd_rn(whf, __fmul_rn(aw, ah)), inter);
      iou[a] = __fdiv_rn(inter, uni);
    }
    int order[9];
    for (int a = 0; a < 9; ++a) order[a] = a;
    for (int i = 1; i < 9; ++i) {  // stable insertion sort, descending; NaN sorts first (torch treats NaN as the largest)
      const int o = order[i];
      const float v = iou[o];
      int j = i - 1;
      while (j >= 0) {
        const float u = iou[order[j]];
        const bool u_ge_v = (u != u) || (!(v != v) && u >= v);
        if (u_ge_v) break;
        order[j + 1] = order[j];
        --j;
      }
      order[j + 1] = o;
    }
    bool has[3] = {false, false, false};
    for (int k = 0; k < 9; ++k) {
      const int a = order[k], l = a / 3, aos = a - 3 * l;
      const int H = P.lv[l].H, W = P.lv[l].W;
      const double fy = __dmul_rn((double)H, y), fx = __dmul_rn((double)W, x);
      const long i = (long)fy, j = (long)fx;  // int(): truncation toward zero
      if (i < 0 || i >= H || j < 0 || j >= W) continue;  // the reference raises IndexError here; never write out of bounds
      const long cell = (((long)b * 3 + aos) * H + i) * W + j;
      int8_t* st = P.state + state_off[l] + cell;
      if (*st != 0) continue;  // `not anchor_taken` is false for 1 and for -1
      const bool obj = !has[l];
      if (!obj && !(iou[a] > P.ignore_thr)) continue;
      *st = obj ? 1 : -1;
      const yb_loss_level& L = P.lv[l];
      const long slot = 3L * r0 + cursor[l]++;
      if (slot < P.cap) {
        L.idx[0 * P.cap + slot] = b;
        L.idx[1 * P.cap + slot] = aos;
        L.idx[2 * P.cap + slot] = i;
        L.idx[3 * P.cap + slot] = j;
        float4 tb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (obj) {
          tb.x = (float)__dsub_rn(fx, (double)j);
          tb.y = (float)__dsub_rn(fy, (double)i);
          tb.z = (float)__dmul_rn(w, (double)W);
          tb.w = (float)__dmul_rn(h, (double)H);
        }
        reinterpret_cast<float4*>(L.tbox)[slot] = tb;
        L.anch[2 * slot + 0] = P.head_anchors[(l * 3 + aos) * 2 + 0];
        L.anch[2 * slot + 1] = P.head_anchors[(l * 3 + aos) * 2 + 1];
        L.tcls[slot] = obj ? (long)cls : -1;  // int(classes[idx]); -1 marks an "ignore" row for loss_rows_kernel
      }
      if (P.dense != nullptr) {
        float* d = P.dense + dense_off[l] + cell * 6;
        if (obj) {
          d[0] = (float)__dsub_rn(fx, (double)j);
          d[1] = (float)__dsub_rn(fy, (double)i);
          d[2] = (float)__dmul_rn(w, (double)W);
          d[3] = (float)__dmul_rn(h, (double)H);
          d[4] = 1.f;
          d[5] = (float)(long)cls;
        } else {
          d[4] = -1.f;
        }
      }
      if (obj) {
        has[l] = true;
        ++objs[l];
      }
    }
  }
  for (int l = 0; l < 3; ++l)
    if (objs[l]) atomicAdd(&P.nobj[l], objs[l]);
}

}  // namespace yb

using namespace yb;

extern "C" int yb_yolo_build_targets(const double* labels, const int* offsets, int B, int nt, const float* anchor_table, int T,
                                     int64_t decay_base, int decay_stride, const float* head_anchors,
                                     const yb_loss_level* levels, int nl, int na, float ignore_thr, int64_t cap,
                                     int8_t* state, float* dense, int* counts, int* nobj, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  YB_REQUIRE(nl == 3 && na == 3, "yolo_build_targets: the reference matcher is written for 3 levels x 3 anchors (got %d x %d)", nl, na);
  YB_REQUIRE(cap >= 3L * nt, "yolo_build_targets: cap=%lld < 3*nt", (long long)cap);
  YB_REQUIRE(T >= 2, "yolo_build_targets: anchor table needs >= 2 rows");
  YBTParams P;
  memset(&P, 0, sizeof(P));
  long cells = 0;
  for (int i = 0; i < 3; ++i) {
    P.lv[i] = levels[i];
    cells += (long)B * 3 * levels[i].H * levels[i].W;
    // unused row slots carry image index -1 (loss_rows_kernel skips them)
    if (nt > 0) YB_CHECK_CUDA(cudaMemsetAsync(levels[i].idx, 0xFF, sizeof(int64_t) * 3 * (size_t)nt, st));
  }
  YB_CHECK_CUDA(cudaMemsetAsync(state, 0, (size_t)cells, st));
  if (dense != nullptr) YB_CHECK_CUDA(cudaMemsetAsync(dense, 0, sizeof(float) * 6 * (size_t)cells, st));
  YB_CHECK_CUDA(cudaMemsetAsync(nobj, 0, sizeof(int) * 3, st));
  P.labels = labels;
  P.offsets = offsets;
  P.table = anchor_table;
  P.head_anchors = head_anchors;
  P.state = state;
  P.dense = dense;
  P.counts = counts;
  P.nobj = nobj;
  P.cap = cap;
  P.decay_base = decay_base;
  P.decay_stride = decay_stride;
  P.B = B;
  P.T = T;
  P.ignore_thr = ignore_thr;
  P.nt = nt;
  YB_REQUIRE(B > 0, "yolo_build_targets: empty batch");
  yolo_build_targets_kernel<<<(B + 63) / 64, 64, 0, st>>>(P);
  YB_LAUNCHED();
  return 0;
}
