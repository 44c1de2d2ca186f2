// ComputeLoss on the GPU (sm_100a; HBM/latency-bound integer + fp32 work, no tensor cores).
//
// Replaces reference ultralytics_loss.py:
//   build_targets  (:122-311)  anchor-ratio match + 5-offset neighbour expansion -> ordered (bit-exact) index lists
//   __call__       (:60-120)   gather pi[b,a,gj,gi], box decode, GIoU (utils/bboxes_utils.py:33-87), BCE-with-logits for the
//                              class one-hot and the dense objectness map with tobj[b,a,gj,gi] = GIoU.clamp(0) (last write wins)
// and the autograd backward of all of it, written straight into the head convolutions' bf16 NHWC gradient operand.
//
// Kernels (one launch each, all levels at once):
//   build_targets_kernel   one CTA per level; candidates enumerated in the reference's order (offset, anchor, target) and
//                          compacted with a block-wide ORDERED prefix sum (no atomics -> row order is bit-exact)
//   loss_rows_kernel       one warp per matched row: 85-float gather, decode, GIoU + analytic d(1-GIoU)/dlogits, class BCE;
//                          rows sharing a cell are chained (atomicExch) so the dense pass can find them
//   loss_obj_kernel        dense objectness BCE; tobj of a cell = GIoU of the HIGHEST row index in its chain
//                          (= "last write wins" of the reference's CPU index_put_, ultralytics_loss.py:89)
//   loss_finalize_kernel   fixed-order reductions -> [loss*bs, lbox, lobj, lcls]
//   loss_bwd_kernel        one warp per pixel: writes the complete gradient row (3 anchors x (5+nc) channels): objectness
//                          channel densely, box/class channels from the cell's chain (summed in ascending row order ->
//                          deterministic); outputs fp32 (B,na,H,W,no) and/or bf16 NHWC (B,H,W,cpad)
#include "../../include/yolov5m_b200.h"
#include "common.cuh"

#include <algorithm>

namespace yb {

static constexpr int kMaxLevels = 4;

struct LossKParams {
  yb_loss_level lv[kMaxLevels];
  int nl, B, na, no, nc;
  long cap;
  const int* counts;
  const int* nobj;   // optional: object rows per level (denominator of the box / class means) when the row lists also hold
                     // "ignore" rows or unused slots (YOLO_LOSS, yolo_loss.cu); NULL -> counts
  int nan_empty;     // 1: a level without object rows yields NaN box / class terms like the reference's mean() of an empty
                     // tensor (loss.py:211,232); 0: ComputeLoss semantics (the level contributes nothing, ultralytics_loss.py:76)
  float lam_box, lam_obj, lam_cls;
  const float* gout;
  int cpad;
  float* out;
  int obj_rows;
};

__device__ __forceinline__ float torch_remainder1(float a) {  // torch `a % 1` for floats (result takes the divisor's sign)
  float m = fmodf(a, 1.0f);
  if (m != 0.f && m < 0.f) m += 1.0f;
  return m;
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float bce_logits(float x, float t) {
  return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}

// ------------------------------------------------------------------------------------------------ build_targets
struct BTParams {
  yb_loss_level lv[kMaxLevels];
  const float* targets;
  const float* anchors;
  int nt, na;
  float anchor_t;
  long cap;
  int* counts;
};

__global__ void __launch_bounds__(1024) build_targets_kernel(const __grid_constant__ BTParams P) {
  __shared__ int warp_tot[32];
  __shared__ int running;
  const int lvl = blockIdx.x;
  const yb_loss_level& L = P.lv[lvl];
  const int H = L.H, W = L.W;
  const float gw = (float)W, gh = (float)H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long total = 5L * P.na * P.nt;
  if (tid == 0) running = 0;
  __syncthreads();
  for (long base = 0; base < total; base += blockDim.x) {
    const long idx = base + tid;
    bool take = false;
    float timg = 0, tc = 0, gx = 0, gy = 0, bw = 0, bh = 0;
    int a = 0, o = 0;
    if (idx < total) {
      const int t = (int)(idx % P.nt);
      const long r = idx / P.nt;
      a = (int)(r % P.na);
      o = (int)(r / P.na);
      const float* tg = P.targets + (long)t * 6;
      timg = tg[0];
      tc = tg[1];
      gx = tg[2] * gw;  // targets * gain, ultralytics_loss.py:175
      gy = tg[3] * gh;
      bw = tg[4] * gw;
      bh = tg[5] * gh;
      const float aw = P.anchors[(lvl * P.na + a) * 2 + 0], ah = P.anchors[(lvl * P.na + a) * 2 + 1];
      const float rw = __fdiv_rn(bw, aw), rh = __fdiv_rn(bh, ah);  // :186
      const float iw = __fdiv_rn(1.0f, rw), ih = __fdiv_rn(1.0f, rh);
      const float m = fmaxf(fmaxf(rw, iw), fmaxf(rh, ih));
      const bool nan = (rw != rw) || (rh != rh) || (iw != iw) || (ih != ih);  // torch.max propagates NaN -> compare false
      take = !nan && (m < P.anchor_t);                                          // :195
      if (take && o > 0) {
        float v;
        if (o == 1) v = gx;
        else if (o == 2) v = gy;
        else if (o == 3) v = __fsub_rn(gw, gx);  // gxi = gain[[2,3]] - gxy, :222-226
        else v = __fsub_rn(gh, gy);
        take = (torch_remainder1(v) < 0.5f) && (v > 1.0f);  // :233, :243
      }
    }
    // ordered block compaction
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int before = running;
    for (int w2 = 0; w2 < warp; ++w2) before += warp_tot[w2];
    const int rank = before + __popc(bal & ((1u << lane) - 1u));
    if (take && rank < P.cap) {
      const float offx = (o == 1) ? 0.5f : (o == 3 ? -0.5f : 0.f);
      const float offy = (o == 2) ? 0.5f : (o == 4 ? -0.5f : 0.f);
      long gi = (long)truncf(__fsub_rn(gx, offx));  // (gxy - offsets).long(), :278
      long gj = (long)truncf(__fsub_rn(gy, offy));
      gi = gi < 0 ? 0 : (gi > W - 1 ? W - 1 : gi);  // clamp_ (in place: tbox sees the clamped cell), :285
      gj = gj < 0 ? 0 : (gj > H - 1 ? H - 1 : gj);
      L.idx[0 * P.cap + rank] = (long)truncf(timg);
      L.idx[1 * P.cap + rank] = a;
      L.idx[2 * P.cap + rank] = gj;
      L.idx[3 * P.cap + rank] = gi;
      float4 tb;
      tb.x = __fsub_rn(gx, (float)gi);  // :296
      tb.y = __fsub_rn(gy, (float)gj);
      tb.z = bw;
      tb.w = bh;
      reinterpret_cast<float4*>(L.tbox)[rank] = tb;
      L.anch[2 * rank + 0] = P.anchors[(lvl * P.na + a) * 2 + 0];  // :301
      L.anch[2 * rank + 1] = P.anchors[(lvl * P.na + a) * 2 + 1];
      L.tcls[rank] = (long)truncf(tc);  // :306
    }
    __syncthreads();
    if (tid == 0) {
      int s = running;
      for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) s += warp_tot[w2];
      running = s;
    }
    __syncthreads();
  }
  if (tid == 0) P.counts[lvl] = running < P.cap ? running : (int)P.cap;
}

// ------------------------------------------------------------------------------------------------ per-row pass
// row_val[0][k] = GIoU.clamp(0) (tobj), row_val[1][k] = 1 - GIoU, row_val[2][k] = sum_c BCE(pcls_c, onehot_c)
// row_grad[k][0..3] = d(1-GIoU)/d(logit xywh), [4] = 0, [5+c] = sigmoid(pcls_c) - onehot_c
__global__ void __launch_bounds__(256) loss_rows_kernel(const __grid_constant__ LossKParams P) {
  const int lane = threadIdx.x & 31;
  const long wglobal = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long wi = wglobal; wi < (long)P.nl * P.cap; wi += nwarps) {
    const int lvl = (int)(wi / P.cap);
    const long k = wi - (long)lvl * P.cap;
    if (k >= P.counts[lvl]) continue;
    const yb_loss_level& L = P.lv[lvl];
    const long b = L.idx[k], a = L.idx[P.cap + k], gj = L.idx[2 * P.cap + k], gi = L.idx[3 * P.cap + k];
    if (b < 0 || b >= P.B) {  // unused slot (YOLO_LOSS) or an image index the reference would raise IndexError on
      if (lane == 0) {
        L.row_val[k] = 0.f;
        L.row_val[P.cap + k] = 0.f;
        L.row_val[2 * P.cap + k] = 0.f;
      }
      continue;
    }
    const long cell = ((b * P.na + a) * L.H + gj) * L.W + gi;
    const float* ps = L.p + cell * P.no;
    float* rg = L.row_grad + k * P.no;
    if (L.tcls[k] < 0) {
      // "ignore" row of YOLO_LOSS (loss.py:190): objectness target -1, no box / class term
      for (int c = lane; c < P.no; c += 32) rg[c] = 0.f;
      if (lane == 0) {
        L.row_val[k] = -1.f;
        L.row_val[P.cap + k] = 0.f;
        L.row_val[2 * P.cap + k] = 0.f;
        L.row_prev[k] = atomicExch(&L.cell_head[cell], (int)k + 1);
      }
      continue;
    }
    // ---- class BCE (ultralytics_loss.py:93-95)
    float csum = 0.f;
    if (P.nc > 1) {
      const int tc = (int)L.tcls[k];
      for (int c = lane; c < P.nc; c += 32) {
        const float x = ps[5 + c];
        const float t = (c == tc) ? 1.f : 0.f;
        csum += bce_logits(x, t);
        rg[5 + c] = sigmoid_acc(x) - t;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
    } else {
      for (int c = lane; c < P.nc; c += 32) rg[5 + c] = 0.f;
    }
    if (lane == 0) {
      // ---- box decode (:81-82) and GIoU (bboxes_utils.py:46-86), midpoint format
      const float sx = sigmoid_acc(ps[0]), sy = sigmoid_acc(ps[1]), sw = sigmoid_acc(ps[2]), sh = sigmoid_acc(ps[3]);
      const float aw = L.anch[2 * k], ah = L.anch[2 * k + 1];
      const float px = sx * 2.f - 0.5f, py = sy * 2.f - 0.5f;
      const float pw = (sw * 2.f) * (sw * 2.f) * aw, ph = (sh * 2.f) * (sh * 2.f) * ah;
      const float4 tb = reinterpret_cast<const float4*>(L.tbox)[k];
      const float b1x1 = px - pw / 2.f, b1x2 = px + pw / 2.f, b1y1 = py - ph / 2.f, b1y2 = py + ph / 2.f;
      const float b2x1 = tb.x - tb.z / 2.f, b2x2 = tb.x + tb.z / 2.f, b2y1 = tb.y - tb.w / 2.f, b2y2 = tb.y + tb.w / 2.f;
      const float w1 = b1x2 - b1x1, h1 = b1y2 - b1y1, w2 = b2x2 - b2x1, h2 = b2y2 - b2y1;
      const float iwr = fminf(b1x2, b2x2) - fmaxf(b1x1, b2x1), ihr = fminf(b1y2, b2y2) - fmaxf(b1y1, b2y1);
      const float iw = fmaxf(iwr, 0.f), ih = fmaxf(ihr, 0.f);
      const float inter = iw * ih;
      const float uni = w1 * h1 + w2 * h2 - inter + 1e-7f;
      const float iou = inter / uni;
      const float cw = fmaxf(b1x2, b2x2) - fminf(b1x1, b2x1), ch = fmaxf(b1y2, b2y2) - fminf(b1y1, b2y1);
      const float carea = cw * ch + 1e-7f;
      const float giou = iou - (carea - uni) / carea;
      L.row_val[k] = fmaxf(giou, 0.f);
      L.row_val[P.cap + k] = 1.0f - giou;
      L.row_val[2 * P.cap + k] = csum;
      // ---- d(giou): giou = I/U - 1 + U/C
      const float gI = 1.f / uni + inter / (uni * uni) - 1.f / carea;  // dU/dI = -1 folded in
      const float gU = -inter / (uni * uni) + 1.f / carea;
      const float gC = -uni / (carea * carea);
      const float giw = (iwr >= 0.f) ? gI * ih : 0.f, gih = (ihr >= 0.f) ? gI * iw : 0.f;
      const float gcw = gC * ch, gch = gC * cw;
      const float gw1 = gU * h1, gh1 = gU * w1;
      auto lt = [](float p, float q) { return p < q ? 1.f : (p == q ? 0.5f : 0.f); };  // torch min/max split ties evenly
      // x: iw_raw = min(b1x2,b2x2) - max(b1x1,b2x1); cw = max(b1x2,b2x2) - min(b1x1,b2x1); w1 = b1x2 - b1x1
      const float gx2 = giw * lt(b1x2, b2x2) + gcw * lt(b2x2, b1x2) + gw1;
      const float gx1 = -giw * lt(b2x1, b1x1) - gcw * lt(b1x1, b2x1) - gw1;
      const float gy2 = gih * lt(b1y2, b2y2) + gch * lt(b2y2, b1y2) + gh1;
      const float gy1 = -gih * lt(b2y1, b1y1) - gch * lt(b1y1, b2y1) - gh1;
      const float gpx = gx1 + gx2, gpw = 0.5f * (gx2 - gx1), gpy = gy1 + gy2, gph = 0.5f * (gy2 - gy1);
      // chain to logits; loss row = 1 - giou  => sign flip
      rg[0] = -gpx * 2.f * sx * (1.f - sx);
      rg[1] = -gpy * 2.f * sy * (1.f - sy);
      rg[2] = -gpw * aw * 8.f * sw * sw * (1.f - sw);
      rg[3] = -gph * ah * 8.f * sh * sh * (1.f - sh);
      rg[4] = 0.f;
      L.row_prev[k] = atomicExch(&L.cell_head[cell], (int)k + 1);
    }
  }
}

// highest row index in the chain that starts at `head` (1-based link values; 0 terminates)
__device__ __forceinline__ int chain_last(const int* __restrict__ prev, int head) {
  int best = 0;
  for (int r = head; r != 0; r = prev[r - 1]) best = r > best ? r : best;
  return best;
}

// ------------------------------------------------------------------------------------------------ dense objectness BCE
__global__ void __launch_bounds__(256) loss_obj_kernel(const __grid_constant__ LossKParams P) {
  __shared__ float red[8];
  float acc[kMaxLevels];
#pragma unroll
  for (int l = 0; l < kMaxLevels; ++l) acc[l] = 0.f;
#pragma unroll
  for (int lvl = 0; lvl < kMaxLevels; ++lvl) {
    if (lvl >= P.nl) break;
    const yb_loss_level& L = P.lv[lvl];
    const long cells = (long)P.B * P.na * L.H * L.W;
    for (long c = (long)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (long)gridDim.x * blockDim.x) {
      const float x = L.p[c * P.no + 4];
      const int head = L.cell_head[c];
      float t = 0.f;
      if (head != 0) t = L.row_val[chain_last(L.row_prev, head) - 1];
      acc[lvl] += bce_logits(x, t);  // :101
    }
  }
#pragma unroll
  for (int lvl = 0; lvl < kMaxLevels; ++lvl) {
    if (lvl >= P.nl) break;
    float v = acc[lvl];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
      P.lv[lvl].obj_partial[blockIdx.x] = s;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ finalize
__device__ double block_sum_ordered(const float* __restrict__ v, long n, double* sh) {
  double a = 0.0;
  for (long i = threadIdx.x; i < n; i += blockDim.x) a += (double)v[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256) loss_finalize_kernel(const __grid_constant__ LossKParams P) {
  __shared__ double sh[256];
  double lbox = 0.0, lobj = 0.0, lcls = 0.0;
  for (int lvl = 0; lvl < P.nl; ++lvl) {
    const yb_loss_level& L = P.lv[lvl];
    const int nrows = P.counts[lvl];
    const int n = P.nobj != nullptr ? P.nobj[lvl] : nrows;
    const double sb = block_sum_ordered(L.row_val + P.cap, nrows, sh);
    const double sc = block_sum_ordered(L.row_val + 2 * P.cap, nrows, sh);
    const double so = block_sum_ordered(L.obj_partial, P.obj_rows, sh);
    if (n > 0) {
      lbox += sb / n;                                   // (1 - iou).mean(), :85
      if (P.nc > 1) lcls += sc / ((double)n * P.nc);    // BCEcls mean over n*nc, :95
    } else if (P.nan_empty) {
      lbox += nan("");
      lcls += nan("");
    }
    lobj += so / ((double)P.B * P.na * L.H * L.W) * (double)L.balance;  // :101-102
  }
  if (threadIdx.x == 0) {
    const float fb = (float)lbox * P.lam_box, fo = (float)lobj * P.lam_obj, fc = (float)lcls * P.lam_cls;  // :104-106
    P.out[0] = (fb + fo + fc) * (float)P.B;                                                              // :120
    P.out[1] = fb;
    P.out[2] = fo;
    P.out[3] = fc;
  }
}

// ------------------------------------------------------------------------------------------------ backward
// one warp per pixel (b, gj, gi) of a level.
__global__ void __launch_bounds__(256) loss_bwd_kernel(const __grid_constant__ LossKParams P) {
  const int lane = threadIdx.x & 31;
  const long wglobal = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const float g = (P.gout != nullptr ? P.gout[0] : 1.f) * (float)P.B;
  long pix_total = 0;
  long lvl_begin[kMaxLevels + 1];
#pragma unroll
  for (int l = 0; l < kMaxLevels; ++l) {
    lvl_begin[l] = pix_total;
    if (l < P.nl) pix_total += (long)P.B * P.lv[l].H * P.lv[l].W;
  }
  lvl_begin[kMaxLevels] = pix_total;
  // A warp takes 32 consecutive pixels at a time: lane l first loads, for pixel base + l, the chain heads and objectness
  // logits of its anchors (the only loads on the common path) -- 32 pixels' worth of independent loads in flight per warp
  // instead of one pixel's (the pass was bound by that latency: 0.33 ms for 0.5 GB) -- then the warp walks the 32 pixels and
  // writes each complete gradient row.
  constexpr int kMaxNa = 4;
  for (long base = wglobal * 32; base < pix_total; base += nwarps * 32) {
    int my_lvl = 0, my_heads[kMaxNa];
    long my_b = 0, my_sp = 0;
    float my_xs[kMaxNa];
#pragma unroll
    for (int a = 0; a < kMaxNa; ++a) {
      my_heads[a] = 0;
      my_xs[a] = 0.f;
    }
    {
      const long wl = base + lane;
      if (wl < pix_total) {
#pragma unroll
        for (int l = 1; l < kMaxLevels; ++l)
          if (l < P.nl && wl >= lvl_begin[l]) my_lvl = l;
        const yb_loss_level& Ll = P.lv[my_lvl];
        const long pixl = wl - lvl_begin[my_lvl];
        const long hwl = (long)Ll.H * Ll.W;
        my_b = pixl / hwl;
        my_sp = pixl - my_b * hwl;
#pragma unroll
        for (int a = 0; a < kMaxNa; ++a)
          if (a < P.na) {
            const long cell = (my_b * P.na + a) * hwl + my_sp;
            my_heads[a] = Ll.cell_head[cell];
            my_xs[a] = Ll.p[cell * P.no + 4];
          }
        // sigmoid of this lane's own cells, once (the pixel loop below used to evaluate it on every lane for every pixel)
#pragma unroll
        for (int a = 0; a < kMaxNa; ++a)
          if (a < P.na) my_xs[a] = sigmoid_acc(my_xs[a]);
      }
    }
    const int npx = (int)min(32L, pix_total - base);
    for (int pj = 0; pj < npx; ++pj) {
    const long wi = base + pj;
    const int lvl = __shfl_sync(0xffffffffu, my_lvl, pj);
    const yb_loss_level& L = P.lv[lvl];
    const long pix = wi - lvl_begin[lvl];
    const long hw = (long)L.H * L.W;
    const long b = __shfl_sync(0xffffffffu, my_b, pj), sp = __shfl_sync(0xffffffffu, my_sp, pj);
    const int n = P.nobj != nullptr ? P.nobj[lvl] : P.counts[lvl];
    const float cbox = n > 0 ? g * P.lam_box / (float)n : 0.f;
    const float ccls = (n > 0 && P.nc > 1) ? g * P.lam_cls / ((float)n * (float)P.nc) : 0.f;
    const float cobj = g * P.lam_obj * L.balance / (float)((long)P.B * P.na * hw);
    // channel ownership for the bf16 NHWC row: lane owns channels lane*8 .. lane*8+7 (cpad <= 256)
    float v8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v8[j] = 0.f;
    int heads[kMaxNa];
    float xs[kMaxNa];
#pragma unroll
    for (int a = 0; a < kMaxNa; ++a) {
      heads[a] = __shfl_sync(0xffffffffu, my_heads[a], pj);
      xs[a] = __shfl_sync(0xffffffffu, my_xs[a], pj);
    }
#pragma unroll
    for (int a = 0; a < kMaxNa; ++a) {
      if (a >= P.na) break;
      const long cell = (b * P.na + a) * hw + sp;
      const int head = heads[a];
      const float x = xs[a];
      float tobj = 0.f;
      // fp32 row of this anchor: lane owns o = lane, lane+32, lane+64
      float f0 = 0.f, f1 = 0.f, f2 = 0.f;
      if (head != 0) {
        int last = 0;
        // ascending row order: repeatedly take the smallest chain member greater than `last` (chains are short)
        while (true) {
          int nxt = 0x7fffffff;
          for (int r = head; r != 0; r = L.row_prev[r - 1])
            if (r > last && r < nxt) nxt = r;
          if (nxt == 0x7fffffff) break;
          last = nxt;
          const float* rg = L.row_grad + (long)(nxt - 1) * P.no;
          if (L.grad_f32 != nullptr) {
            if (lane < P.no) f0 += rg[lane] * (lane < 4 ? cbox : ccls);
            if (lane + 32 < P.no) f1 += rg[lane + 32] * ccls;
            if (lane + 64 < P.no) f2 += rg[lane + 64] * ccls;
          }
          if (L.grad_bf16 != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int o = lane * 8 + j - a * P.no;
              if (o >= 0 && o < P.no && o != 4) v8[j] += rg[o] * (o < 4 ? cbox : ccls);
            }
          }
        }
        tobj = L.row_val[last - 1];
      }
      const float gobj = (x - tobj) * cobj;  // x = sigmoid(objectness logit), evaluated by the owning lane above
      if (L.grad_f32 != nullptr) {
        float* go = L.grad_f32 + cell * P.no;
        if (lane == 4) f0 = gobj;
        if (lane < P.no) go[lane] = f0;
        if (lane + 32 < P.no) go[lane + 32] = f1;
        if (lane + 64 < P.no) go[lane + 64] = f2;
      }
      if (L.grad_bf16 != nullptr) {
        const int co = a * P.no + 4;
        if ((co >> 3) == lane) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j == (co & 7)) v8[j] = gobj;
        }
      }
    }
    if (L.grad_bf16 != nullptr && lane * 8 < P.cpad) {
      uint4 o4;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o4);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v8[2 * j], v8[2 * j + 1]);
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(L.grad_bf16) + pix * P.cpad + lane * 8) = o4;
    }
    }  // pixels of this batch
  }
}

// ------------------------------------------------------------------------------------------------ standalone IoU / GIoU
// utils/bboxes_utils.py:33-87: (n,4) x (n,4) -> (n,), same operation order as the reference
__global__ void box_iou_kernel(const float* __restrict__ a, const float* __restrict__ b, long n, int midpoint, int giou,
                               float eps, float* __restrict__ out) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float4 p = reinterpret_cast<const float4*>(a)[i], t = reinterpret_cast<const float4*>(b)[i];
    float b1x1, b1y1, b1x2, b1y2, b2x1, b2y1, b2x2, b2y2;
    if (midpoint) {
      b1x1 = __fsub_rn(p.x, p.z / 2.f); b1y1 = __fsub_rn(p.y, p.w / 2.f); b1x2 = __fadd_rn(p.x, p.z / 2.f); b1y2 = __fadd_rn(p.y, p.w / 2.f);
      b2x1 = __fsub_rn(t.x, t.z / 2.f); b2y1 = __fsub_rn(t.y, t.w / 2.f); b2x2 = __fadd_rn(t.x, t.z / 2.f); b2y2 = __fadd_rn(t.y, t.w / 2.f);
    } else {
      b1x1 = p.x; b1y1 = p.y; b1x2 = p.z; b1y2 = p.w;
      b2x1 = t.x; b2y1 = t.y; b2x2 = t.z; b2y2 = t.w;
    }
    const float w1 = __fsub_rn(b1x2, b1x1), h1 = __fsub_rn(b1y2, b1y1), w2 = __fsub_rn(b2x2, b2x1), h2 = __fsub_rn(b2y2, b2y1);
    const float iw = fmaxf(__fsub_rn(fminf(b1x2, b2x2), fmaxf(b1x1, b2x1)), 0.f);
    const float ih = fmaxf(__fsub_rn(fminf(b1y2, b2y2), fmaxf(b1y1, b2y1)), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(w1, h1), __fmul_rn(w2, h2)), inter), eps);
    const float iou = __fdiv_rn(inter, uni);
    float r = iou;
    if (giou) {
      const float cw = __fsub_rn(fmaxf(b1x2, b2x2), fminf(b1x1, b2x1)), ch = __fsub_rn(fmaxf(b1y2, b2y2), fminf(b1y1, b2y1));
      const float carea = __fadd_rn(__fmul_rn(cw, ch), eps);
      r = __fsub_rn(iou, __fdiv_rn(__fsub_rn(carea, uni), carea));
    }
    out[i] = r;
  }
}

static int fill_params(LossKParams& P, const yb_loss_level* levels, int nl, int B, int na, int no, int64_t cap,
                       const int* counts) {
  YB_REQUIRE(nl >= 1 && nl <= kMaxLevels, "loss: nl=%d (max %d)", nl, kMaxLevels);
  YB_REQUIRE(no >= 6 && no <= 96, "loss: no=%d unsupported (5+nc must be in [6,96])", no);
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < nl; ++i) P.lv[i] = levels[i];
  P.nl = nl;
  P.B = B;
  P.na = na;
  P.no = no;
  P.nc = no - 5;
  P.cap = cap;
  P.counts = counts;
  return 0;
}

static int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
      sms = 148;
  }
  return sms;
}

}  // namespace yb

using namespace yb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int yb_loss_obj_rows(void) { return sm_count() * 8; }

int yb_box_iou(const float* boxes_preds, const float* boxes_labels, int64_t n, int midpoint, int giou, float eps, float* out,
               void* stream) {
  if (n == 0) return 0;
  const int blocks = (int)std::max<long>(1, std::min<long>((n + 255) / 256, (long)sm_count() * 8));
  box_iou_kernel<<<blocks, 256, 0, ST(stream)>>>(boxes_preds, boxes_labels, n, midpoint, giou, eps, out);
  YB_LAUNCHED();
  return 0;
}

int yb_build_targets(const float* targets, int nt, const float* anchors, const yb_loss_level* levels, int nl, int na,
                     float anchor_t, int64_t cap, int* counts, void* stream) {
  YB_REQUIRE(nl >= 1 && nl <= kMaxLevels, "build_targets: nl=%d", nl);
  if (nt == 0) {  // ultralytics_loss.py:262-265: no rows on any level
    YB_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * nl, ST(stream)));
    return 0;
  }
  YB_REQUIRE(cap >= 5L * na * nt, "build_targets: cap=%lld < 5*na*nt", (long long)cap);
  BTParams P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < nl; ++i) P.lv[i] = levels[i];
  P.targets = targets;
  P.anchors = anchors;
  P.nt = nt;
  P.na = na;
  P.anchor_t = anchor_t;
  P.cap = cap;
  P.counts = counts;
  build_targets_kernel<<<nl, 1024, 0, ST(stream)>>>(P);
  YB_LAUNCHED();
  return 0;
}

int yb_loss_fwd(const yb_loss_level* levels, int nl, int B, int na, int no, int64_t cap, const int* counts,
                const int* nobj, int nan_on_empty, float lam_box, float lam_obj, float lam_cls, float* out4, void* stream) {
  LossKParams P;
  if (fill_params(P, levels, nl, B, na, no, cap, counts)) return -1;
  P.nobj = nobj;
  P.nan_empty = nan_on_empty;
  P.lam_box = lam_box;
  P.lam_obj = lam_obj;
  P.lam_cls = lam_cls;
  P.out = out4;
  P.obj_rows = yb_loss_obj_rows();
  for (int i = 0; i < nl; ++i)
    YB_CHECK_CUDA(cudaMemsetAsync(levels[i].cell_head, 0, sizeof(int) * (size_t)B * na * levels[i].H * levels[i].W,
                                  ST(stream)));
  if (cap > 0) {
    const long warps = (long)nl * cap;
    const int blocks = (int)std::max<long>(1, std::min<long>((warps + 7) / 8, (long)sm_count() * 8));
    loss_rows_kernel<<<blocks, 256, 0, ST(stream)>>>(P);
    YB_LAUNCHED();
  }
  loss_obj_kernel<<<P.obj_rows, 256, 0, ST(stream)>>>(P);
  YB_LAUNCHED();
  loss_finalize_kernel<<<1, 256, 0, ST(stream)>>>(P);
  YB_LAUNCHED();
  return 0;
}

int yb_loss_bwd(const yb_loss_level* levels, int nl, int B, int na, int no, int64_t cap, const int* counts,
                const int* nobj, float lam_box, float lam_obj, float lam_cls, const float* gout, int cpad, void* stream) {
  LossKParams P;
  if (fill_params(P, levels, nl, B, na, no, cap, counts)) return -1;
  P.nobj = nobj;
  P.lam_box = lam_box;
  P.lam_obj = lam_obj;
  P.lam_cls = lam_cls;
  P.gout = gout;
  P.cpad = cpad;
  for (int i = 0; i < nl; ++i)
    YB_REQUIRE(levels[i].grad_bf16 == nullptr || (cpad % 8 == 0 && cpad >= na * no && cpad <= 256),
               "loss_bwd: cpad=%d must be a multiple of 8 in [na*no, 256]", cpad);
  YB_REQUIRE(na <= 4, "loss_bwd: na=%d > 4 anchors per level", na);
  long pix = 0;
  for (int i = 0; i < nl; ++i) pix += (long)B * levels[i].H * levels[i].W;
  const int blocks = (int)std::max<long>(1, std::min<long>((pix + 7) / 8, (long)sm_count() * 16));
  loss_bwd_kernel<<<blocks, 256, 0, ST(stream)>>>(P);
  YB_LAUNCHED();
  return 0;
}

}  // extern "C"
